"""Loader for the unmodified reference package copied to ``oracle/_ref/`` (see ``oracle/make_ref.py``).

TEST INFRASTRUCTURE.  ``load()`` returns the imported ``pyAudioDspTools`` module or None when the copy is
absent; the product never calls this (tests/test_host_logic.py::test_product_never_imports_the_oracle).
"""
import contextlib
import importlib
import io
import os
import sys

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mod = None


def available() -> bool:
    return os.path.isfile(os.path.join(_REF_DIR, "pyAudioDspTools", "__init__.py"))


def load():
    global _mod
    if _mod is None and available():
        sys.path.insert(0, _REF_DIR)
        try:
            with contextlib.redirect_stdout(io.StringIO()):      # the "No cupy" info line (__init__.py:8)
                _mod = importlib.import_module("pyAudioDspTools")
        finally:
            sys.path.remove(_REF_DIR)
    return _mod
