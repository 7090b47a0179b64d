"""Recipe for ``oracle/_ref/`` — the UNMODIFIED reference package as a CPU baseline that can travel.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference is pure Python, so "building" it is a copy:
``/root/reference/pyAudioDspTools/*.py`` -> ``oracle/_ref/pyAudioDspTools/``.  ``oracle/_ref/`` is git-ignored
(reference sources never enter the history) but not gpurun-ignored, so the copy reaches the GPU box where
``/root/reference`` does not exist; ``bench.py --impl reference`` and the ``cpu_baseline`` leg then time the
real ``EffectFFTFilter.apply`` / ``EffectEQ3BandFFT.apply`` (``cpu_baseline.kind = "reference"``).
Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present; a no-op otherwise.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ADT_REFERENCE_DIR", "/root/reference")


def make_ref(verbose=True):
    src = os.path.join(SRC, "pyAudioDspTools")
    dst = os.path.join(HERE, "_ref", "pyAudioDspTools")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/_ref: {src} not present, nothing to do")
        return os.path.isdir(dst)
    os.makedirs(dst, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
            n += 1
    if verbose:
        print(f"oracle/_ref: copied {n} files of the unmodified reference package")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
