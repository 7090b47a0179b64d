import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_sessionstart(session):
    """The C-ABI library is git-ignored (built in-tree, it travels to the GPU box with the snapshot): on a
    fresh checkout build it once (nvcc cross-compiles for sm_100a without a GPU, ~90 s)."""
    lib = os.path.join(ROOT, "pyaudiodsptools_b200", "libadt_b200.so")
    if not os.path.exists(lib) and not os.environ.get("ADT_LIB_PATH"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pyaudiodsptools_b200", "csrc"), "-j4"])


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def golden_fft_cases():
    names = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        n = os.path.basename(p)[:-4]
        if n.startswith(("lowcut", "highcut", "eq3fft")) and not n.endswith("int16"):
            names.append(n)
    return names


def rms(a):
    a = np.asarray(a, dtype=np.float64)
    return float(np.sqrt(np.mean(a * a))) if a.size else 0.0


@pytest.fixture(scope="session")
def gpu_lib():
    """The ctypes-loaded C-ABI library on a machine with a GPU (fails loudly otherwise)."""
    from pyaudiodsptools_b200 import _native
    lib = _native.load()
    n = _native.device_count()
    if n < 1:
        # a plain `pytest tests` on a CPU box skips the GPU tier; where a GPU is expected (the driver's
        # `-m gpu` run) set ADT_REQUIRE_GPU=1 to turn a missing device into a failure instead of a skip
        if os.environ.get("ADT_REQUIRE_GPU") == "1":
            pytest.fail("gpu-marked test running without a CUDA device")
        pytest.skip("no CUDA device (GPU tier; set ADT_REQUIRE_GPU=1 to fail instead)")
    return lib
