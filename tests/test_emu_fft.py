"""Development check of the CUDA kernel's index algebra on the CPU.

tests/emu/emu_fir.cpp compiles the kernel's own __host__ __device__ phase
functions (pyaudiodsptools_b200/csrc/fft_core.cuh) with g++ and runs them with
a loop over thread ids.  This is NOT a product path (the product fails loudly
without a GPU) and NOT the oracle; it only lets the decomposition
N = N1*N2*32, the twiddle tables, the in-place tile exchanges and the mask
permutation be verified against numpy.fft in the CPU test tier.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
SO = os.path.join(EMU_DIR, "libemu_fir.so")
SRC = [os.path.join(EMU_DIR, "emu_fir.cpp"),
       os.path.join(HERE, "..", "pyaudiodsptools_b200", "csrc", "fft_core.cuh"),
       os.path.join(HERE, "..", "pyaudiodsptools_b200", "csrc", "fft_core16.cuh"),
       os.path.join(HERE, "..", "pyaudiodsptools_b200", "csrc", "fir_tables.h")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRC):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC[0]])
    lib = ctypes.CDLL(SO)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.emu_fir_block.argtypes = [ctypes.c_int, fp, fp, ctypes.c_longlong, ctypes.c_longlong, fp, ctypes.c_int, fp]
    lib.emu_dft.argtypes = [ctypes.c_int, ctypes.c_int, fp]
    lib.emu_fir16_block.argtypes = lib.emu_fir_block.argtypes
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.mark.parametrize("r", [2, 4, 8, 16, 32])
@pytest.mark.parametrize("direction", [-1, 1])
def test_register_dft(emu, r, direction):
    rng = np.random.default_rng(r)
    x = (rng.standard_normal(r) + 1j * rng.standard_normal(r)).astype(np.complex64)
    buf = np.ascontiguousarray(x.view(np.float32)).copy()
    emu.emu_dft(r, direction, _p(buf))
    got = buf.view(np.complex64)
    want = np.fft.fft(x.astype(np.complex128)) if direction < 0 else np.fft.ifft(x.astype(np.complex128)) * r
    assert np.max(np.abs(got - want)) < 2e-6 * np.sqrt(r) * np.max(np.abs(want))


@pytest.mark.parametrize("n,variant", [(2048, 32), (4096, 32), (8192, 32), (16384, 32), (32768, 32), (4096, 16), (8192, 16)])
@pytest.mark.parametrize("real_mask", [0, 1])
def test_fir_block_equals_circular_convolution(emu, n, variant, real_mask):
    """variant = complex points held per thread (32: fft_core.cuh, 16: fft_core16.cuh)."""
    run = emu.emu_fir_block if variant == 32 else emu.emu_fir16_block
    rng = np.random.default_rng(n + real_mask)
    n_in = 3 * n
    xa = rng.uniform(-1, 1, n_in).astype(np.float32)
    xb = rng.uniform(-1, 1, n_in).astype(np.float32)
    if real_mask:
        H = rng.uniform(0, 1, n) + 0j
    else:
        H = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    Hc = np.ascontiguousarray(H.astype(np.complex64).view(np.float32))
    for ws in (n // 2 + 3, -100, n_in - n // 3):   # interior, left edge, right edge (zero fill)
        z = np.zeros(2 * n, dtype=np.float32)
        assert run(n, _p(xa), _p(xb), n_in, ws, _p(Hc), real_mask, _p(z)) == 0
        idx = ws + np.arange(n)
        ok = (idx >= 0) & (idx < n_in)
        w = np.where(ok, xa[np.clip(idx, 0, n_in - 1)], 0) + 1j * np.where(ok, xb[np.clip(idx, 0, n_in - 1)], 0)
        want = np.fft.ifft(np.fft.fft(w.astype(np.complex128)) * H.astype(np.complex64).astype(np.complex128))
        got = z.view(np.complex64)
        scale = np.sqrt(np.mean(np.abs(want) ** 2))
        err = np.sqrt(np.mean(np.abs(got - want) ** 2)) / scale
        print(n, ws, err)
        assert err < (2e-6 if n > 16384 else 1e-6), (n, ws, err)   # float32 FFT pair, twiddles from power chains
