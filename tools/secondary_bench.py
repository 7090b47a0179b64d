#!/usr/bin/env python
"""Throughput of the widened rows: streaming biquad (N1), wave-shapers and delay (N4).
Device-resident buffers, CUDA-event timing through the C ABI; CPU figures are the numpy/python oracle."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import pyaudiodsptools_b200 as adt
from pyaudiodsptools_b200 import _native, devices

ctx = _native.default_context(0)
lib = ctx.lib


def timed(fn, reps=5):
    fn(); ctx.sync()
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    return e0.elapsed_ms(e1) / reps


# ---- biquad: 1000 channels x 10 s, one band, float32 ------------------------------------------------
chans, n = 1000, 441000
x = np.random.default_rng(0).uniform(-1, 1, (chans, n)).astype(np.float32)
dx, dy = ctx.malloc(x.nbytes), ctx.malloc(x.nbytes)
ctx.h2d(dx, x)
coef = devices.biquad_coefficients(100, 2, 700, -4, 8000, 5)[1]
h = C.c_void_p()
ctx.check(lib.adt_biquad_create(ctx.h, (C.c_double * 5)(*coef), chans, 0, C.byref(h)))
ms = timed(lambda: ctx.check(lib.adt_biquad_apply_dev(h, dx, dy, n, n)), reps=3)
y = np.empty((1, n), np.float32); ctx.d2h(y, dy)
o = oracle.Eq3BandBiquad(100, 2, 700, -4, 8000, 5)
t0 = time.perf_counter(); want = o.applymidband(x[0, :20000].copy()); cpu = 20000 / (time.perf_counter() - t0) / 1e6
print(f"biquad (peaking band, f32, bit-exact recurrence): {chans} ch x {n}: {ms:.2f} ms -> {chans * n / ms / 1e3:.0f} Msamples/s "
      f"({8 * chans * n / ms / 1e6:.0f} GB/s);  python reference loop: {cpu:.2f} Msamples/s per core")
# (state continues across the timed repetitions, so compare a fresh object for parity)
h2 = C.c_void_p(); ctx.check(lib.adt_biquad_create(ctx.h, (C.c_double * 5)(*coef), chans, 0, C.byref(h2)))
ctx.check(lib.adt_biquad_apply_dev(h2, dx, dy, n, n)); ctx.d2h(y, dy)
assert np.array_equal(y[0, :20000], want), "biquad parity"

# ---- shapers: pointwise over the same buffer ------------------------------------------------------------
for name, dev in (("saturator", adt.CreateSaturator()), ("soft clipper", adt.CreateSoftClipper())):
    p = dev._params()
    ms = timed(lambda: ctx.check(lib.adt_shape_apply_dev(ctx.h, dev.kind, p.ctypes.data, dx, dy, chans * n)))
    print(f"{name}: {chans * n / ms / 1e3:.0f} Msamples/s ({8 * chans * n / ms / 1e6:.0f} GB/s, pointwise)")

# ---- fused epilogue: FIR + saturator in one kernel vs FIR alone ----------------------------------------
adt.config.initialize(44100, 4096)
n_out = -(-n // 4096) * 4096
dz = ctx.malloc(chans * n_out * 4)
plain = adt.CreateLowCutFilter(800)
fused = adt.CreateLowCutFilter(800, epilogue=adt.CreateSaturator())
for name, d in (("FIR alone", plain), ("FIR + fused saturator", fused)):
    ms = timed(lambda: d.process_device(dx, n, n, dz, n_out, n_out, chans), reps=20)
    print(f"{name}: {ms:.3f} ms -> {chans * n_out / ms / 1e3:.0f} Msamples/s")

# ---- delay: 500 ms, 2 feedback loops, chunk 4096, 1000 channels ----------------------------------------
dl = C.c_void_p()
ramp = np.linspace(0.5, 0.1, num=2, dtype="float32")
ctx.check(lib.adt_delay_create(ctx.h, 22050, 2, ramp.ctypes.data, 0, chans, C.byref(dl)))
c = 4096
dxc, dyc = ctx.malloc(chans * c * 4), ctx.malloc(chans * c * 4)
ms = timed(lambda: ctx.check(lib.adt_delay_apply_dev(dl, dxc, dyc, c)), reps=50)
print(f"delay (500 ms, 2 loops) {chans} ch x {c}: {ms * 1e3:.1f} us per chunk -> {chans * c / ms / 1e3:.0f} Msamples/s")
