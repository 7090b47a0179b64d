import sys, time
sys.path.insert(0, '.')
import numpy as np, bench
import pyaudiodsptools_b200 as adt
from pyaudiodsptools_b200 import _native
ctx = _native.default_context(0)
eq = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=1000)
n = 441000
base = np.random.default_rng(78).uniform(-1, 1, (16, n)).astype(np.float32)
dx, dy = ctx.malloc(1000 * n * 4), ctx.malloc(1000 * n * 4)
bench._fill_device_rows(ctx, dx, base, 1000, n * 4)
eq.apply_device(dx, dy, n, n); ctx.sync()
s = bench.ClockSampler(0); s.start()
e0, e1 = ctx.event(), ctx.event()
e0.record()
for _ in range(20):
    eq.apply_device(dx, dy, n, n)
e1.record()
ms = e0.elapsed_ms(e1) / 20
c = s.stop()
print("ms per chain", round(ms, 2), "clocks", c)
