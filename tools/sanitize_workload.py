"""Small workload for compute-sanitizer (tools/sanitize.sh): every FIR kernel size, real and complex masks,
whole-buffer and streaming mode, int16 I/O, the shaped epilogue, the persistent variant (ADT_FIR_PERSIST=2
forces it), and the biquad — a handful of channels each so racecheck finishes in minutes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import pyaudiodsptools_b200 as adt

rng = np.random.default_rng(0)
fams = os.environ.get("ADT_SANITIZE_FAMILIES", "p32").split(",")
for fam in fams:
    os.environ["ADT_FIR_KERNEL"] = fam
    sizes = ((4096, 8192) if fam == "p16" else (16384,) if fam == "c2" else
             tuple(int(v) for v in os.environ.get("ADT_SANITIZE_SIZES", "4096,8192,16384,32768").split(",")))
    for fft in sizes:
        if fft not in adt.design.SUPPORTED_FFT:
            continue
        fs, c, ch = 44100, 1024, 3
        adt.config.initialize(fs, c)
        x = rng.uniform(-1, 1, (ch, 5 * c + 100)).astype(np.float32)
        for kind in ("lowcut", "eq"):
            if kind == "lowcut":
                dev = adt.CreateLowCutFilter(800, channels=ch, fft_size=fft)
                taps = oracle.lowcut_taps(fs, c, 800)
            else:
                dev = adt.CreateEQ3BandFFT(100, 2, 700, -4, 8000, 5, channels=ch, fft_size=fft)
                taps = oracle.eq3_composite_taps(fs, c, 100, 2, 700, -4, 8000, 5)
            y = dev.process(x)
            err = max(float(np.sqrt(np.mean((y[k] - oracle.fir_stream_f64(taps, c, x[k])) ** 2))) for k in range(ch))
            xp = np.pad(x, ((0, 0), (0, 6 * c - x.shape[1])))
            ys = np.concatenate([dev.apply(xp[:, i:i + c]) for i in range(0, 6 * c, c)], axis=1)
            errs = float(np.sqrt(np.mean((ys - y) ** 2)))
            assert err <= 2e-6 and errs <= 1e-6, (fam, fft, kind, err, errs)
            print(f"{fam} N={fft} {kind}: rms {err:.2e} (streaming vs whole {errs:.1e})", flush=True)
        if fam in ("c2",) or fft == 32768:
            continue                                  # cluster transforms: float32 I/O only
        xi = (x * 8000).astype(np.int16)
        dev = adt.CreateLowCutFilter(800, channels=ch, fft_size=fft)
        yi = dev.process_int16(xi)
        assert yi.dtype == np.int16
        if os.environ.get("ADT_SANITIZE_EPILOGUE", "1") == "1":
            dev.set_epilogue(adt.CreateSaturator(-12.0, 1.0, "soft"))
            dev.process(x)
os.environ.pop("ADT_FIR_KERNEL", None)
# a partitioned (segmented) filter: chunk 32768 -> 16383 taps in tap segments, store + accumulate kernels
fs, c = 44100, 32768
adt.config.initialize(fs, c)
dev = adt.CreateLowCutFilter(800, channels=2, fft_size=16384)      # forced N: 3 tap segments, 2 accumulating
xs = rng.uniform(-1, 1, (2, 2 * c + 50)).astype(np.float32)
ys = dev.process(xs)
from scipy.signal import fftconvolve
taps, d = oracle.lowcut_taps(fs, c, 800), oracle.stream_delay(c)
want = np.zeros(ys.shape[1]); want[d:] = fftconvolve(xs[1].astype(np.float64), taps)[: ys.shape[1] - d]
err = float(np.sqrt(np.mean((ys[1] - want) ** 2)))
assert err <= 2e-6, err
print(f"segmented C={c}: {dev.n_segments} segments of N={dev.plan.fft_size}, rms {err:.2e}", flush=True)
eq = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=40)
xb = rng.uniform(-1, 1, (40, 700)).astype(np.float32)
yb = eq.apply(xb)
o = oracle.Eq3BandBiquad(100, 2, 700, -4, 8000, 5)
w = o.applyhighband(o.applymidband(o.applylowband(xb[39].copy())))
assert np.array_equal(yb[39], w)
print("biquad chain bit-exact", flush=True)
print("sanitize workload ok", flush=True)
