"""GPU tier: the CUDA path (through the C ABI) against the golden vectors of
the live reference and against the oracle.  Tolerance (north_star): RMS error
<= 1e-5 on float32 chunks; we assert 2e-6 (measured ~2e-7) so regressions show.
"""
import numpy as np
import pytest

import oracle
import pyaudiodsptools_b200 as adt
from conftest import golden_fft_cases, load_golden, rms

pytestmark = pytest.mark.gpu

RMS_TOL = 2e-6      # north_star allows 1e-5
MAX_TOL = 2e-5


def _make(meta, **kw):
    adt.config.initialize(meta["fs"], meta["chunk"])
    ctor = {"lowcut": adt.CreateLowCutFilter, "highcut": adt.CreateHighCutFilter, "eq3fft": adt.CreateEQ3BandFFT}
    return ctor[meta["kind"]](*meta["args"], **kw)


@pytest.mark.parametrize("name", golden_fft_cases())
def test_streaming_apply_matches_reference(gpu_lib, name):
    meta, arr = load_golden(name)
    dev = _make(meta)
    c = meta["chunk"]
    outs = [dev.apply(arr["x"][i:i + c]) for i in range(0, len(arr["x"]), c)]
    assert all(o.dtype == np.float32 and o.shape == (c,) for o in outs)
    y = np.concatenate(outs)
    assert rms(y - arr["y"]) <= RMS_TOL
    assert np.max(np.abs(y - arr["y"])) <= MAX_TOL


@pytest.mark.parametrize("name", golden_fft_cases())
def test_whole_buffer_matches_reference(gpu_lib, name):
    meta, arr = load_golden(name)
    dev = _make(meta)
    y = dev.process(arr["x"])
    assert y.shape == arr["y"].shape
    assert rms(y - arr["y"]) <= RMS_TOL
    assert np.max(np.abs(y - arr["y"])) <= MAX_TOL


@pytest.mark.parametrize("fft_size,kernel", [(4096, "p16"), (8192, "p16"), (4096, "p32"), (8192, "p32"), (16384, "p32"),
                                             (16384, "c2"), (16384, "b2"), (32768, "c4")])
@pytest.mark.parametrize("name", ["highcut4000_c512_noise", "eq3fft_c512_noise", "lowcut160_default_c1024_noise"])
def test_every_kernel_variant(gpu_lib, monkeypatch, name, fft_size, kernel):
    """p16 / p32 = 16 / 32 complex points per thread (fft_core16.cuh / fft_core.cuh), c2 / c4 = one transform on a
    2- / 4-CTA thread-block cluster with the transposition through distributed shared memory (fir_cluster.cuh;
    b2 = the 2-CTA cluster with the bulk-copy-engine exchange through a staging buffer);
    real and complex masks."""
    meta, arr = load_golden(name)
    if kernel == "p16" and b"ab_variants=1" not in gpu_lib.adt_version():
        pytest.skip("the p16 A/B family is only in the AB=1 build (ADT_LIB_PATH=.../libadt_b200_ab.so)")
    monkeypatch.setenv("ADT_FIR_KERNEL", kernel)
    dev = _make(meta, fft_size=fft_size)
    assert dev.plan.fft_size == fft_size
    y = dev.process(arr["x"])
    assert rms(y - arr["y"]) <= RMS_TOL
    c = meta["chunk"]
    ys = np.concatenate([dev.apply(arr["x"][i:i + c]) for i in range(0, len(arr["x"]), c)])
    assert rms(ys - arr["y"]) <= RMS_TOL


@pytest.mark.parametrize("channels", [1, 2, 5, 64])
def test_batched_channels_are_independent(gpu_lib, channels):
    fs, c, n = 44100, 1024, 7 * 1024 + 300     # ragged tail: MakeChunks zero-pads
    adt.config.initialize(fs, c)
    dev = adt.CreateLowCutFilter(800, channels=channels)
    x = np.stack([np.random.default_rng(100 + ch).uniform(-1, 1, n).astype(np.float32) for ch in range(channels)])
    y = dev.process(x)
    assert y.shape == (channels, 8 * c)
    taps = oracle.lowcut_taps(fs, c, 800)
    for ch in range(channels):
        want = oracle.fir_stream_f64(taps, c, x[ch])
        assert rms(y[ch] - want) <= RMS_TOL, ch
    # streaming with a batch: [channels, C] in, [channels, C] out
    xp = np.pad(x, ((0, 0), (0, 8 * c - n)))
    ys = np.concatenate([dev.apply(xp[:, i:i + c]).reshape(channels, c) for i in range(0, 8 * c, c)], axis=1)
    assert rms(ys - y) <= 1e-6


@pytest.mark.parametrize("chunk", [256, 442, 1000, 1001, 2048, 8192, 16384, 32768])
@pytest.mark.parametrize("kind", ["lowcut", "eq3fft"])
def test_other_chunk_sizes(gpu_lib, chunk, kind):
    """Chunk sizes beyond the golden set — not powers of two, not multiples of 4 (even filter length), odd,
    and beyond one transform (partitioned taps): streaming apply and whole-buffer mode against the oracle's
    closed form.  The reference accepts every chunk size (EffectFFTFilter.py:22-25)."""
    fs = 48000
    adt.config.initialize(fs, chunk)
    if kind == "eq3fft":
        dev = adt.CreateEQ3BandFFT(120, 3, 900, -5, 7000, 4, channels=3)
        taps = oracle.eq3_composite_taps(fs, chunk, 120, 3, 900, -5, 7000, 4)
    else:
        dev = adt.CreateLowCutFilter(500, channels=3)
        taps = oracle.lowcut_taps(fs, chunk, 500)
    n = 5 * chunk + chunk // 3
    x = np.random.default_rng(chunk).uniform(-1, 1, (3, n)).astype(np.float32)
    y = dev.process(x)
    assert y.shape == (3, 6 * chunk)
    from scipy.signal import fftconvolve
    d = oracle.stream_delay(chunk)
    for ch in range(3):
        want = np.zeros(y.shape[1])
        want[d:] = fftconvolve(x[ch].astype(np.float64), taps)[: y.shape[1] - d]
        assert rms(y[ch] - want) <= RMS_TOL
    xp = np.pad(x, ((0, 0), (0, 6 * chunk - n)))
    ys = np.concatenate([dev.apply(xp[:, i:i + chunk]) for i in range(0, 6 * chunk, chunk)], axis=1)
    assert rms(ys - y) <= 1e-6


@pytest.mark.parametrize("seed", range(8))
def test_engine_with_arbitrary_filters(gpu_lib, seed):
    """The FIR engine is generic (y[m] = sum_k h[k] x[m - D - k]): random non-symmetric / even-length taps,
    every kernel size, against numpy.convolve."""
    from pyaudiodsptools_b200 import design, devices
    rng = np.random.default_rng(seed)
    chunk = [512, 1024, 4096][seed % 3]
    t = int(rng.integers(2, chunk - 4))
    taps = rng.standard_normal(t) / np.sqrt(t)
    fft = [4096, 8192, 16384, 32768][(seed // 2) % 4]
    dev = devices._FirDevice(taps, chunk, channels=3, fft_size=fft)
    assert dev.plan.mask_is_real == bool(t % 2 == 1 and np.allclose(taps, taps[::-1]))
    n = 5 * chunk + 17
    x = rng.uniform(-1, 1, (3, n)).astype(np.float32)
    y = dev.process(x)
    d = design.stream_delay(chunk)
    for ch in range(3):
        want = np.zeros(y.shape[1])
        full = np.convolve(x[ch].astype(np.float64), taps)
        want[d:] = full[: y.shape[1] - d]
        assert rms(y[ch] - want) <= 3e-6 * max(1.0, rms(want))


def test_eq_batched_stereo_pairs(gpu_lib):
    # BASELINE config 5 shape in miniature: 96 kHz, stereo = two planar rows per stream
    fs, c = 96000, 4096
    adt.config.initialize(fs, c)
    dev = adt.CreateLowCutFilter(800, channels=6)
    x = np.random.default_rng(9).uniform(-1, 1, (6, 5 * c)).astype(np.float32)
    y = dev.process(x)
    taps = oracle.lowcut_taps(fs, c, 800)
    for ch in range(6):
        assert rms(y[ch] - oracle.fir_stream_f64(taps, c, x[ch])) <= RMS_TOL


def test_first_calls_and_latency(gpu_lib):
    # call 0 returns (almost) silence: only the pre-ring of chunk 0 in its last C/4-1 samples
    adt.config.initialize(44100, 512)
    dev = adt.CreateHighCutFilter(4000)
    x = np.zeros(512, dtype=np.float32); x[0] = 1.0
    y0 = dev.apply(x)
    assert np.max(np.abs(y0[: 512 - 127])) < 1e-6 and np.max(np.abs(y0[512 - 127:])) > 1e-4   # FFT rounding noise only
    y1 = dev.apply(np.zeros(512, dtype=np.float32))
    taps = oracle.highcut_taps(44100, 512, 4000)
    full = np.concatenate([y0, y1])
    d = oracle.stream_delay(512)
    assert rms(full[d:d + len(taps)] - taps) <= 1e-7      # impulse response == taps at delay D
    dev.reset()
    assert np.array_equal(dev.apply(x), y0)


def test_wrong_length_raises_and_keeps_state(gpu_lib):
    adt.config.initialize(44100, 512)
    a, b = adt.CreateLowCutFilter(300), adt.CreateLowCutFilter(300)
    x = np.random.default_rng(3).uniform(-1, 1, 3 * 512).astype(np.float32)
    a.apply(x[:512]); b.apply(x[:512])
    with pytest.raises(ValueError):
        a.apply(x[:100])
    assert np.array_equal(a.apply(x[512:1024]), b.apply(x[512:1024]))   # history untouched by the failed call


def test_array_likes_accepted(gpu_lib):
    adt.config.initialize(44100, 512)
    a, b, c = adt.CreateHighCutFilter(4000), adt.CreateHighCutFilter(4000), adt.CreateHighCutFilter(4000)
    x = np.random.default_rng(4).uniform(-1, 1, 512).astype(np.float32)
    ya = a.apply(x)
    assert np.array_equal(ya, b.apply(x.reshape(2, 256)))      # concatenate(axis=None) flattens
    assert np.array_equal(ya, c.apply(list(x)))


def test_linearity_and_shift_invariance_large(gpu_lib):
    # size-independent properties at a BASELINE-sized chunk: 64 channels x 10 s, C = 4096
    fs, c, n = 44100, 4096, 441000
    adt.config.initialize(fs, c)
    dev = adt.CreateLowCutFilter(800, channels=64)
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, (64, n)).astype(np.float32)
    b = rng.uniform(-1, 1, (64, n)).astype(np.float32)
    ya, yb, yab = dev.process(a), dev.process(b), dev.process(a + 0.5 * b)
    assert rms(yab - (ya + 0.5 * yb)) <= 2e-6
    shifted = np.zeros_like(a); shifted[:, 4096:] = a[:, :-4096]
    ys = dev.process(shifted)
    assert rms(ys[:, 4096:] - ya[:, :-4096]) <= 2e-6
    taps = oracle.lowcut_taps(fs, c, 800)
    for ch in (0, 31, 63):
        from scipy.signal import fftconvolve
        full = fftconvolve(a[ch].astype(np.float64), taps)
        want = np.zeros(ya.shape[1]); d = oracle.stream_delay(c)
        want[d:] = full[: ya.shape[1] - d]
        assert rms(ya[ch] - want) <= RMS_TOL


_PERSIST_WORKER = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np
import pyaudiodsptools_b200 as adt
import oracle
from scipy.signal import fftconvolve
fs, c, rows, n = 44100, 4096, 240, 441000
adt.config.initialize(fs, c)
dev = adt.CreateLowCutFilter(800, channels=rows, fft_size=16384)   # the size whose persistent kernel is in the default build
x = np.random.default_rng(1).uniform(-1, 1, (rows, n)).astype(np.float32)
y = dev.process(x)                      # 3+ row groups on concurrent copy streams, each a persistent launch
taps, d = oracle.lowcut_taps(fs, c, 800), oracle.stream_delay(c)
for ch in (0, 75, 76, 151, 152, rows - 1):
    want = np.zeros(y.shape[1]); full = fftconvolve(x[ch].astype(np.float64), taps); want[d:] = full[: y.shape[1] - d]
    err = float(np.sqrt(np.mean((y[ch] - want) ** 2)))
    assert err <= 2e-6, (ch, err)
print("persistent host pipeline ok")
"""


def test_persistent_variant_on_concurrent_streams(gpu_lib, tmp_path):
    """The dynamic-queue persistent kernel keeps one work counter per stream slot: row groups of
    process() run on three copy streams at once (regression test for a shared-counter race)."""
    import os, subprocess, sys
    from conftest import ROOT
    script = tmp_path / "p.py"
    script.write_text(_PERSIST_WORKER.format(root=ROOT))
    env = dict(os.environ, ADT_FIR_PERSIST="1", ADT_FIR_GROUP_MB="128")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "persistent host pipeline ok" in r.stdout


def test_device_resident_whole_buffer(gpu_lib):
    fs, c, rows, n = 44100, 4096, 9, 5 * 4096
    adt.config.initialize(fs, c)
    dev = adt.CreateEQ3BandFFT(100, 2, 700, -4, 8000, 5, channels=rows)
    ctx = dev.context
    x = np.random.default_rng(5).uniform(-1, 1, (rows, n)).astype(np.float32)
    dx, dy = ctx.malloc(x.nbytes), ctx.malloc(x.nbytes)
    ctx.h2d(dx, x)
    before = ctx.launch_count()
    dev.process_device(dx, n, n, dy, n, n, rows)
    ctx.sync()
    assert ctx.launch_count() == before + 1
    y = np.empty_like(x)
    ctx.d2h(y, dy)
    ctx.free(dx); ctx.free(dy)
    taps = oracle.eq3_composite_taps(fs, c, 100, 2, 700, -4, 8000, 5)
    for ch in range(rows):
        assert rms(y[ch] - oracle.fir_stream_f64(taps, c, x[ch])) <= RMS_TOL


def test_out_buffer_is_validated(gpu_lib):
    """A wrong-sized / wrong-typed out= never reaches the C library (explicit checks, not asserts)."""
    adt.config.initialize(44100, 512)
    dev = adt.CreateLowCutFilter(300, channels=2)
    x = np.zeros((2, 1024), dtype=np.float32)
    with pytest.raises(ValueError):
        dev.process(x, out=np.empty((2, 512), dtype=np.float32))
    with pytest.raises(TypeError):
        dev.process(x, out=np.empty((2, 1024), dtype=np.float64))
    with pytest.raises(ValueError):
        dev.apply(x[:, :512], out=np.empty(100, dtype=np.float32))
    with pytest.raises(ValueError):
        dev.process(x, out=np.empty((2, 2048), dtype=np.float32)[:, ::2])
    dev.process(x, out=np.empty((2, 1024), dtype=np.float32))


def test_capi_rejects_bad_geometry(gpu_lib):
    import ctypes as C
    from pyaudiodsptools_b200 import _native
    ctx = _native.default_context(0)
    mask = np.zeros(2 * 8192, dtype=np.float32)
    h = C.c_void_p()
    for desc in (_native.FirDesc(1000, 10, 0, 0, 0, 0, 0, 0),       # unsupported N
                 _native.FirDesc(8192, 8000, 500, 0, 0, 0, 0, 0),   # n0 + hop > N
                 _native.FirDesc(8192, 64, 0, 0, 0, 512, 0, 0)):    # chunk without channels
        rc = gpu_lib.adt_fir_create(ctx.h, C.byref(desc), mask.ctypes.data, C.byref(h))
        assert rc in (-1, -4) and not h.value
        assert gpu_lib.adt_last_error(ctx.h)


# ---- 16-bit PCM mode ----------------------------------------------------
@pytest.mark.parametrize("name", ["lowcut800_c4096_wav_int16", "lowcut800_c4096_stereo_int16"])
def test_int16_mode_matches_reference_chain(gpu_lib, name):
    """int16 in -> /32768 -> filter -> *32767 -> truncate, all fused.  The float path differs from the
    reference by ~3e-7, so after truncation a sample may land one LSB away; never more."""
    meta, arr = load_golden(name)
    adt.config.initialize(meta["fs"], meta["chunk"])
    rows = 1 if arr["x"].ndim == 1 else arr["x"].shape[0]
    dev = adt.CreateLowCutFilter(*meta["args"], channels=rows)
    y = dev.process_int16(arr["x"])
    assert y.dtype == np.int16 and y.shape == arr["y"].shape
    diff = np.abs(y.astype(np.int32) - arr["y"].astype(np.int32))
    assert diff.max() <= 1
    assert np.mean(diff != 0) < 0.05
    # and it is exactly the float path with the two conversions applied around it
    yf = dev.process(arr["x"].astype(np.float32) / 32768)
    assert np.array_equal(y, (yf * 32767).astype(np.int16))


# ---- in-repo consumers of the path -------------------------------------------
@pytest.mark.parametrize("name", ["saturator_default", "saturator_soft", "softclipper_default", "softclipper_drive2"])
def test_shapers_match_reference(gpu_lib, name):
    meta, arr = load_golden(name)
    dev = (adt.CreateSaturator if meta["kind"] == "saturator" else adt.CreateSoftClipper)(**meta["kwargs"])
    y = dev.apply(arr["x"])
    assert y.dtype == np.float32
    if meta["kind"] == "saturator":
        np.testing.assert_array_equal(y, arr["y"])                  # same float32 operations in the same order
    else:
        assert np.max(np.abs(y - arr["y"])) <= 4e-7                 # powf: last-ulp differences only


@pytest.mark.parametrize("name", ["delay_default", "delay_wet3"])
def test_delay_bit_exact(gpu_lib, name):
    meta, arr = load_golden(name)
    adt.config.initialize(meta["fs"], meta["chunk"])
    dev = adt.CreateDelay(**meta["kwargs"])
    c = meta["chunk"]
    y = np.concatenate([dev.apply(arr["x"][i:i + c]) for i in range(0, len(arr["x"]), c)])
    np.testing.assert_array_equal(y, arr["y"])


def test_delay_batched_with_prefilters(gpu_lib):
    # pre-filter flags raise AttributeError in the reference (EffectDelay.py:56,58); here they compose
    fs, c, ch = 44100, 512, 3
    adt.config.initialize(fs, c)
    dev = adt.CreateDelay(time_in_ms=30, feedback_loops=2, lowcut_filter_frequency=300, use_lowcut_filter=True,
                          channels=ch)
    lc = adt.CreateLowCutFilter(300, channels=ch)
    x = np.random.default_rng(2).uniform(-1, 1, (ch, 12 * c)).astype(np.float32)
    got = np.concatenate([dev.apply(x[:, i:i + c]) for i in range(0, 12 * c, c)], axis=1)
    filt_all = lc.process(x)                                      # same filter, whole-buffer mode
    for k in range(ch):
        o = oracle.FeedbackDelay(fs, time_in_ms=30, feedback_loops=2)
        want = np.concatenate([o.apply(filt_all[k, i:i + c]) for i in range(0, 12 * c, c)])
        assert rms(got[k] - want) <= 2e-6


def test_fused_epilogue_equals_separate_pass(gpu_lib):
    fs, c = 44100, 4096
    adt.config.initialize(fs, c)
    x = np.random.default_rng(6).uniform(-1, 1, (5, 4 * c + 77)).astype(np.float32)
    for shaper in (adt.CreateSaturator(-12.0, 1.0, 'soft'), adt.CreateSoftClipper(0.8)):
        plain = adt.CreateHighCutFilter(6000, channels=5)
        fused = adt.CreateHighCutFilter(6000, channels=5, epilogue=shaper)
        y = plain.process(x)
        # fused = same filter kernel + the shaper with fast division / __powf in the store phase
        assert np.max(np.abs(fused.process(x) - shaper.apply(y))) <= 3e-6
        ys = np.concatenate([fused.apply(np.pad(x, ((0, 0), (0, 5 * c - x.shape[1])))[:, i:i + c])
                             for i in range(0, 5 * c, c)], axis=1)
        assert rms(ys - shaper.apply(y)) <= 1e-6
        fused.set_epilogue(None)
        assert np.array_equal(fused.process(x), y)


# ---- the streaming biquad: bit-exact ---------------------------------------------
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_biquad_bit_exact(gpu_lib, tag):
    meta, arr = load_golden("eq3biquad_" + tag)
    blk, x = meta["block"], arr["x"]
    eq, eq2 = adt.CreateEQ3Band(*meta["args"]), adt.CreateEQ3Band(*meta["args"])
    outs = {"low": [], "mid": [], "high": [], "chain": []}
    for i in range(0, len(x), blk):
        b = x[i:i + blk]
        outs["low"].append(eq.applylowband(b)); outs["mid"].append(eq.applymidband(b))
        outs["high"].append(eq.applyhighband(b)); outs["chain"].append(eq2.apply(b))
    for k, v in outs.items():
        got = np.concatenate(v)
        assert got.dtype == arr[k].dtype
        np.testing.assert_array_equal(got, arr[k])


def test_biquad_batched_channels(gpu_lib):
    chans, n = 70, 1000     # not a multiple of 32 channels, ragged time tile
    eq = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=chans)
    x = np.random.default_rng(8).uniform(-1, 1, (chans, 2 * n)).astype(np.float32)
    y = np.concatenate([eq.applymidband(x[:, :n]), eq.applymidband(x[:, n:])], axis=1)
    for ch in (0, 31, 32, 69):
        o = oracle.Eq3BandBiquad(100, 2, 700, -4, 8000, 5)
        want = np.concatenate([o.applymidband(x[ch, :n].copy()), o.applymidband(x[ch, n:].copy())])
        np.testing.assert_array_equal(y[ch], want)


@pytest.mark.parametrize("pipe,round_int", [("1", "0"), ("0", "0"), ("0", "1")])
def test_biquad_fused_chain_equals_three_launches(gpu_lib, pipe, round_int, tmp_path):
    """The one-launch low->mid->high pipeline — helper + chain warp per band (pipe 1, the default) or one warp per
    band (pipe 0) — is bit-identical to the three per-band launches (and to the
    oracle), shares their state, handles ragged tiles / channel counts, both rounding implementations, and
    float32 denormals (a decaying tail crosses 2^-126, where the integer rounding hands over to F2F)."""
    import os, subprocess, sys
    from conftest import ROOT
    script = tmp_path / "b.py"
    script.write_text(f"""
import sys
sys.path.insert(0, {ROOT!r})
import numpy as np
import oracle
import pyaudiodsptools_b200 as adt
chans, n = 70, 1517
rng = np.random.default_rng(21)
x = rng.uniform(-1, 1, (chans, 3 * n)).astype(np.float32)
x[:, n:] *= 1e-36                      # outputs decay through the float32 denormal range
x[:, 2 * n:] = 0
a = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=chans)
b = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=chans)
for i in range(3):
    blk = x[:, i * n:(i + 1) * n]
    ya = a.apply(blk)
    yb = b.applyhighband(b.applymidband(b.applylowband(blk)))
    assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32)), i
o = oracle.Eq3BandBiquad(100, 2, 700, -4, 8000, 5)
c = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=chans)
y = np.concatenate([c.apply(x[:, :n]), c.applymidband(x[:, n:2 * n])], axis=1)     # mixed styles share state
want = np.concatenate([o.applyhighband(o.applymidband(o.applylowband(x[69, :n].copy()))), o.applymidband(x[69, n:2 * n].copy())])
assert np.array_equal(y[69], want)
xd = rng.uniform(-1, 1, (3, 500))
d = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=3)
od = oracle.Eq3BandBiquad(100, 2, 700, -4, 8000, 5)
assert np.array_equal(d.apply(xd)[2], od.applyhighband(od.applymidband(od.applylowband(xd[2].copy()))))
print("fused chain ok")
""")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, ADT_BIQUAD_ROUND_INT=round_int, ADT_BIQUAD_PIPE=pipe))
    assert r.returncode == 0 and "fused chain ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
