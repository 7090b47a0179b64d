"""The oracle (oracle/) against golden vectors produced by the live reference.

CPU only.  Pins both forms of the oracle: the windowed restatement of
EffectFFTFilter.py:125-151 / EffectEQ3BandFFT.py:156-211 and the closed-form
float64 FIR, plus the biquad restatement.
"""
import numpy as np
import pytest

import oracle
from conftest import golden_fft_cases, load_golden, rms


def _windowed(meta):
    fs, c, a = meta["fs"], meta["chunk"], meta["args"]
    if meta["kind"] == "lowcut":
        return oracle.SlidingFftFilter(fs, c, a[0] if a else 160, "lowcut")
    if meta["kind"] == "highcut":
        return oracle.SlidingFftFilter(fs, c, a[0] if a else 8000, "highcut")
    return oracle.SlidingFftEq3(fs, c, *a)


def _composite(meta):
    fs, c, a = meta["fs"], meta["chunk"], meta["args"]
    if meta["kind"] == "lowcut":
        return oracle.lowcut_taps(fs, c, a[0] if a else 160)
    if meta["kind"] == "highcut":
        return oracle.highcut_taps(fs, c, a[0] if a else 8000)
    return oracle.eq3_composite_taps(fs, c, *a)


@pytest.mark.parametrize("name", golden_fft_cases())
def test_windowed_oracle_matches_reference(name):
    meta, arr = load_golden(name)
    dev = _windowed(meta)
    c = meta["chunk"]
    y = np.concatenate([dev.apply(arr["x"][i:i + c]) for i in range(0, len(arr["x"]), c)])
    # same numpy calls in the same order: identical up to last-bit effects
    assert rms(y - arr["y"]) <= 2e-8
    assert np.max(np.abs(y - arr["y"])) <= 5e-7


@pytest.mark.parametrize("name", golden_fft_cases())
def test_closed_form_matches_reference(name):
    meta, arr = load_golden(name)
    y = oracle.fir_stream_f64(_composite(meta), meta["chunk"], arr["x"])
    err = rms(y - arr["y"])
    # the reference's own float32/complex64 rounding noise is ~2e-8 RMS
    assert err <= 1e-7, err
    assert np.max(np.abs(y - arr["y"])) <= 1e-6


@pytest.mark.parametrize("name", ["lowcut800_c4096_wav_int16", "lowcut800_c4096_stereo_int16"])
def test_int16_chain_oracle(name):
    meta, arr = load_golden(name)
    x = np.atleast_2d(arr["x"]); want = np.atleast_2d(arr["y"])
    for row, w in zip(x, want):
        dev = oracle.SlidingFftFilter(meta["fs"], meta["chunk"], meta["args"][0], "lowcut")
        c = meta["chunk"]
        xf = row.astype("float32") / 32768                               # Utility.py:236-237
        y = np.concatenate([dev.apply(xf[i:i + c]) for i in range(0, len(xf), c)])
        got = (y * 32767).astype("int16")                                # Utility.py:306
        assert np.max(np.abs(got.astype(int) - w.astype(int))) <= 1 and np.mean(got != w) < 0.01


@pytest.mark.parametrize("name", ["saturator_default", "saturator_soft", "softclipper_default", "softclipper_drive2"])
def test_shaper_oracle_bit_exact(name):
    meta, arr = load_golden(name)
    fn = oracle.saturator if meta["kind"] == "saturator" else oracle.soft_clipper
    kw = dict(meta["kwargs"])
    if "saturation_threshold_in_db" in kw:
        kw["threshold_db"] = kw.pop("saturation_threshold_in_db")
    y = fn(arr["x"], **kw)
    assert y.dtype == arr["y"].dtype
    np.testing.assert_array_equal(y, arr["y"])


@pytest.mark.parametrize("name", ["delay_default", "delay_wet3"])
def test_delay_oracle_bit_exact(name):
    meta, arr = load_golden(name)
    dev = oracle.FeedbackDelay(meta["fs"], **meta["kwargs"])
    c = meta["chunk"]
    y = np.concatenate([dev.apply(arr["x"][i:i + c]) for i in range(0, len(arr["x"]), c)])
    np.testing.assert_array_equal(y, arr["y"])


def test_mask_design_matches_reference():
    meta, arr = load_golden("masks_c4096")
    from oracle.fftfilter import _padded_mask
    fs, c = meta["fs"], meta["chunk"]
    np.testing.assert_allclose(_padded_mask(oracle.lowcut_taps(fs, c, 800), c), arr["lowcut800"], rtol=0, atol=1e-13)
    hs, ls, mlp, mhp = oracle.eq3_band_taps(fs, c, 100, 700, 8000)
    for h, key in ((hs, "eq_hs"), (ls, "eq_ls"), (mlp, "eq_mlp"), (mhp, "eq_mhp")):
        np.testing.assert_allclose(_padded_mask(h, c), arr[key], rtol=0, atol=1e-13)


def test_known_answers_lowcut_800():
    # |H| ~ 0 at 100/400 Hz, ~0.506 at 800 Hz, 1.000 at >= 1.6 kHz
    fs, c = 44100, 4096
    h = oracle.lowcut_taps(fs, c, 800)
    H = lambda f: abs(np.sum(h * np.exp(-2j * np.pi * f / fs * np.arange(len(h)))))
    assert H(100) < 1e-3 and H(400) < 1e-3
    assert abs(H(800) - 0.5) < 1e-2  # -6 dB point of a windowed sinc (SURVEY quotes 0.506 at the nearest FFT bin)
    assert abs(H(1600) - 1.0) < 1e-3 and abs(H(10000) - 1.0) < 1e-3


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_biquad_oracle_matches_reference(tag):
    meta, arr = load_golden("eq3biquad_" + tag)
    blk = meta["block"]
    x = arr["x"]
    eq, eq2 = oracle.Eq3BandBiquad(*meta["args"]), oracle.Eq3BandBiquad(*meta["args"])
    outs = {"low": [], "mid": [], "high": [], "chain": []}
    for i in range(0, len(x), blk):
        b = x[i:i + blk]
        outs["low"].append(eq.applylowband(b.copy()))
        outs["mid"].append(eq.applymidband(b.copy()))
        outs["high"].append(eq.applyhighband(b.copy()))
        outs["chain"].append(eq2.applyhighband(eq2.applymidband(eq2.applylowband(b.copy()))))
    for k, v in outs.items():
        got = np.concatenate(v)
        assert got.dtype == arr[k].dtype
        np.testing.assert_array_equal(got, arr[k])  # bit-exact: same arithmetic, same order
