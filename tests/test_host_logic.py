"""CPU tier: host-side logic of the product (no GPU, no compute calls into the .so)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
import pyaudiodsptools_b200 as adt
from pyaudiodsptools_b200 import _native, design, devices
from conftest import ROOT, golden_fft_cases, load_golden, rms


def _taps_for(meta):
    fs, c, a = meta["fs"], meta["chunk"], meta["args"]
    if meta["kind"] == "lowcut":
        return design.lowcut_taps(fs, c, a[0] if a else 160)
    if meta["kind"] == "highcut":
        return design.highcut_taps(fs, c, a[0] if a else 8000)
    return design.eq3_taps(fs, c, *a)


def _overlap_save_numpy(plan, x, n_out):
    """What the CUDA engine computes, in float64 numpy, from the plan alone."""
    n, hop, n0, back = plan.fft_size, plan.hop, plan.n0, plan.back
    y = np.zeros(n_out)
    mask = plan.mask.astype(np.complex128)
    for b in range(-(-n_out // hop)):
        ws = b * hop - back
        idx = ws + np.arange(n)
        ok = (idx >= 0) & (idx < len(x))
        w = np.where(ok, x[np.clip(idx, 0, len(x) - 1)], 0.0)
        z = np.fft.ifft(np.fft.fft(w) * mask).real
        m0 = b * hop
        k = min(hop, n_out - m0)
        y[m0:m0 + k] = z[n0:n0 + k]
    return y


@pytest.mark.parametrize("name", golden_fft_cases())
def test_block_plan_reproduces_reference(name):
    """Every golden vector of the reference is reproduced by the block plan(s) replayed in float64 numpy —
    one plan for ordinary filters, several accumulated tap segments for the long ones (EQ at C = 16384,
    Example4.py's chunk 88200), including chunk sizes that give an even filter length (882) or are odd (441)."""
    meta, arr = load_golden(name)
    plans = design.plan_filter(_taps_for(meta), design.stream_delay(meta["chunk"]))
    y = np.zeros(len(arr["y"]))
    for plan in plans:
        assert plan.n0 % 32 == 0 and plan.hop % 32 == 0
        assert plan.back % 32 == 0 or (plan.mask_is_real and meta["chunk"] % 4)   # unaligned windows only for odd chunk sizes
        assert plan.n0 + plan.hop <= plan.fft_size
        y += _overlap_save_numpy(plan, arr["x"].astype(np.float64), len(arr["y"]))
    assert rms(y - arr["y"]) <= 1e-7          # complex64 mask rounding + the reference's own noise
    assert np.max(np.abs(y - arr["y"])) <= 2e-6


@pytest.mark.parametrize("fft_size", [4096, 8192, 16384, 32768])
def test_block_plan_any_fft_size(fft_size):
    meta, arr = load_golden("highcut4000_c1024_noise")
    plan = design.plan_block(_taps_for(meta), design.stream_delay(1024), fft_size)
    assert plan.mask_is_real and plan.fft_size == fft_size
    y = _overlap_save_numpy(plan, arr["x"].astype(np.float64), len(arr["y"]))
    assert rms(y - arr["y"]) <= 1e-7


def test_design_equals_oracle_design():
    for c in (512, 1024, 4096, 16384):
        assert np.array_equal(design.lowcut_taps(44100, c, 800), oracle.lowcut_taps(44100, c, 800))
        assert np.array_equal(design.highcut_taps(96000, c, 4000), oracle.highcut_taps(96000, c, 4000))
        a = design.eq3_taps(44100, c, 100, 2, 700, -4, 8000, 5)
        b = oracle.eq3_composite_taps(44100, c, 100, 2, 700, -4, 8000, 5)
        assert np.max(np.abs(a - b)) <= 1e-15
        assert design.stream_delay(c) == oracle.stream_delay(c) == 3 * c // 4 + 1


def test_long_filters_are_partitioned():
    """A filter that does not fit one transform is split into equal tap segments whose delays tile the taps."""
    with pytest.raises(ValueError):
        design.plan_block(np.ones(40000), design.stream_delay(16384))        # one block plan cannot hold it ...
    taps = np.random.default_rng(0).standard_normal(44099)
    plans = design.plan_filter(taps, 100)                                     # ... the segmented planner can
    assert len(plans) > 1 and sum(p.n_taps for p in plans) == len(taps)
    assert [p.delay for p in plans] == list(np.cumsum([100] + [p.n_taps for p in plans[:-1]]))
    assert len(design.plan_filter(taps[:2047], 3073)) == 1


def test_biquad_coefficients_match_oracle():
    from oracle.biquad import band_coefficients
    for args in ((100, 2, 700, -4, 8000, 5), (250, -6, 1200, 3, 6000, -2), (80, 0, 1000, 0, 10000, 0)):
        for (b0, b1, b2, a0, a1, a2), got in zip(band_coefficients(*args), devices.biquad_coefficients(*args)):
            assert tuple(got) == (b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0)


def test_config_mirror():
    adt.config.initialize(48000, 256)
    assert (adt.config.sampling_rate, adt.config.chunk_size, adt.config.use_gpu) == (48000, 256, False)
    adt.config.initialize(44100, 512, use_gpu=True)
    assert adt.config.use_gpu is True
    adt.config.sampling_rate = None
    with pytest.raises(TypeError):
        adt.CreateLowCutFilter(800)
    adt.config.initialize(44100, 512)


def test_make_and_combine_chunks():
    adt.config.initialize(44100, 512)
    x = np.arange(1300, dtype=np.float32)
    chunks = adt.MakeChunks(x)
    assert len(chunks) == 3 and all(len(c) == 512 for c in chunks)
    back = adt.CombineChunks(chunks)
    assert back.dtype == np.float32 and np.array_equal(back[:1300], x) and not back[1300:].any()
    assert len(adt.MakeChunks(np.zeros(1024, dtype=np.float32))) == 2


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "adt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)      # declarations only, not prose
    declared = set(re.findall(r"\b(adt_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _native.load().adt_version().startswith(b"adt_b200")


def test_no_cpu_fallback_without_device():
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    adt.config.initialize(44100, 512)
    with pytest.raises(adt.AdtError):
        adt.CreateHighCutFilter(4000)
    with pytest.raises(adt.AdtError):
        adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyaudiodsptools_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), fn


def test_header_is_plain_c_and_links(tmp_path):
    """include/adt_b200.h is the drop-in boundary: it must compile as C (no C++ / CUDA / torch types) and
    a C program must link against the shared library; device count works without a GPU."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "adt_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { int n = -1; adt_fir_desc d = {8192, 6144, 1024, 5120, 1, 4096, 1, 0};\n'
                   '  if (adt_device_count(&n) != ADT_OK || d.hop != 6144) return 1;\n'
                   '  printf("%s devices=%d %s\\n", adt_version(), n, adt_status_string(ADT_ERR_UNSUPPORTED)); return 0; }\n')
    exe = tmp_path / "abi"
    lib_dir = os.path.dirname(_native.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", lib_dir, "-l:libadt_b200.so", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.check_output([str(exe)], text=True)
    assert out.startswith("adt_b200") and "unsupported geometry" in out


@pytest.mark.parametrize("seed", range(12))
def test_block_plan_random_filters(seed):
    """plan_block for arbitrary (also non-symmetric, even-length) FIRs and delays: overlap-save replay of the
    plan equals the direct convolution."""
    rng = np.random.default_rng(seed)
    t = int(rng.integers(1, 6000))
    taps = rng.standard_normal(t)
    if seed % 3 == 0 and t % 2 == 1:
        taps = taps + taps[::-1]                       # symmetric -> real mask
    delay = int(rng.integers(0, 5000))
    fft = [4096, 8192, 16384][seed % 3]
    if t + 64 > fft:
        fft = 16384
    plan = design.plan_block(taps, delay, fft)
    assert plan.n0 + plan.hop <= plan.fft_size and plan.hop % 32 == 0 and plan.n0 % 32 == 0
    n = 3 * plan.hop + 1234
    x = rng.uniform(-1, 1, n)
    want = np.zeros(n)
    full = np.convolve(x, taps)
    want[delay:] = full[: n - delay]
    got = _overlap_save_numpy(plan, x, n)
    scale = np.sqrt(np.mean(want ** 2)) + 1e-30
    assert rms(got - want) / scale < 2e-6       # complex64 mask rounding only


def test_wav_helpers_round_trip(tmp_path):
    """The file boundary of Example1/2.py (Utility.py:218-312): int16/32768 on load, int16(x*32767) on store,
    [2, n] transposed to frames; plus the aliases that make the reference's scripts run after an import swap."""
    import pyaudiodsptools_b200 as adt
    adt.config.initialize(22050, 512)
    rng = np.random.default_rng(0)
    mono = rng.uniform(-1, 1, 1000).astype(np.float32)
    p = str(tmp_path / "m.wav")
    adt.NumpyFloatToWav(p, mono)
    back = adt.Utility.MonoWavToNumpyFloat(p)
    assert back.dtype == np.float32
    assert np.array_equal(back, (mono * 32767).astype(np.int16).astype(np.float32) / 32768)
    assert np.array_equal(adt.MonoWavToNumpy16BitInt(p), (mono * 32767).astype(np.int16))
    st = rng.uniform(-1, 1, (2, 700)).astype(np.float32)
    q = str(tmp_path / "s.wav")
    adt.NumpyFloatToWav(q, st)
    left, right = adt.Utility.StereoWavToNumpyFloat(q)
    assert np.array_equal(left, (st[0] * 32767).astype(np.int16).astype(np.float32) / 32768)
    assert np.array_equal(right, (st[1] * 32767).astype(np.int16).astype(np.float32) / 32768)
    with pytest.raises(ValueError):
        adt.Utility.StereoWavToNumpyFloat(p)
    assert np.array_equal(adt.MixSignals(mono, mono), np.clip(mono.astype(np.float64) * 2, -1, 1))
    assert adt.CreateLowCutFilterGPU is adt.CreateLowCutFilter and adt.CreateEQ3BandFFTGPU is adt.CreateEQ3BandFFT
