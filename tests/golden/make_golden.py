#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (pyAudioDspTools, imported from /root/reference) ships no golden
vectors or assertions of its own, so these fixtures — outputs of
the reference's own classes on seeded inputs and on its shipped WAV — are what
pins the oracle and the CUDA path.  Each .npz holds: ``meta`` (json), ``x``
(float32 input stream), ``y`` (concatenated ``apply`` outputs, float32).
"""
import io
import json
import os
import sys
import contextlib

import numpy as np

REF = os.environ.get("ADT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):  # "No cupy" info line
        import pyAudioDspTools as ref
    return ref


def _noise(seed, n):
    return np.random.default_rng(seed).uniform(-1, 1, n).astype("float32")


def _run_chunks(dev, x, chunk):
    outs = [dev.apply(x[i:i + chunk]) for i in range(0, len(x), chunk)]
    return np.concatenate(outs).astype("float32")


def _save(name, meta, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name}: " + ", ".join(f"{k}{v.shape}" for k, v in arrays.items()))


def main():
    ref = _import_reference()
    numpy_version = np.__version__

    only = [a for a in sys.argv[1:] if not a.startswith("-")]     # optional name prefixes: regenerate a subset

    def fft_case(name, fs, chunk, kind, args, x):
        if only and not any(name.startswith(o) for o in only):
            return
        ref.config.initialize(fs, chunk)
        ctor = {"lowcut": ref.CreateLowCutFilter, "highcut": ref.CreateHighCutFilter,
                "eq3fft": ref.CreateEQ3BandFFT}[kind]
        dev = ctor(*args)
        y = _run_chunks(dev, x, chunk)
        meta = dict(kind=kind, fs=fs, chunk=chunk, args=list(args), numpy=numpy_version,
                    source="pyAudioDspTools @ /root/reference, EffectFFTFilter.py / EffectEQ3BandFFT.py .apply")
        _save(name, meta, x=x, y=y)

    # --- BASELINE configs 1/2: Example1.py parameters, C=4096 -------------------
    fft_case("lowcut800_c4096_noise", 44100, 4096, "lowcut", (800,), _noise(1234, 6 * 4096))
    # Example1.py path on the shipped WAV (first 8 chunks), Utility.py:218-238 loader
    ref.config.initialize(44100, 4096)
    wav = ref.MonoWavToNumpyFloat(os.path.join(REF, "TestFile16BitMono.wav"))
    fft_case("lowcut800_c4096_wav", 44100, 4096, "lowcut", (800,), wav[20000:20000 + 8 * 4096].astype("float32"))
    # --- ModuleTests.py parameter sets at C=512 on a 1 kHz sine (ModuleTests.py:34,57,81-84)
    ref.config.initialize(44100, 512)
    sine = ref.CreateSinewave(1000, 44100 // 4).astype("float32")
    sine = sine[: (len(sine) // 512) * 512]
    fft_case("lowcut200_c512_sine", 44100, 512, "lowcut", (200,), sine)
    fft_case("highcut8000_c512_sine", 44100, 512, "highcut", (8000,), sine)
    fft_case("eq3fft_c512_sine", 44100, 512, "eq3fft", (100, 2, 700, -4, 8000, 5), sine)
    # --- config 3: EQ at C=4096 and C=512 on noise -----------------------------------
    fft_case("eq3fft_c4096_noise", 44100, 4096, "eq3fft", (100, 2, 700, -4, 8000, 5), _noise(77, 6 * 4096))
    fft_case("eq3fft_c512_noise", 44100, 512, "eq3fft", (100, 2, 700, -4, 8000, 5), _noise(78, 12 * 512))
    # --- config 4: high-cut 4 kHz chunk sweep ----------------------------------------
    for c, nch in ((512, 12), (1024, 8), (4096, 5), (16384, 4)):
        fft_case(f"highcut4000_c{c}_noise", 44100, c, "highcut", (4000,), _noise(4000 + c, nch * c))
    # --- round 2: chunk sizes beyond one transform and chunk sizes that are not multiples of 4 ---------
    # EQ at C = 16384 (16381-tap composite), Example4.py's chunk 88200 (44099 taps -> partitioned filter),
    # C = 882 / 441 (even / odd filter length: the reference runs them, EffectFFTFilter.py:22-25)
    fft_case("eq3fft_c16384_noise", 44100, 16384, "eq3fft", (100, 2, 700, -4, 8000, 5), _noise(16384, 4 * 16384))
    fft_case("lowcut800_c88200_noise", 44100, 88200, "lowcut", (800,), _noise(88200, 3 * 88200))
    fft_case("lowcut300_c882_noise", 44100, 882, "lowcut", (300,), _noise(882, 7 * 882))
    fft_case("eq3fft_c882_noise", 44100, 882, "eq3fft", (100, 2, 700, -4, 8000, 5), _noise(883, 7 * 882))
    fft_case("highcut4000_c441_noise", 44100, 441, "highcut", (4000,), _noise(441, 9 * 441))
    # --- config 5: 96 kHz low-cut ----------------------------------------------------
    fft_case("lowcut800_c4096_96k_noise", 96000, 4096, "lowcut", (800,), _noise(96, 5 * 4096))
    # --- defaults and edge inputs ----------------------------------------------------
    fft_case("lowcut160_default_c1024_noise", 44100, 1024, "lowcut", (), _noise(5, 6 * 1024))
    fft_case("highcut8000_default_c2048_noise", 44100, 2048, "highcut", (), _noise(6, 5 * 2048))
    imp = np.zeros(5 * 512, dtype="float32"); imp[0] = 1; imp[511] = -1; imp[512] = 0.5; imp[3 * 512 - 1] = 1
    fft_case("lowcut800_c512_impulses", 44100, 512, "lowcut", (800,), imp)
    sq = np.where((np.arange(6 * 512) // 37) % 2 == 0, 1.0, -1.0).astype("float32")
    fft_case("eq3fft_c512_square", 44100, 512, "eq3fft", (250, -6, 1200, 3, 6000, -2), sq)
    fft_case("highcut4000_c512_zeros", 44100, 512, "highcut", (4000,), np.zeros(4 * 512, dtype="float32"))

    if only:
        return
    # --- the streaming biquad (EffectEQ3Band.py:90-180) -------------------------------
    for tag, dtype in (("f32", "float32"), ("f64", "float64")):
        eq = ref.CreateEQ3Band(100, 2, 700, -4, 8000, 5)
        x = np.random.default_rng(321).uniform(-1, 1, 3 * 1024).astype(dtype)
        outs = {"low": [], "mid": [], "high": [], "chain": []}
        eq2 = ref.CreateEQ3Band(100, 2, 700, -4, 8000, 5)
        for i in range(0, len(x), 1024):
            blk = x[i:i + 1024]
            outs["low"].append(eq.applylowband(blk.copy()))
            outs["mid"].append(eq.applymidband(blk.copy()))
            outs["high"].append(eq.applyhighband(blk.copy()))
            outs["chain"].append(eq2.applyhighband(eq2.applymidband(eq2.applylowband(blk.copy()))))
        meta = dict(kind="eq3biquad", args=[100, 2, 700, -4, 8000, 5], block=1024, dtype=dtype,
                    numpy=numpy_version, source="EffectEQ3Band.py:90-180")
        _save(f"eq3biquad_{tag}", meta, x=x, **{k: np.concatenate(v) for k, v in outs.items()})

    # --- 16-bit PCM end to end: the Example1.py / Example2.py chains with their WAV
    #     conversions, Utility.py:218-238 (int16 -> float32/32768) and :306 ((y*32767).astype(int16))
    import wave
    ref.config.initialize(44100, 4096)
    with wave.open(os.path.join(REF, "TestFile16BitMono.wav")) as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)[20000:20000 + 8 * 4096].copy()
    dev = ref.CreateLowCutFilter(800)
    yf = _run_chunks(dev, pcm.astype("float32") / 32768, 4096)
    _save("lowcut800_c4096_wav_int16", dict(kind="lowcut", fs=44100, chunk=4096, args=[800], numpy=numpy_version,
                                             source="Example1.py chain incl. Utility.py:236-237 and :306"),
          x=pcm, y=(yf * 32767).astype("int16"))
    with wave.open(os.path.join(REF, "TestFile16BitStereo.wav")) as w:
        st = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).reshape(-1, 2)[30000:30000 + 6 * 4096].T.copy()
    outs = []
    for row in st:                                                                             # Example2.py:13-22
        d = ref.CreateLowCutFilter(800)
        outs.append((_run_chunks(d, row.astype("float32") / 32768, 4096) * 32767).astype("int16"))
    _save("lowcut800_c4096_stereo_int16", dict(kind="lowcut", fs=44100, chunk=4096, args=[800], numpy=numpy_version,
                                                source="Example2.py chain, planar L/R rows"),
          x=st, y=np.stack(outs))

    # --- the in-repo consumers of the path: wave-shapers and the delay ------------
    xs = (np.random.default_rng(55).uniform(-1.3, 1.3, 4096)).astype("float32")    # incl. |x| > 1
    for tag, kw in (("default", {}), ("soft", dict(saturation_threshold_in_db=-12.0, makeup_gain=0.0, mode="soft"))):
        dev = ref.CreateSaturator(**kw)
        _save(f"saturator_{tag}", dict(kind="saturator", kwargs=kw, numpy=numpy_version,
                                       source="EffectSaturator.py:18-49"), x=xs, y=dev.apply(xs.copy()))
    for tag, kw in (("default", {}), ("drive2", dict(drive=2.0))):
        dev = ref.CreateSoftClipper(**kw)
        _save(f"softclipper_{tag}", dict(kind="softclipper", kwargs=kw, numpy=numpy_version,
                                         source="EffectSoftClipper.py:19-44"), x=xs, y=dev.apply(xs.copy()))
    ref.config.initialize(44100, 512)
    xd = _noise(91, 40 * 512)
    for tag, kw in (("default", dict(time_in_ms=50)), ("wet3", dict(time_in_ms=20, feedback_loops=3, wet=True))):
        dev = ref.CreateDelay(**kw)
        y = np.concatenate([np.array(dev.apply(xd[i:i + 512].copy())) for i in range(0, len(xd), 512)])
        _save(f"delay_{tag}", dict(kind="delay", fs=44100, chunk=512, kwargs=kw, numpy=numpy_version,
                                   source="EffectDelay.py:31-74"), x=xd, y=y.astype("float32"))

    # --- tap designs (float64) so the host-side design code is pinned too --------------
    ref.config.initialize(44100, 4096)
    d = ref.CreateLowCutFilter(800)
    e = ref.CreateEQ3BandFFT(100, 2, 700, -4, 8000, 5)
    _save("masks_c4096", dict(fs=44100, chunk=4096, numpy=numpy_version,
                              note="sinc_filter attributes after __init__ (complex128 [3C])"),
          lowcut800=d.sinc_filter, eq_hs=e.sinc_filter_highshelf, eq_ls=e.sinc_filter_lowshelf,
          eq_mlp=e.sinc_filter_mid_lowpass, eq_mhp=e.sinc_filter_mid_highpass)


if __name__ == "__main__":
    main()
