"""Host-side filter design and block planning (init time, not hot).

Tap arithmetic follows the reference constructors in float64 numpy so that
sinc / Blackman / Kaiser(i0) are not part of the parity surface:
  EffectFFTFilter.py:20-37 (high cut), :93-113 (low cut = spectral inversion),
  EffectEQ3BandFFT.py:65-143 (four Kaiser(6.0) designs).
The EQ's three masked inverse transforms collapse into ONE composite FIR
(SURVEY.md §8(a) A5); `plan_block` then turns any composite FIR into the
overlap-save geometry + spectral mask the CUDA engine consumes.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

SUPPORTED_FFT = (4096, 8192, 16384, 32768)

def filter_length(chunk: int) -> int:
    return chunk // 2 - 1  # EffectFFTFilter.py:22


def stream_delay(chunk: int) -> int:
    """D with out[m] = sum_k h[k] x[m - D - k]: the reference slices the
    3-chunk window at C + L//2 (EffectFFTFilter.py:24,97), so the sample that
    meets h[0] is 2C - (C + L//2) = C - L//2 behind the output index — for the
    usual odd L that is zero-phase on the previous chunk, D = C - (L-1)/2; the
    same formula covers chunk sizes that give an even L (e.g. 882, 441)."""
    return chunk - filter_length(chunk) // 2


def _lowpass(cut_hz, fs, n, window):
    k = np.arange(n) - (n - 1) / 2
    h = np.sinc(2 * cut_hz / fs * k) * window
    return h / h.sum()


def _spectral_inversion(h):
    g = -h
    g[(len(h) - 1) // 2] += 1
    return g


def highcut_taps(fs, chunk, cutoff_hz):
    n = filter_length(chunk)
    return _lowpass(cutoff_hz, fs, n, np.blackman(n))


def lowcut_taps(fs, chunk, cutoff_hz):
    return _spectral_inversion(highcut_taps(fs, chunk, cutoff_hz))


def eq3_taps(fs, chunk, f_low, db_low, f_mid, db_mid, f_high, db_high):
    """Composite FIR of CreateEQ3BandFFT: dry tap + (g-1)-weighted shelves +
    (g_mid-1) * (mid high-pass (*) mid low-pass).  The mid product is 2L-1 taps
    long and, sliced like the L-tap bands, arrives C/4-1 samples late — kept."""
    n = filter_length(chunk)
    win = np.kaiser(n, 6.0)
    hs = _spectral_inversion(_lowpass(0.75 * f_high, fs, n, win))   # :72-83
    ls = _lowpass(1.25 * f_low, fs, n, win)                          # :95-102
    mid = np.convolve(_spectral_inversion(_lowpass(0.75 * f_mid, fs, n, win)),  # :122-133
                      _lowpass(1.25 * f_mid, fs, n, win))                       # :112-119
    g_hs, g_ls, g_mid = (10.0 ** (db / 20.0) for db in (db_high, db_low, db_mid))  # :195,:200,:205
    h = (g_mid - 1.0) * mid
    h[:n] += (g_hs - 1.0) * hs + (g_ls - 1.0) * ls
    h[n // 2] += 1.0  # the dry middle chunk (:209) sits exactly C behind: tap index C - D = L//2
    return h


@dataclasses.dataclass
class BlockPlan:
    fft_size: int
    hop: int
    n0: int
    back: int
    mask_is_real: bool
    mask: np.ndarray      # complex64 [fft_size], natural bin order, no 1/N
    n_taps: int
    delay: int


# Measured cost of one FFT point (kernel time per transform point, relative to N = 8192) on B200
# (gpurun_out/q7, r2d): the 4-CTA/SM N = 4096 kernels overlap memory and FP phases best, the 1-CTA/SM
# N = 16384 kernel is worse, and the 4-CTA-cluster N = 32768 transform is bound by the distributed-shared-
# memory exchange (~11 B/clk/SM of st.async) — it is chosen only where it saves whole
# passes (e.g. the 16381-tap EQ at chunk 16384).  The planner minimises  segments * cost * N / hop.
_POINT_COST = {4096: 0.82, 8192: 1.00, 16384: 1.36, 32768: 2.30}
_ALIGN_SLACK = 64          # worst-case loss of hop to the 32-sample alignment of n0 and hop (plan_block)
_MAX_SEGMENTS = 64


def _plan_sizes(n_taps: int, fft_size: int | None = None):
    """(fft_size, n_segments, taps_per_segment) with the lowest estimated cost per output sample.

    One segment whenever the filter fits a transform with a useful hop; otherwise the taps are partitioned
    in time into equal segments (each its own overlap-save pass, the later ones accumulating — DESIGN.md
    §3.1), which is what makes every chunk size the reference accepts work here too."""
    forced = fft_size or (int(os.environ["ADT_FFT_SIZE"]) if os.environ.get("ADT_FFT_SIZE") else None)
    if forced and forced not in SUPPORTED_FFT:
        raise ValueError(f"fft_size {forced} not in {SUPPORTED_FFT}")
    best = None
    for n, cost in _POINT_COST.items():
        if n not in SUPPORTED_FFT or (forced and n != forced):
            continue
        for segs in range(1, _MAX_SEGMENTS + 1):
            per = -(-n_taps // segs)
            hop = n - (per - 1) - _ALIGN_SLACK
            if hop < 32:
                continue
            score = segs * cost * n / hop
            if best is None or score < best[0] - 1e-12:
                best = (score, n, segs, per)
            if hop >= n // 2:
                break              # more segments only add passes from here on
    if best is None:
        raise ValueError(f"a {n_taps}-tap filter cannot be planned with FFT sizes {SUPPORTED_FFT}")
    return best[1], best[2], best[3]


def plan_filter(taps: np.ndarray, delay: int, fft_size: int | None = None):
    """List of BlockPlans (one per tap segment) for  y[m] = sum_k taps[k] x[m - delay - k]."""
    taps = np.asarray(taps, dtype=np.float64)
    n, segs, per = _plan_sizes(len(taps), fft_size)
    if segs == 1:
        return [plan_block(taps, delay, n)]
    return [plan_block(taps[i:i + per], delay + i, n) for i in range(0, len(taps), per)]


def plan_block(taps: np.ndarray, delay: int, fft_size: int | None = None) -> BlockPlan:
    """Overlap-save geometry for  y[m] = sum_k taps[k] x[m - delay - k].

    The filter is placed circularly with shift s = -c (c = its centre), so a
    symmetric (linear-phase) filter gets a purely REAL spectrum.  Circular
    output index n is alias-free for n in [T-1-c, N-1-c]; block b reads the
    window starting at stream index b*hop - back and keeps [n0, n0+hop), with
    back = delay + c + n0."""
    taps = np.asarray(taps, dtype=np.float64)
    t = len(taps)
    n = fft_size or _plan_sizes(t)[0]
    if n not in SUPPORTED_FFT:
        raise ValueError(f"fft_size {n} not in {SUPPORTED_FFT}")
    c = (t - 1) // 2
    symmetric = (t % 2 == 1) and np.allclose(taps, taps[::-1], rtol=0, atol=1e-15)
    if not symmetric:
        # the shift is free for a complex mask: make (delay + c), hence `back`, a multiple of 32
        # so that every window starts on a 128-byte boundary
        c += (-(delay + c)) % 32
    lo = max(t - 1 - c, 0)              # first alias-free index
    n0 = -(-lo // 32) * 32              # 128-byte aligned slice start
    hop = (n - c - n0) // 32 * 32       # last alias-free index is N-1-c
    if hop < 32:
        raise ValueError(f"{t} taps do not fit an FFT of {n}")
    g = np.zeros(n)
    idx = (np.arange(t) - c) % n
    g[idx] = taps
    spec = np.fft.fft(g)
    if symmetric:
        spec = spec.real + 0j
    return BlockPlan(fft_size=n, hop=int(hop), n0=int(n0), back=int(delay + c + n0), mask_is_real=bool(symmetric),
                     mask=spec.astype(np.complex64), n_taps=t, delay=int(delay))
