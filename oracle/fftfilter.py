"""numpy restatement of the reference's FFT filter / EQ devices (test oracle).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Every function cites the
reference lines it restates (paths relative to /root/reference).

Two independent forms are given:

* the *windowed* form (``SlidingFftFilter`` / ``SlidingFftEq3``): literally the
  reference algorithm — keep three chunks, FFT the 3C window, multiply by a
  complex mask, inverse FFT, keep the middle slice;
* the *closed* form (``fir_stream_f64``): the same result written as one causal
  float64 linear convolution of the whole stream (SURVEY.md §8(a) A3/A5,
  Appendix A.4).  This is what the CUDA path is ultimately compared with at
  sizes where the windowed form is slow, and it guards against a numpy
  version silently moving the windowed form.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------
# tap design (float64, exactly the arithmetic of the reference constructors)
# --------------------------------------------------------------------------
def _tap_count(chunk: int) -> int:
    # pyAudioDspTools/EffectFFTFilter.py:22 / :95, EffectEQ3BandFFT.py:65
    return chunk // 2 - 1


def _windowed_sinc(cutoff_hz: float, fs: float, n_taps: int, window: np.ndarray) -> np.ndarray:
    """Unity-DC-gain windowed-sinc low-pass.

    EffectFFTFilter.py:28-37 (Blackman) and EffectEQ3BandFFT.py:72-79, 95-102,
    112-119, 122-129 (Kaiser beta=6): sinc(2*fc/fs*(k-(L-1)/2)) * window, / sum.
    """
    k = np.arange(n_taps)
    h = np.sinc(2 * cutoff_hz / fs * (k - (n_taps - 1) / 2))
    h = h * window
    return h / np.sum(h)


def _invert(h: np.ndarray) -> np.ndarray:
    """Spectral inversion delta - h (EffectFFTFilter.py:112-113,
    EffectEQ3BandFFT.py:82-83, 132-133)."""
    g = -h
    g[(len(h) - 1) // 2] += 1
    return g


def highcut_taps(fs: float, chunk: int, cutoff_hz: float) -> np.ndarray:
    """CreateHighCutFilter.__init__ taps, EffectFFTFilter.py:20-37."""
    n = _tap_count(chunk)
    return _windowed_sinc(cutoff_hz, fs, n, np.blackman(n))


def lowcut_taps(fs: float, chunk: int, cutoff_hz: float) -> np.ndarray:
    """CreateLowCutFilter.__init__ taps, EffectFFTFilter.py:93-113."""
    return _invert(highcut_taps(fs, chunk, cutoff_hz))


def eq3_band_taps(fs: float, chunk: int, f_low: float, f_mid: float, f_high: float):
    """The four Kaiser(6.0) designs of CreateEQ3BandFFT.__init__.

    Returns (h_highshelf, h_lowshelf, h_mid_lowpass, h_mid_highpass):
    EffectEQ3BandFFT.py:72-83 (inverted LP at 0.75*f_high), :95-102 (LP at
    1.25*f_low), :112-119 (LP at 1.25*f_mid), :122-133 (inverted LP at 0.75*f_mid).
    """
    n = _tap_count(chunk)
    win = np.kaiser(n, 6.0)
    h_hs = _invert(_windowed_sinc(f_high - f_high / 4, fs, n, win))
    h_ls = _windowed_sinc(f_low + f_low / 4, fs, n, win)
    h_mlp = _windowed_sinc(f_mid + f_mid / 4, fs, n, win)
    h_mhp = _invert(_windowed_sinc(f_mid - f_mid / 4, fs, n, win))
    return h_hs, h_ls, h_mlp, h_mhp


def _padded_mask(h: np.ndarray, chunk: int) -> np.ndarray:
    """Zero-pad taps to 3*chunk and FFT (EffectFFTFilter.py:45-47 / :121-123:
    append C-L+1 zeros -> C+1, then 2*(C+1)-3 more -> 3C)."""
    buf = np.zeros(3 * chunk)
    buf[: len(h)] = h
    return np.fft.fft(buf)


def _slice_bounds(chunk: int):
    # EffectFFTFilter.py:24-25: start = C + L//2, end offset = C - L//2
    n = _tap_count(chunk)
    return chunk + n // 2, chunk - n // 2


# --------------------------------------------------------------------------
# windowed form: the reference's per-chunk algorithm
# --------------------------------------------------------------------------
class _ThreeChunkWindow:
    """History of the last three chunks, zeros (float64) at start
    (EffectFFTFilter.py:116-118, rotation :139-141)."""

    def __init__(self, chunk: int):
        self.chunk = chunk
        self.hist = [np.zeros(chunk), np.zeros(chunk), np.zeros(chunk)]  # oldest .. newest

    def push(self, x) -> np.ndarray:
        self.hist = [self.hist[1], self.hist[2], x]
        # numpy.concatenate(axis=None) flattens each piece (EffectFFTFilter.py:143-144)
        return np.concatenate(self.hist, axis=None)


class SlidingFftFilter:
    """CreateHighCutFilter / CreateLowCutFilter .apply (EffectFFTFilter.py:49-75,
    125-151): out = ifft(fft(window3C) * mask)[S0:-S1].real as float32."""

    def __init__(self, fs: float, chunk: int, cutoff_hz: float, kind: str):
        taps = {"highcut": highcut_taps, "lowcut": lowcut_taps}[kind](fs, chunk, cutoff_hz)
        self.taps = taps
        self.mask = _padded_mask(taps, chunk)
        self.s0, self.s1 = _slice_bounds(chunk)
        self.win = _ThreeChunkWindow(chunk)

    def apply(self, x) -> np.ndarray:
        spec = np.fft.fft(self.win.push(x))
        y = np.fft.ifft(spec * self.mask)[self.s0:-self.s1]
        return y.real.astype("float32")


class SlidingFftEq3:
    """CreateEQ3BandFFT.apply (EffectEQ3BandFFT.py:156-211): one forward FFT,
    three masked inverse FFTs, each band scaled by (10**(dB/20) - 1) and mixed
    with the dry middle chunk."""

    def __init__(self, fs, chunk, f_low, db_low, f_mid, db_mid, f_high, db_high):
        h_hs, h_ls, h_mlp, h_mhp = eq3_band_taps(fs, chunk, f_low, f_mid, f_high)
        self.m_hs = _padded_mask(h_hs, chunk)
        self.m_ls = _padded_mask(h_ls, chunk)
        self.m_mid = _padded_mask(h_mhp, chunk) * _padded_mask(h_mlp, chunk)  # :188
        self.g = tuple(10 ** (d / 20) for d in (db_high, db_low, db_mid))
        self.s0, self.s1 = _slice_bounds(chunk)
        self.win = _ThreeChunkWindow(chunk)

    def apply(self, x) -> np.ndarray:
        spec = np.fft.fft(self.win.push(x))
        mix = self.win.hist[1]  # dry middle chunk, EffectEQ3BandFFT.py:209
        for mask, gain in zip((self.m_hs, self.m_ls, self.m_mid), self.g):
            band = np.fft.ifft(spec * mask)[self.s0:-self.s1]
            mix = mix + (band * gain - band)  # :195, :200, :205
        return mix.real.astype("float32")


# --------------------------------------------------------------------------
# closed form: one causal float64 FIR over the whole stream
# --------------------------------------------------------------------------
def stream_delay(chunk: int) -> int:
    """Delay D such that out[m] = (h * x)[m - D]: the slice starts at C + L//2
    (EffectFFTFilter.py:24,97) of a window whose newest chunk starts at 2C, so D = C - L//2 — for odd L that is
    C - (L-1)/2 = 3C/4 + 1; identical for filters and the EQ composite, valid for even L too."""
    return chunk - _tap_count(chunk) // 2


def eq3_composite_taps(fs, chunk, f_low, db_low, f_mid, db_mid, f_high, db_high) -> np.ndarray:
    """Single FIR equal to CreateEQ3BandFFT: dry unit tap at
    (L-1)/2 plus (g-1)-weighted shelves plus the (g_mid-1)-weighted *linear
    convolution* of the two mid filters (2L-1 taps, sliced as if centred at
    (L-1)/2 — the reference's extra mid-band delay is inside this vector)."""
    h_hs, h_ls, h_mlp, h_mhp = eq3_band_taps(fs, chunk, f_low, f_mid, f_high)
    n = len(h_hs)
    g_hs, g_ls, g_mid = (10 ** (d / 20) for d in (db_high, db_low, db_mid))
    tot = np.zeros(2 * n - 1)
    tot[:n] += (g_hs - 1) * h_hs + (g_ls - 1) * h_ls
    tot[n // 2] += 1.0          # the dry middle chunk is exactly C behind: tap index C - D = L//2
    tot += (g_mid - 1) * np.convolve(h_mhp, h_mlp)
    return tot


def fir_stream_f64(taps: np.ndarray, chunk: int, x: np.ndarray, n_out: int | None = None) -> np.ndarray:
    """out[m] = sum_k taps[k] * x[m - D - k], x[n<0] = x[n>=len] = 0, float64.

    Equals the concatenation of successive ``apply`` outputs over the
    zero-padded chunked stream (MakeChunks semantics, Utility.py:22-27)."""
    x = np.asarray(x, dtype=np.float64)
    if n_out is None:
        n_out = -(-len(x) // chunk) * chunk
    full = np.convolve(x, np.asarray(taps, dtype=np.float64))
    d = stream_delay(chunk)
    out = np.zeros(n_out)
    m0, m1 = d, min(n_out, d + len(full))
    if m1 > m0:
        out[m0:m1] = full[: m1 - m0]
    return out
