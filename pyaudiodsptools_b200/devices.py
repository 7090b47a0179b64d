"""The reference's device classes for the FFT filter / EQ path, on the GPU.

Same names, constructor arguments and ``.apply(float_array)`` protocol as
  pyAudioDspTools/EffectFFTFilter.py:5-75   CreateHighCutFilter
  pyAudioDspTools/EffectFFTFilter.py:78-151 CreateLowCutFilter
  pyAudioDspTools/EffectEQ3BandFFT.py:23-211 CreateEQ3BandFFT
  pyAudioDspTools/EffectEQ3Band.py:4-180    CreateEQ3Band (biquad; + .apply chain)
so a script switches by changing the import.  Extensions that the reference
does not have (all backward compatible — such calls raise there):
``channels=`` (batch of independent mono streams, input ``[channels, C]``),
``device=`` (GPU ordinal), ``.process(x)`` (whole-buffer mode, equal to the
concatenation of successive ``apply`` results over ``MakeChunks(x)``),
``.reset()``.

Differences kept on purpose: the input chunk is copied to
the device, not retained by reference; a wrong-sized input raises ValueError
*before* the history is touched (the reference corrupts its state first).
Everything runs through the C ABI in ``_native``; without a CUDA device the
constructors raise — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native, config, design


def _snapshot_config():
    fs, chunk = config.sampling_rate, config.chunk_size
    # the reference fails with TypeError on `None // 2` when config was never initialised
    if fs is None or chunk is None:
        raise TypeError("config.initialize(sampling_rate, chunk_size) must be called before creating devices")
    chunk = int(chunk)
    if chunk < 8:
        raise ValueError("chunk_size must be at least 8 (the filter has chunk_size // 2 - 1 taps)")
    return fs, chunk


def _check_out(out, shape, dtype, flat_ok=False):
    """The C library writes through the raw pointer: a wrong-sized or wrong-typed ``out=`` must never reach it
    (explicit checks, not asserts — they survive ``python -O``)."""
    if not isinstance(out, np.ndarray):
        raise TypeError("out must be a numpy array")
    if out.dtype != dtype:
        raise TypeError(f"out must have dtype {np.dtype(dtype).name}, got {out.dtype}")
    if not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
        raise ValueError("out must be C-contiguous and writeable")
    if (out.size != int(np.prod(shape))) if flat_ok else (out.shape != tuple(shape)):
        raise ValueError(f"out must have shape {tuple(shape)}, got {out.shape}")


class _FirDevice:
    """Shared engine: composite taps -> block plan -> adt_fir."""

    def __init__(self, taps, chunk, channels=1, device=0, fft_size=None, epilogue=None):
        if channels < 1:
            raise ValueError("channels must be >= 1")
        self.chunk_size = chunk
        self.channels = int(channels)
        self.taps = taps
        # one block plan per tap segment: a single one unless the filter is too long for one transform
        self.plans = design.plan_filter(taps, design.stream_delay(chunk), fft_size)
        self.plan = self.plans[0]
        self.n_segments = len(self.plans)
        self._ctx = _native.default_context(device)
        descs = (_native.FirDesc * self.n_segments)(*[
            _native.FirDesc(p.fft_size, p.hop, p.n0, p.back, int(p.mask_is_real), chunk, self.channels, 0)
            for p in self.plans])
        masks = [np.ascontiguousarray(p.mask).view(np.float32) for p in self.plans]
        mask_ptrs = (C.c_void_p * self.n_segments)(*[m.ctypes.data for m in masks])
        h = C.c_void_p()
        self._ctx.check(self._ctx.lib.adt_fir_create_segmented(self._ctx.h, self.n_segments, descs, mask_ptrs,
                                                               C.byref(h)))
        self._h = h
        if epilogue is not None and self.n_segments > 1:
            raise ValueError("a store epilogue cannot be fused into a partitioned (long) filter")
        self.set_epilogue(epilogue)

    def set_epilogue(self, shaper=None):
        """Fuse a wave-shaper (consumers.CreateSaturator / CreateSoftClipper) into the kernel's store, or
        detach it with None: every output sample y becomes shaper.apply(y) with no extra pass over HBM."""
        if shaper is None:
            self._ctx.check(self._ctx.lib.adt_fir_set_epilogue(self._h, 0, None))
        else:
            p = shaper._params()
            self._ctx.check(self._ctx.lib.adt_fir_set_epilogue(self._h, shaper.kind, p.ctypes.data))
        self.epilogue = shaper

    # -- streaming: one reference .apply() ---------------------------------
    def apply(self, float32_array_input, out=None):
        """One reference-style step.  For large batches pass arrays from ``dev.context.pinned_empty`` (and an
        ``out=`` buffer of the same kind): page-locked buffers are DMA'd directly, pageable ones are staged."""
        # numpy.concatenate(axis=None) is what makes the reference accept lists / 2-D / int16 input
        flat = np.concatenate((float32_array_input,), axis=None)
        if flat.size != self.channels * self.chunk_size:
            raise ValueError(f"operands could not be broadcast together: expected {self.channels} x "
                             f"{self.chunk_size} samples, got {flat.size}")
        x = np.ascontiguousarray(flat, dtype=np.float32)
        if out is None:
            y = np.empty(self.channels * self.chunk_size, dtype=np.float32)
        else:
            _check_out(out, (flat.size,), np.float32, flat_ok=True)
            y = out.reshape(-1)
        self._ctx.check(self._ctx.lib.adt_fir_apply_host(self._h, x.ctypes.data, y.ctypes.data))
        return y if self.channels == 1 else y.reshape(self.channels, self.chunk_size)

    # -- whole buffer ---------------------------------------------------------
    def out_length(self, n_samples: int) -> int:
        """ceil(n / C) * C — what chunking + per-chunk apply produces (Utility.py:22-27)."""
        return -(-int(n_samples) // self.chunk_size) * self.chunk_size

    def process(self, x, out=None):
        """x: [n] or [rows, n] float32 host array -> [rows, ceil(n/C)*C].
        Independent of the streaming history (starts from silence, like a fresh device)."""
        x = np.asarray(x)
        one_d = x.ndim == 1
        x2 = np.ascontiguousarray(x.reshape(1, -1) if one_d else x, dtype=np.float32)
        rows, n = x2.shape
        n_out = self.out_length(n)
        if out is None:
            out = np.empty((rows, n_out), dtype=np.float32)
        _check_out(out, (rows, n_out), np.float32)
        self._ctx.check(self._ctx.lib.adt_fir_process_host(self._h, x2.ctypes.data, n, n, out.ctypes.data, n_out,
                                                           n_out, rows))
        return out[0] if one_d else out

    def process_int16(self, x, out=None):
        """16-bit PCM in and out: equals ``(process(x / 32768) * 32767).astype(int16)``, i.e. the reference's
        MonoWavToNumpyFloat -> chunks -> apply -> CombineChunks -> NumpyFloatToWav chain (Utility.py:218-238,
        278-312) with both conversions fused into the kernel.  x: int16 [n] or [rows, n]."""
        x = np.asarray(x)
        if x.dtype != np.int16:
            raise TypeError("process_int16 takes int16 samples")
        one_d = x.ndim == 1
        x2 = np.ascontiguousarray(x.reshape(1, -1) if one_d else x)
        rows, n = x2.shape
        n_out = self.out_length(n)
        if out is None:
            out = np.empty((rows, n_out), dtype=np.int16)
        _check_out(out, (rows, n_out), np.int16)
        self._ctx.check(self._ctx.lib.adt_fir_process_host_i16(self._h, x2.ctypes.data, n, n, out.ctypes.data, n_out,
                                                               n_out, rows))
        return out[0] if one_d else out

    def process_device_int16(self, x_dev, in_pitch, n_in, y_dev, out_pitch, n_out, rows):
        self._ctx.check(self._ctx.lib.adt_fir_process_dev_i16(self._h, x_dev, in_pitch, n_in, y_dev, out_pitch,
                                                              n_out, rows))

    def process_device(self, x_dev: int, in_pitch: int, n_in: int, y_dev: int, out_pitch: int, n_out: int, rows: int):
        """Whole-buffer mode on device pointers (async on the context stream)."""
        self._ctx.check(self._ctx.lib.adt_fir_process_dev(self._h, x_dev, in_pitch, n_in, y_dev, out_pitch, n_out,
                                                          rows))

    def reset(self):
        self._ctx.check(self._ctx.lib.adt_fir_reset(self._h))

    @property
    def context(self):
        return self._ctx

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._ctx.h:
                self._ctx.lib.adt_fir_destroy(self._h)
        except Exception:
            pass


class CreateHighCutFilter(_FirDevice):
    """FFT high-cut (low-pass) filter; latency = config.chunk_size.
    Reference: pyAudioDspTools/EffectFFTFilter.py:18-75."""

    def __init__(self, cutoff_frequency=8000, channels=1, device=0, fft_size=None, epilogue=None):
        self.fS, chunk = _snapshot_config()
        self.fH = cutoff_frequency
        self.filter_length = design.filter_length(chunk)
        super().__init__(design.highcut_taps(self.fS, chunk, cutoff_frequency), chunk, channels, device, fft_size,
                         epilogue)


class CreateLowCutFilter(_FirDevice):
    """FFT low-cut (high-pass) filter; latency = config.chunk_size.
    Reference: pyAudioDspTools/EffectFFTFilter.py:91-151."""

    def __init__(self, cutoff_frequency=160, channels=1, device=0, fft_size=None, epilogue=None):
        self.fS, chunk = _snapshot_config()
        self.fH = cutoff_frequency
        self.filter_length = design.filter_length(chunk)
        super().__init__(design.lowcut_taps(self.fS, chunk, cutoff_frequency), chunk, channels, device, fft_size,
                         epilogue)


class CreateEQ3BandFFT(_FirDevice):
    """FFT 3-band EQ (NOT overloaded with defaults, like the reference).
    Reference: pyAudioDspTools/EffectEQ3BandFFT.py:47-211; the three masked
    inverse FFTs + dry mix are one composite mask here."""

    def __init__(self, lowshelf_frequency, lowshelf_db, midband_frequency, midband_db, highshelf_frequency,
                 highshelf_db, channels=1, device=0, fft_size=None, epilogue=None):
        self.fS, chunk = _snapshot_config()
        self.fH_lowshelf, self.lowshelf_db = lowshelf_frequency, lowshelf_db
        self.fH_midband, self.midband_db = midband_frequency, midband_db
        self.fH_highshelf, self.highshelf_db = highshelf_frequency, highshelf_db
        self.filter_length = design.filter_length(chunk)
        taps = design.eq3_taps(self.fS, chunk, lowshelf_frequency, lowshelf_db, midband_frequency, midband_db,
                               highshelf_frequency, highshelf_db)
        super().__init__(taps, chunk, channels, device, fft_size, epilogue)


# ---------------------------------------------------------------------------
# the streaming biquad (the only biquad the reference has)
# ---------------------------------------------------------------------------
class _BiquadBand:
    def __init__(self, ctx, coef5, channels):
        self.ctx, self.coef, self.channels = ctx, np.asarray(coef5, dtype=np.float64), channels
        self._h = {}

    def _handle(self, f64):
        if f64 not in self._h:
            h = C.c_void_p()
            c = (C.c_double * 5)(*self.coef)
            self.ctx.check(self.ctx.lib.adt_biquad_create(self.ctx.h, c, self.channels, int(f64), C.byref(h)))
            self._h[f64] = h
        return self._h[f64]

    def run(self, x):
        x = np.asarray(x)
        f64 = x.dtype == np.float64
        if not f64 and x.dtype != np.float32:
            raise TypeError("CreateEQ3Band on the GPU takes float32 or float64 arrays")
        if len(self._h) == 1 and f64 not in self._h:
            raise TypeError("dtype changed between calls; the filter state lives in the first dtype")
        shape = x.shape
        x2 = np.ascontiguousarray(x.reshape(self.channels, -1))
        y = np.empty_like(x2)
        n = x2.shape[1]
        self.ctx.check(self.ctx.lib.adt_biquad_apply_host(self._handle(f64), x2.ctypes.data, y.ctypes.data, n, n))
        return y.reshape(shape)

    def close(self):
        for h in self._h.values():
            self.ctx.lib.adt_biquad_destroy(h)
        self._h = {}


class CreateEQ3Band:
    """RBJ-cookbook 3-band biquad EQ, per-sample recurrence on the GPU (one
    thread per channel).  Reference: pyAudioDspTools/EffectEQ3Band.py:29-180,
    including its quirks: Fs fixed at 44100 (:33), shelf Q = 1, A =
    sqrt(10**(dB/20)), and the numerator seeing x[n-1..n-3] (:106-113).
    ``apply`` (not in the reference) chains low -> mid -> high."""

    def __init__(self, low_shelf_frequency, low_shelf_gain, mid_frequency, mid_gain, high_shelf_frequency,
                 high_shelf_gain, channels=1, device=0):
        self.Fs = 44100.0
        self.channels = int(channels)
        ctx = _native.default_context(device)
        coefs = biquad_coefficients(low_shelf_frequency, low_shelf_gain, mid_frequency, mid_gain,
                                    high_shelf_frequency, high_shelf_gain)
        self._low, self._mid, self._high = (_BiquadBand(ctx, c, self.channels) for c in coefs)

    def applylowband(self, float_array_input):
        return self._low.run(float_array_input)

    def applymidband(self, float_array_input):
        return self._mid.run(float_array_input)

    def applyhighband(self, float_array_input):
        return self._high.run(float_array_input)

    def apply(self, float_array_input):
        """low -> mid -> high in ONE launch (adt_biquad_chain_apply_host): the three bands run as a pipeline
        over 32-sample tiles, bit-identical to ``applyhighband(applymidband(applylowband(x)))`` and sharing
        the per-band state with those methods."""
        x = np.asarray(float_array_input)
        f64 = x.dtype == np.float64
        if not f64 and x.dtype != np.float32:
            raise TypeError("CreateEQ3Band on the GPU takes float32 or float64 arrays")
        bands = (self._low, self._mid, self._high)
        if any(len(b._h) == 1 and f64 not in b._h for b in bands):
            raise TypeError("dtype changed between calls; the filter state lives in the first dtype")
        x2 = np.ascontiguousarray(x.reshape(self.channels, -1))
        y = np.empty_like(x2)
        n = x2.shape[1]
        ctx = self._low.ctx
        ctx.check(ctx.lib.adt_biquad_chain_apply_host(*(b._handle(f64) for b in bands), x2.ctypes.data,
                                                      y.ctypes.data, n, n))
        return y.reshape(x.shape)

    def apply_device(self, x_dev: int, y_dev: int, pitch: int, n: int, f64: bool = False):
        """The same chain on device-resident planar rows [channels][pitch] (async on the context stream)."""
        ctx = self._low.ctx
        ctx.check(ctx.lib.adt_biquad_chain_apply_dev(*(b._handle(f64) for b in (self._low, self._mid, self._high)),
                                                     x_dev, y_dev, pitch, n))

    def reset(self):
        for b in (self._low, self._mid, self._high):
            for h in b._h.values():
                b.ctx.check(b.ctx.lib.adt_biquad_reset(h))

    def __del__(self):
        try:
            for b in (self._low, self._mid, self._high):
                b.close()
        except Exception:
            pass


def biquad_coefficients(f_low, db_low, f_mid, db_mid, f_high, db_high, fs=44100.0):
    """Normalised (b0,b1,b2,a1,a2)/a0 per band, float64 (EffectEQ3Band.py:45-88)."""
    def shelf(f, db, high):
        a = np.sqrt(10 ** (db / 20))
        w0 = 2 * np.pi * f / fs
        cw, q = np.cos(w0), 1.0
        alpha = np.sin(w0) / 2 * np.sqrt((a + 1 / a) * (1 / q - 1) + 2)
        r = 2 * np.sqrt(a) * alpha
        s = -1.0 if high else 1.0   # the high shelf mirrors the cos terms
        b0 = a * ((a + 1) - s * (a - 1) * cw + r)
        b1 = s * 2 * a * ((a - 1) - s * (a + 1) * cw)
        b2 = a * ((a + 1) - s * (a - 1) * cw - r)
        a0 = (a + 1) + s * (a - 1) * cw + r
        a1 = -s * 2 * ((a - 1) + s * (a + 1) * cw)
        a2 = (a + 1) + s * (a - 1) * cw - r
        return b0, b1, b2, a0, a1, a2

    def peak(f, db):
        a = np.sqrt(10 ** (db / 20))
        w0 = 2 * np.pi * f / fs
        alpha = np.sin(w0) / (2 * 2.5)
        return 1 + alpha * a, -2 * np.cos(w0), 1 - alpha * a, 1 + alpha / a, -2 * np.cos(w0), 1 - alpha / a

    out = []
    for b0, b1, b2, a0, a1, a2 in (shelf(f_low, db_low, False), peak(f_mid, db_mid), shelf(f_high, db_high, True)):
        out.append((b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0))
    return out
