"""The reference's in-repo consumers of the FFT filter path, on the GPU.

  pyAudioDspTools/EffectSaturator.py:4-49    CreateSaturator   (pointwise wave-shaper)
  pyAudioDspTools/EffectSoftClipper.py:3-44  CreateSoftClipper (pointwise wave-shaper)
  pyAudioDspTools/EffectDelay.py:6-74        CreateDelay       (feedback delay with optional pre-filters)

Same constructor arguments and ``.apply(array)``.  A shaper can also be fused into a filter's store
phase: ``CreateLowCutFilter(800, epilogue=CreateSaturator())`` applies it to every output sample inside
the FIR kernel (no extra pass over HBM).  ``CreateDelay``'s pre-filter flags work here — the reference
calls ``applylowcutfilter`` / ``applyhighcutfilter``, which do not exist (EffectDelay.py:56,58).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native, config


class _Shaper:
    kind = 0

    def __init__(self, device=0):
        self._device = device

    def _params(self):
        raise NotImplementedError

    def apply(self, float_array_input):
        x = np.ascontiguousarray(float_array_input, dtype=np.float32)
        y = np.empty_like(x)
        ctx = _native.default_context(self._device)
        p = self._params()
        ctx.check(ctx.lib.adt_shape_apply_host(ctx.h, self.kind, p.ctypes.data, x.ctypes.data, y.ctypes.data, x.size))
        return y


class CreateSaturator(_Shaper):
    """EffectSaturator.py:18-49."""
    kind = 1

    def __init__(self, saturation_threshold_in_db=-20.0, makeup_gain=2.0, mode='hard', device=0):
        super().__init__(device)
        self.saturation_coeff = 10 ** (saturation_threshold_in_db / 20)
        self.makeup_gain = makeup_gain
        if mode == 'soft':
            self.mode = 2
        if mode == 'hard':
            self.mode = 1      # (any other string leaves .mode unset and .apply fails, as in the reference)

    def _params(self):
        s = self.saturation_coeff
        return np.array([s, 1 - s, (s + 1) / 2, 10 ** (self.makeup_gain / 20), self.mode], dtype=np.float32)


class CreateSoftClipper(_Shaper):
    """EffectSoftClipper.py:19-44."""
    kind = 2

    def __init__(self, drive=0.44, device=0):
        super().__init__(device)
        self.placeholder = True
        self.drive = drive + 1

    def _params(self):
        return np.array([self.drive, 0, 0, 0, 0], dtype=np.float32)


class CreateDelay:
    """EffectDelay.py:31-74, batched over ``channels`` independent mono streams."""

    def __init__(self, time_in_ms=500, feedback_loops=2, lowcut_filter_frequency=40, highcut_filter_frequency=12000,
                 use_lowcut_filter=False, use_highcut_filter=False, wet=False, channels=1, device=0):
        from .devices import CreateHighCutFilter, CreateLowCutFilter
        self.time_in_samples = int(time_in_ms * (config.sampling_rate / 1000))
        self.wet = wet
        self.channels = int(channels)
        self.feedback_ramp = np.linspace(0.5, 0.1, num=feedback_loops, dtype="float32")
        self.use_lowcut_filter = use_lowcut_filter
        self.use_highcut_filter = use_highcut_filter
        self.LowCutFilter = CreateLowCutFilter(lowcut_filter_frequency, channels=channels, device=device)
        self.HighcutFilter = CreateHighCutFilter(highcut_filter_frequency, channels=channels, device=device)
        self._ctx = _native.default_context(device)
        h = C.c_void_p()
        ramp = np.ascontiguousarray(self.feedback_ramp)
        self._ctx.check(self._ctx.lib.adt_delay_create(self._ctx.h, self.time_in_samples, feedback_loops,
                                                       ramp.ctypes.data, int(bool(wet)), self.channels, C.byref(h)))
        self._h = h

    def apply(self, float32_array_input):
        x = np.asarray(float32_array_input, dtype=np.float32)
        shape = x.shape
        if self.use_lowcut_filter:
            x = self.LowCutFilter.apply(x)
        if self.use_highcut_filter:
            x = self.HighcutFilter.apply(x)
        x2 = np.ascontiguousarray(x.reshape(self.channels, -1))
        y = np.empty_like(x2)
        self._ctx.check(self._ctx.lib.adt_delay_apply_host(self._h, x2.ctypes.data, y.ctypes.data, x2.shape[1]))
        return y.reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._ctx.h:
                self._ctx.lib.adt_delay_destroy(self._h)
        except Exception:
            pass
