"""Restatement of the reference's only streaming biquad, CreateEQ3Band
(pyAudioDspTools/EffectEQ3Band.py:29-180).  TEST INFRASTRUCTURE — see
``oracle/__init__.py``.

Quirks kept on purpose: Fs is hard-coded to 44100
(:33); shelf alpha uses Q=1 (:50,:64); A = sqrt(10**(dB/20)) (:48,:55,:62);
the feed-forward path sees x[n-1], x[n-2], x[n-3] because three raw samples
but only two outputs are prepended (:106-109); with float32 input every output
is rounded to float32 before it is fed back (numpy.insert keeps the dtype).
"""
from __future__ import annotations

import numpy as np

_FS = 44100.0


def _shelf_alpha(w0, a):
    q = 1.0
    return np.sin(w0) / 2 * np.sqrt((a + 1 / a) * (1 / q - 1) + 2)


def band_coefficients(f_low, db_low, f_mid, db_mid, f_high, db_high):
    """(b0,b1,b2,a0,a1,a2) for low shelf, peaking mid, high shelf
    (EffectEQ3Band.py:45-88)."""
    out = []
    # low shelf :46-50, :67-72
    a = np.sqrt(10 ** (db_low / 20)); w0 = 2 * np.pi * f_low / _FS
    al = _shelf_alpha(w0, a); c = np.cos(w0); r = 2 * np.sqrt(a) * al
    out.append((a * ((a + 1) - (a - 1) * c + r), 2 * a * ((a - 1) - (a + 1) * c),
                a * ((a + 1) - (a - 1) * c - r),
                (a + 1) + (a - 1) * c + r, -2 * ((a - 1) + (a + 1) * c), (a + 1) + (a - 1) * c - r))
    # peaking :53-57, :75-80
    a = np.sqrt(10 ** (db_mid / 20)); w0 = 2 * np.pi * f_mid / _FS
    al = np.sin(w0) / (2 * 2.5); c = np.cos(w0)
    out.append((1 + al * a, -2 * c, 1 - al * a, 1 + al / a, -2 * c, 1 - al / a))
    # high shelf :60-64, :83-88
    a = np.sqrt(10 ** (db_high / 20)); w0 = 2 * np.pi * f_high / _FS
    al = _shelf_alpha(w0, a); c = np.cos(w0); r = 2 * np.sqrt(a) * al
    out.append((a * ((a + 1) + (a - 1) * c + r), -2 * a * ((a - 1) + (a + 1) * c),
                a * ((a + 1) + (a - 1) * c - r),
                (a + 1) - (a - 1) * c + r, 2 * ((a - 1) - (a + 1) * c), (a + 1) - (a - 1) * c - r))
    return out


class _Band:
    def __init__(self, coef):
        b0, b1, b2, a0, a1, a2 = coef
        # the reference divides inside the loop: b/a0 and a/a0 as Python floats (:112)
        self.c = (b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0)
        self.prev_in = np.zeros(3)   # :37 raw x[-3:]
        self.prev_out = np.zeros(2)  # :36 y[-2:]

    def run(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x)
        u = np.concatenate([self.prev_in.astype(x.dtype), x])    # numpy.insert keeps x.dtype (:107)
        y = np.concatenate([self.prev_out.astype(x.dtype), x])   # (:109)
        self.prev_in = x[-3:].copy()
        c0, c1, c2, c3, c4 = self.c
        for i in range(2, len(y)):
            # float64 arithmetic on numpy scalars, then stored in y's dtype (:112)
            y[i] = (c0 * u[i]) + (c1 * u[i - 1]) + (c2 * u[i - 2]) - (c3 * y[i - 1]) - (c4 * y[i - 2])
        self.prev_out = y[-2:].copy()
        return y[2:]


class Eq3BandBiquad:
    """applylowband / applymidband / applyhighband of CreateEQ3Band
    (EffectEQ3Band.py:90-118, 121-149, 152-180)."""

    def __init__(self, f_low, db_low, f_mid, db_mid, f_high, db_high):
        lo, mid, hi = band_coefficients(f_low, db_low, f_mid, db_mid, f_high, db_high)
        self.low, self.mid, self.high = _Band(lo), _Band(mid), _Band(hi)

    def applylowband(self, x):
        return self.low.run(x)

    def applymidband(self, x):
        return self.mid.run(x)

    def applyhighband(self, x):
        return self.high.run(x)
