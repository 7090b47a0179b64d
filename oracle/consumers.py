"""Restatements of the reference's in-repo consumers of the FFT filter path:
the pointwise wave-shapers and the feedback delay.  TEST INFRASTRUCTURE — see ``oracle/__init__.py``.
Pinned against golden vectors from the live reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def saturator(x, threshold_db=-20.0, makeup_gain=2.0, mode="hard"):
    """CreateSaturator.apply, pyAudioDspTools/EffectSaturator.py:18-49: above the threshold s the
    magnitude follows s + (a-s)/(1 + ((a-s)/(1-s))**m) (m = 1 'hard', 2 'soft'), anything still above
    1.0 becomes (s+1)/2, sign restored, makeup gain 10**(dB/20).  float32 in -> float32 arithmetic."""
    s = 10 ** (threshold_db / 20)                                  # :19
    m = {"hard": 1, "soft": 2}[mode]                                # :21-24
    x = np.asarray(x)
    neg = x < 0                                                     # :41
    a = np.abs(x)                                                   # :42
    a = np.where(a > s, s + (a - s) / (1 + ((a - s) / (1 - s)) ** m), a)   # :45
    a = np.where(a > 1.0, (s + 1) / 2, a)                           # :46
    a = np.where(neg, -a, a)                                        # :47
    return 10 ** (makeup_gain / 20) * a                             # :48


def soft_clipper(x, drive=0.44):
    """CreateSoftClipper.apply, EffectSoftClipper.py:19-44: y = sign(x) * (1 - |clip(|x|,-1,1) - 1|**(drive+1))."""
    d = drive + 1                                                   # :21
    x = np.asarray(x)
    neg = x < 0                                                     # :37
    a = np.clip(np.abs(x), -1.0, 1.0)                               # :38-40
    a = -1 * (np.abs(a - 1)) ** d + 1                               # :42
    return np.where(neg, -a, a)                                     # :44


class FeedbackDelay:
    """CreateDelay.apply, EffectDelay.py:31-74 (pre-filter flags off: with them on the reference calls
    methods that do not exist, :56,:58).  delay_buffer[D(k+1) : D(k+1)+n] += x * ramp[k] for every
    feedback loop k, output = x + delay_buffer[:n] (or just the buffer when wet), buffer shifted by n."""

    def __init__(self, fs, time_in_ms=500, feedback_loops=2, wet=False):
        self.d = int(time_in_ms * (fs / 1000))                      # :32
        self.wet = wet
        self.buf = np.zeros(int(self.d * (feedback_loops + 2)), dtype="float32")   # :34-35
        self.ramp = np.linspace(0.5, 0.1, num=feedback_loops, dtype="float32")     # :36

    def apply(self, x):
        x = np.array(x, dtype=np.float32)      # (the reference updates the caller's array in place, :66)
        n = len(x)
        for k, r in enumerate(self.ramp):                            # :60-64
            lo = self.d * (k + 1)
            self.buf[lo:lo + n] += x * r
        out = self.buf[:n].copy() if self.wet else x + self.buf[:n]  # :66-69
        self.buf = np.append(self.buf[n:], np.zeros(n, dtype="float32"))   # :71-72
        return out
