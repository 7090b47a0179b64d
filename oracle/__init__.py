"""CPU oracle for the FFT filter / EQ hot path of pyAudioDspTools.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.

Parity status: PINNED.  The restatement in :mod:`oracle.fftfilter` and
:mod:`oracle.biquad` is checked against outputs of the live reference
(``/root/reference``, imported in the dev container by
``tests/golden/make_golden.py``) stored as fixtures in ``tests/golden/*.npz``.
The reference itself ships no golden vectors or assertions.
"""
from .fftfilter import (  # noqa: F401
    highcut_taps, lowcut_taps, eq3_band_taps, eq3_composite_taps,
    SlidingFftFilter, SlidingFftEq3, fir_stream_f64, stream_delay,
)
from .biquad import Eq3BandBiquad  # noqa: F401
from .consumers import FeedbackDelay, saturator, soft_clipper  # noqa: F401
