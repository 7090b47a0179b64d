"""pyaudiodsptools_b200 — B200-native drop-in for the FFT filter / EQ path of
pyAudioDspTools (ArjaanAuinger/pyaudiodsptools).

    import pyaudiodsptools_b200 as pyAudioDspTools
    pyAudioDspTools.config.initialize(44100, 4096)
    dev = pyAudioDspTools.CreateLowCutFilter(800)
    y = dev.apply(chunk)

Python host code -> ctypes -> libadt_b200.so (hand-written CUDA for sm_100a).
No PyTorch, no cupy, and no CPU fallback: constructing a device without a CUDA
GPU raises ``AdtError``.
"""
from . import config  # noqa: F401
from ._native import AdtError  # noqa: F401
from .devices import CreateEQ3Band, CreateEQ3BandFFT, CreateHighCutFilter, CreateLowCutFilter  # noqa: F401
from .consumers import CreateDelay, CreateSaturator, CreateSoftClipper  # noqa: F401
from .utility import CombineChunks, MakeChunks  # noqa: F401

__all__ = ["config", "CreateHighCutFilter", "CreateLowCutFilter", "CreateEQ3BandFFT", "CreateEQ3Band",
           "CreateSaturator", "CreateSoftClipper", "CreateDelay", "MakeChunks", "CombineChunks", "AdtError"]
