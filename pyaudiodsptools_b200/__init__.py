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
from . import utility as Utility  # noqa: F401  (the reference's scripts call pyAudioDspTools.Utility.*)
from .utility import (CombineChunks, MakeChunks, MixSignals, MonoWavToNumpy16BitInt, MonoWavToNumpyFloat,  # noqa: F401
                      NumpyFloatToWav, StereoWavToNumpyFloat)

# The reference's cupy twins (EffectFFTFilterGPU.py, EffectEQ3BandFFTGPU.py) have the same constructors and
# .apply protocol; here every device already runs on the GPU, so the *GPU names are the same classes (numpy
# arrays in and out instead of cupy arrays).
CreateLowCutFilterGPU = CreateLowCutFilter
CreateHighCutFilterGPU = CreateHighCutFilter
CreateEQ3BandFFTGPU = CreateEQ3BandFFT

__all__ = ["config", "CreateHighCutFilter", "CreateLowCutFilter", "CreateEQ3BandFFT", "CreateEQ3Band",
           "CreateSaturator", "CreateSoftClipper", "CreateDelay", "MakeChunks", "CombineChunks", "AdtError",
           "CreateLowCutFilterGPU", "CreateHighCutFilterGPU", "CreateEQ3BandFFTGPU", "Utility", "MixSignals",
           "MonoWavToNumpyFloat", "MonoWavToNumpy16BitInt", "StereoWavToNumpyFloat", "NumpyFloatToWav"]
