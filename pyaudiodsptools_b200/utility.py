"""Chunk plumbing either side of the hot path (host side).

Behaviour-compatible with pyAudioDspTools/Utility.py:8-48, without its
accidents: ``MakeChunks`` there pads only when ``len % number_of_chunks != 0``
(:23, a typo for ``% chunk_size``) and then ``numpy.split`` raises for lengths
the test missed; ``CombineChunks`` re-allocates per chunk (O(n^2), :45-47).
Here padding is by ``chunk_size`` and combining is one concatenate; whenever the
reference's version runs to completion and pads at all (``len % number_of_chunks
!= 0``) the results are identical, and the lengths it trips over work here.  With the batched
``process()`` entry of the devices neither is needed at all — a
``[channels, samples]`` buffer goes in whole.
"""
import wave

import numpy as np

from . import config


def MakeChunks(float32_array_input):
    x = np.asarray(float32_array_input)
    c = int(config.chunk_size)
    n_chunks = -(-len(x) // c)
    pad = n_chunks * c - len(x)
    if pad:
        x = np.append(x, np.zeros(pad, dtype="float32"))
    return np.split(x, n_chunks) if n_chunks else []


def CombineChunks(float_array_input):
    chunks = [np.asarray(c) for c in float_array_input]
    if not chunks:
        return np.array([], dtype="float32")
    return np.concatenate(chunks).astype("float32", copy=False)


# ---- the file boundary of the reference's example scripts (Utility.py:197-312), host side -----------------
# Thin ports so that Example1/2/4 run after swapping the import; for bulk work the devices take 16-bit PCM
# directly (``process_int16``), with both conversions fused into the kernel.
def _read_pcm16(path, channels=None):
    with wave.open(path, "rb") as w:
        if channels is not None and w.getnchannels() != channels:
            raise ValueError(f"This function supports only {'stereo' if channels == 2 else 'mono'} .wav files.")
        return np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16), w.getnchannels()


def MonoWavToNumpy16BitInt(wav_file_path):
    """int16 samples of a 16-bit WAV as they are stored (Utility.py:197-216)."""
    return _read_pcm16(wav_file_path)[0]


def MonoWavToNumpyFloat(wav_file_path):
    """float32 in [-1, 1): int16 / 32768 (Utility.py:218-238)."""
    return _read_pcm16(wav_file_path)[0].astype("float32") / 32768


def StereoWavToNumpyFloat(wav_file_path):
    """(left, right) float32 views of a stereo 16-bit WAV (Utility.py:241-276)."""
    pcm, _ = _read_pcm16(wav_file_path, channels=2)
    frames = pcm.reshape(-1, 2).astype("float32") / 32768
    return frames[:, 0], frames[:, 1]


def NumpyFloatToWav(wav_file_path, numpy_array):
    """Write float samples as 16-bit PCM at config.sampling_rate: int16(x * 32767), truncating, mono [n] or
    stereo [n, 2] / [2, n] (Utility.py:278-312)."""
    a = np.asarray(numpy_array)
    if a.ndim == 2 and a.shape[0] == 2:
        a = a.T
    if not np.any((a >= -1) & (a <= 1)):
        raise ValueError("Array values should be in the range [-1.0, 1.0]")
    with wave.open(wav_file_path, "wb") as w:
        w.setnchannels(1 if a.ndim == 1 else a.shape[1])
        w.setsampwidth(2)
        w.setframerate(int(config.sampling_rate))
        w.writeframes((a * 32767).astype("int16").tobytes())


def MixSignals(*args):
    """Sum of equally long signals, clipped to [-1, 1] (Utility.py:51-72)."""
    total = np.zeros(len(args[0]))
    for sig in args:
        if len(sig) != len(total):
            raise Exception("Something went wrong. Make sure, that the Numpy arrays are equal in length.")
        total = total + sig
    return np.clip(total, -1.0, 1.0)
