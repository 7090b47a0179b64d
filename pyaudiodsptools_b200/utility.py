"""Chunk plumbing either side of the hot path (host side).

Behaviour-compatible with pyAudioDspTools/Utility.py:8-48, without its
accidents: ``MakeChunks`` there pads only when ``len % number_of_chunks != 0``
(:23, a typo for ``% chunk_size``) and then ``numpy.split`` raises for lengths
the test missed; ``CombineChunks`` re-allocates per chunk (O(n^2), :45-47).
Here padding is by ``chunk_size`` and combining is one concatenate; for every
length the reference handles the results are identical.  With the batched
``process()`` entry of the devices neither is needed at all — a
``[channels, samples]`` buffer goes in whole.
"""
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
