"""Global settings read by the device classes at construction time.

Mirror of pyAudioDspTools/config.py:20-39: module globals ``sampling_rate``,
``chunk_size``, ``use_gpu`` set by ``initialize``; devices snapshot them in
``__init__`` and later re-initialisation does not affect existing devices.
``use_gpu`` is stored and ignored here exactly as in the reference (nothing
reads it) — this package *always* runs on the GPU.
"""
sampling_rate = None
chunk_size = None
use_gpu = False
_gpu_available = True


def initialize(sampling_rate, chunk_size, use_gpu=False):  # noqa: A002 - reference signature (config.py:31)
    g = globals()
    g["sampling_rate"] = sampling_rate
    g["chunk_size"] = chunk_size
    g["use_gpu"] = use_gpu
