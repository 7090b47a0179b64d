#!/usr/bin/env python
"""Latency of the streaming .apply() (one reference-style call per chunk), vs the numpy oracle port."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyaudiodsptools_b200 as adt
import oracle

for chunk, channels in ((4096, 1), (512, 1), (4096, 2), (4096, 64), (4096, 1000)):
    adt.config.initialize(44100, chunk)
    dev = adt.CreateLowCutFilter(800, channels=channels)
    x = np.random.default_rng(0).uniform(-1, 1, (channels, chunk)).astype(np.float32)
    for _ in range(20):
        dev.apply(x)
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        dev.apply(x)
    gpu_us = (time.perf_counter() - t0) / n * 1e6
    o = oracle.SlidingFftFilter(44100, chunk, 800, "lowcut")
    t0 = time.perf_counter()
    for _ in range(100):
        o.apply(x[0])
    cpu_us = (time.perf_counter() - t0) / 100 * 1e6
    print(f"chunk {chunk:5d} channels {channels:5d}: apply() {gpu_us:8.1f} us/call  ({channels * chunk / gpu_us:9.1f} Msamples/s)   "
          f"numpy port {cpu_us:7.1f} us/call/channel   real-time budget {chunk / 44100 * 1e6:8.0f} us")
