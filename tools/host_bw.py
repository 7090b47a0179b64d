"""Aggregate host-memory copy bandwidth with P concurrent processes (explains the e2e ceiling at 8 GPUs:
every rank streams 1.76 GB up and 1.77 GB down through host DRAM per step)."""
import multiprocessing as mp
import sys
import time

import numpy as np


def worker(args):
    i, mb, reps = args
    a = np.ones(mb << 18, dtype=np.float32)
    b = np.empty_like(a)
    np.copyto(b, a)
    t0 = time.perf_counter()
    for _ in range(reps):
        np.copyto(b, a)
    return (time.perf_counter() - t0), 2.0 * a.nbytes * reps


if __name__ == "__main__":
    for p in [int(v) for v in (sys.argv[1:] or ["1", "8", "16", "32"])]:
        with mp.get_context("fork").Pool(p) as pool:
            res = pool.map(worker, [(i, 512, 6) for i in range(p)])
        dt = max(r[0] for r in res)
        print(f"{p:3d} processes: {sum(r[1] for r in res) / dt / 1e9:7.1f} GB/s (read + write bytes, numpy copy of 512 MB buffers)", flush=True)
