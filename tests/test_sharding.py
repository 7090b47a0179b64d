"""Channel sharding: host-side partition logic (CPU, gloo world_size 2) and, on GPUs, the NCCL
scatter / gather of channel rows (adt_comm_*)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from pyaudiodsptools_b200 import sharding


def test_partition_is_contiguous_and_complete():
    for n, w in ((8192, 8), (1000, 8), (7, 4), (3, 8), (4096 * 2, 4)):
        parts = [sharding.channel_range(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 2                      # balanced ...
        assert all(a % 2 == 0 or a == n for a, _ in parts)       # ... and stereo pairs never straddle two GPUs


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from pyaudiodsptools_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_ch, n = 10, 64
full = np.arange(n_ch * n, dtype=np.float32).reshape(n_ch, n) if rank == 0 else None
shard = sharding.scatter_channels_host(full, n_ch, n, src=0)          # host-side plumbing over gloo
lo, hi = sharding.channel_range(n_ch, world, rank)
assert shard.shape == (hi - lo, n) and shard[0, 0] == lo * n
out = sharding.gather_channels_host(shard * 2, n_ch, dst=0)
if rank == 0:
    assert np.array_equal(out, np.arange(n_ch * n, dtype=np.float32).reshape(n_ch, n) * 2)
mx = sharding.max_over_ranks(float(rank + 1))
assert mx == world
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world2_scatter_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from pyaudiodsptools_b200 import sharding, _native
import pyaudiodsptools_b200 as adt
import oracle
dist.init_process_group("gloo")                      # rendezvous only; the data path is our own NCCL communicator
rank, world = dist.get_rank(), dist.get_world_size()
ctx = _native.default_context(rank)
comm = sharding.Communicator(ctx, rank, world)
fs, c, per, n = 44100, 1024, 6, 5 * 1024
adt.config.initialize(fs, c)
full = np.random.default_rng(0).uniform(-1, 1, (world * per, n)).astype(np.float32)
d_full = ctx.malloc(full.nbytes); d_in = ctx.malloc(per * n * 4); d_out = ctx.malloc(per * n * 4)
if rank == 0: ctx.h2d(d_full, full)
comm.scatter_rows(d_full, d_in, per, n, root=0)
dev = adt.CreateLowCutFilter(800, channels=per, device=rank)
dev.process_device(d_in, n, n, d_out, n, n, per)
comm.gather_rows(d_out, d_full, per, n, root=0)
comm.barrier()
if rank == 0:
    y = np.empty_like(full); ctx.d2h(y, d_full)
    taps = oracle.lowcut_taps(fs, c, 800)
    for ch in (0, per - 1, per, world * per - 1):
        want = oracle.fir_stream_f64(taps, c, full[ch])
        assert np.sqrt(np.mean((y[ch] - want) ** 2)) < 2e-6, ch
# uneven shards: 2*world + 3 channels split by channel_range (pairs never straddle), scatterv / gatherv
n_ch = 2 * world + 3
lo, hi = sharding.channel_range(n_ch, world, rank)
full2 = np.random.default_rng(1).uniform(-1, 1, (n_ch, n)).astype(np.float32)
d_f2 = ctx.malloc(full2.nbytes); d_i2 = ctx.malloc(max(hi - lo, 1) * n * 4); d_o2 = ctx.malloc(max(hi - lo, 1) * n * 4)
if rank == 0: ctx.h2d(d_f2, full2)
comm.scatter_channels(d_f2, d_i2, n_ch, n, root=0)
if hi > lo:
    dev2 = adt.CreateLowCutFilter(800, channels=hi - lo, device=rank)
    dev2.process_device(d_i2, n, n, d_o2, n, n, hi - lo)
comm.gather_channels(d_o2, d_f2, n_ch, n, root=0)
comm.barrier()
if rank == 0:
    y2 = np.empty_like(full2); ctx.d2h(y2, d_f2)
    for ch in range(n_ch):
        want = oracle.fir_stream_f64(taps, c, full2[ch])
        assert np.sqrt(np.mean((y2[ch] - want) ** 2)) < 2e-6, ("uneven", ch)
comm.close()
dist.destroy_process_group()
print("rank", rank, "nccl ok")
"""


@pytest.mark.gpu
def test_nccl_scatter_filter_gather(tmp_path):
    from pyaudiodsptools_b200 import _native
    n_gpus = _native.device_count()
    if n_gpus < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2); logs of the 2- and 8-GPU runs: profiles/r02_nccl_*.log")
    world = int(os.environ.get("ADT_TEST_WORLD", "0")) or min(n_gpus, 8)
    script = tmp_path / "n.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29614", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("nccl ok") == world
    print(r.stdout[-1500:])
