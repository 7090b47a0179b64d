#!/usr/bin/env python
"""bench.py — Msamples/s of the FFT filter hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic input
(default: BASELINE configs[1] — low-cut 800 Hz, 1000 mono channels x 10 s @
44.1 kHz, chunk 4096, per GPU).  Rank 0 prints ONE JSON line:

* ``value``      whole-job Msamples/s with inputs resident in HBM (device-timed,
                 CUDA events on the launching stream, max over ranks);
* ``e2e``        the same metric through the public device-class API with
                 pinned HOST buffers, H2D + kernel + D2H inside the timed region;
* ``roofline``   achieved algorithmic GB/s (8 B per output sample, SURVEY.md
                 §8(d) M2) of the fused FIR kernel vs the measured HBM peak;
* ``cpu_baseline`` the oracle port (numpy restatement of the reference's
                 3-chunk FFT algorithm) on the host cores, bounded sample.

``--impl reference`` times only the CPU arm (same metric / config).
No PyTorch on the compute path: torch.distributed is used only for the
barrier / max-over-ranks when launched under torchrun.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, ctor args, fs, chunk, channels per GPU, seconds, description)
    "lowcut": ("lowcut", (800,), 44100, 4096, 1000, 10,
               "EffectFFTFilter low-cut 800 Hz, 1000 mono channels x 10 s @ 44.1 kHz, chunk=4096 (BASELINE configs[1])"),
    "eq": ("eq3fft", (100, 2, 700, -4, 8000, 5), 44100, 4096, 1024, 10,
           "EffectEQ3BandFFT (100,2,700,-4,8000,5), 1024 mono channels per GPU x 10 s, chunk=4096 (BASELINE configs[2])"),
    "stereo96k": ("lowcut", (800,), 96000, 4096, 2048, 10,
                  "stereo CreateLowCutFilter pairs @ 96 kHz, 1024 stereo streams (2048 planar rows) per GPU x 10 s, chunk=4096 (BASELINE configs[4])"),
}
for _c in (512, 1024, 4096, 16384):
    WORKLOADS[f"highcut{_c}"] = ("highcut", (4000,), 44100, _c, 1000, 10,
                                 f"CreateHighCutFilter 4 kHz, 1000 mono channels x 10 s, chunk={_c} (BASELINE configs[3])")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------
def _cpu_worker(args):
    kind, ctor_args, fs, chunk, n_samples, n_channels, seed = args
    import oracle
    rng = np.random.default_rng(seed)
    done = 0
    for _ in range(n_channels):
        x = rng.uniform(-1, 1, n_samples).astype(np.float32)
        if kind == "eq3fft":
            dev = oracle.SlidingFftEq3(fs, chunk, *ctor_args)
        else:
            dev = oracle.SlidingFftFilter(fs, chunk, ctor_args[0], kind)
        pad = (-len(x)) % chunk
        if pad:
            x = np.concatenate([x, np.zeros(pad, dtype=np.float32)])
        for i in range(0, len(x), chunk):            # the Example1.py:15-18 loop
            dev.apply(x[i:i + chunk])
        done += len(x)
    return done


def cpu_throughput(wl, channels_per_core, pool, cores):
    """One bounded CPU step: every core filters `channels_per_core` whole channels."""
    kind, ctor_args, fs, chunk, _, seconds, _ = wl
    n_samples = fs * seconds
    jobs = [(kind, ctor_args, fs, chunk, n_samples, channels_per_core, 1234 + i) for i in range(cores)]
    t0 = time.perf_counter()
    done = sum(pool.map(_cpu_worker, jobs))
    dt = time.perf_counter() - t0
    return done, dt


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    kind, ctor_args, fs, chunk, channels, seconds, desc = wl
    budget_s = 120.0                    # whole timed region, whatever K is
    with mp.get_context("fork").Pool(cores) as pool:
        t1 = None
        for _ in range(max(1, args.warmup)):          # warm-up steps of one channel per core, last one timed
            _, t1 = cpu_throughput(wl, 1, pool, cores)
        per_core = int(max(1, min(25, budget_s / max(args.steps, 1) / max(t1, 1e-3))))
        tot, tt = 0, 0.0
        for _ in range(args.steps):
            d, dt = cpu_throughput(wl, per_core, pool, cores)
            tot += d
            tt += dt
    v = tot / tt / 1e6
    sample = (f"{per_core * cores} channels x {fs * seconds} samples per step over {cores} processes (one device object "
              f"per channel; channels per step sized so that {args.steps} steps take about {budget_s:.0f} s)")
    line = {
        "impl": "reference", "metric": "Msamples/s overlap-add FFT filter, chunk=%d" % chunk, "value": v,
        "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128/float32 (numpy pocketfft)", "data": "synthetic",
        "config": {"workload": desc, "chunk": chunk, "note": "CPU arm: numpy oracle port of EffectFFTFilter.apply / "
                   "EffectEQ3BandFFT.apply (the reference is pure Python and /root/reference does not exist on the GPU box)"},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML (nvidia_ml_py) every ~2 ms from a thread while the timed region runs; the main
    thread sits in a ctypes call that releases the GIL.  Falls back to `nvidia-smi -lms`."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4), ("hw_power_brake", 0x80))

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.nv, self.err = index, [], False, None, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.err = f"nvml unavailable: {e}"
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv, i, power, mask = self.nv, 0, 0.0, 0
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)      # every iteration (cheap)
                if i % 4 == 0:                                                # power / reasons every 4th
                    power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    mask = int(reasons_fn(self.h))
                self.rows.append((sm, power, mask))
            except Exception:
                pass
            i += 1
            time.sleep(0.0005)     # ~2 kHz: plenty of samples, no contention with the launch loop

    def stop(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not started"]}
        self.stop_flag = True
        self.thread.join(timeout=1)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": float(self.max_sm), "reasons": ["no samples"]}
        sm = [r[0] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= int(r[2])
        reasons = [name for name, bit in self.REASONS if mask & bit]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(self.max_sm),
                "power_w_max": max(r[1] for r in self.rows), "samples": len(sm), "reasons": reasons,
                "how": "NVML polled at ~2 kHz from a thread during the timed region"}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lowcut", choices=sorted(WORKLOADS))
    ap.add_argument("--channels", type=int, default=0, help="override channels per GPU")
    ap.add_argument("--fft-size", type=int, default=0, help="override the planner's FFT size")
    ap.add_argument("--io", default="f32", choices=["f32", "i16"],
                    help="sample format at the HBM/PCIe boundary: f32 = the reference's float32 chunks (the "
                         "BASELINE metric); i16 = 16-bit PCM fused into load/store (SURVEY 8(f) N3, separate mode)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    kind, ctor_args, fs, chunk, channels, seconds, desc = wl
    if args.channels:
        channels = args.channels

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    cpu_leg = None
    if world == 1 and not args.no_cpu:
        # before any CUDA initialisation, so the worker processes can simply be forked
        import multiprocessing as mp
        cores = len(os.sched_getaffinity(0))
        per_core, done, dt = 25, 0, 0.0
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_throughput(wl, 2, pool, cores)
            while dt < 10.0:
                d_, t_ = cpu_throughput(wl, per_core, pool, cores)
                done += d_
                dt += t_
        cpu_leg = {"value": done / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                   "sample": f"{done // (fs * seconds)} channels x {fs * seconds} samples in {dt:.1f} s over {cores} "
                             "processes (one device object per channel), numpy oracle port of the reference's "
                             "3-chunk FFT apply loop"}

    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (no "NCCL version ..." banner)
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = (torch, dist_mod)

    def barrier():
        if dist:
            dist[1].barrier()

    def max_over_ranks(v):
        if not dist:
            return v
        t = dist[0].tensor([v], dtype=dist[0].float64, device="cuda")
        dist[1].all_reduce(t, op=dist[1].ReduceOp.MAX)
        return float(t.item())

    import pyaudiodsptools_b200 as adt
    adt.config.initialize(fs, chunk)
    ctor = {"lowcut": adt.CreateLowCutFilter, "highcut": adt.CreateHighCutFilter, "eq3fft": adt.CreateEQ3BandFFT}[kind]
    dev = ctor(*ctor_args, channels=1, device=local_rank, fft_size=args.fft_size or None)
    ctx = dev.context
    n_in = fs * seconds
    n_out = dev.out_length(n_in)

    # synthetic input: uniform(-1,1) float32, seeded per (rank, channel block); 64 distinct rows tiled
    i16 = args.io == "i16"
    io_dtype, es = (np.int16, 2) if i16 else (np.float32, 4)
    x_host = ctx.pinned_empty((channels, n_in), io_dtype)
    base = np.random.default_rng(1234 + rank).uniform(-1, 1, (min(64, channels), n_in)).astype(np.float32)
    if i16:
        base = (base * 0.25 * 32767).astype(np.int16)   # -12 dBFS: filter overshoot must not wrap the int16 output
    for r0 in range(0, channels, base.shape[0]):
        k = min(base.shape[0], channels - r0)
        x_host[r0:r0 + k] = base[:k]
    y_host = ctx.pinned_empty((channels, n_out), io_dtype)
    dx, dy = ctx.malloc(x_host.nbytes), ctx.malloc(y_host.nbytes)
    ctx.h2d(dx, x_host)

    def step():
        if i16:
            dev.process_device_int16(dx, n_in, n_in, dy, n_out, n_out, channels)
        else:
            dev.process_device(dx, n_in, n_in, dy, n_out, n_out, channels)

    for _ in range(args.warmup):
        step()
    ctx.sync()
    # parity spot check of what is being timed (row 0 and last row vs the oracle closed form)
    import oracle
    taps = dev.taps
    yrow = np.empty((1, n_out), io_dtype)
    errs = []
    for row in (0, channels - 1):
        ctx.d2h(yrow, dy + row * n_out * es)
        from scipy.signal import fftconvolve
        xin = x_host[row].astype(np.float32) / 32768 if i16 else x_host[row]
        full = fftconvolve(xin.astype(np.float64), taps)
        want = np.zeros(n_out); d = oracle.stream_delay(chunk)
        want[d:] = full[: n_out - d]
        got = yrow[0].astype(np.float64) / 32767 if i16 else yrow[0]
        errs.append(float(np.sqrt(np.mean((got - want) ** 2))))
    # float32 I/O: north_star tolerance 1e-5 RMS.  int16 I/O: truncation to 16 bits adds ~1/(32767*sqrt(3)) RMS.
    assert max(errs) <= (3e-5 if i16 else 1e-5), f"parity broken in bench: rms {errs}"

    e0, e1 = ctx.event(), ctx.event()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier(); ctx.sync()
    l0 = ctx.launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    ctx.sync(); barrier()
    ms_total = max_over_ranks(e0.elapsed_ms(e1))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and clocks.get("samples", 0) < 5:
        # a very short timed region (few steps x ~1 ms) can end before NVML answers a handful of queries:
        # repeat the same launches untimed for ~0.2 s purely to observe clocks / throttle reasons under this load
        probe = ClockSampler(local_rank)
        probe.start()
        t_end = time.perf_counter() + 0.2
        while time.perf_counter() < t_end:
            for _ in range(10):
                step()
            ctx.sync()
        pc = probe.stop()
        pc["how"] = "timed region too short for 5 NVML samples; probed for 0.2 s of the same launches right after it"
        pc["samples_in_timed_region"] = clocks.get("samples", 0)
        clocks = pc
    ms_step = ms_total / args.steps
    samples_step_all = channels * n_out * world
    value = samples_step_all / (ms_step * 1e-3) / 1e6

    # ---- end to end through the public API with pinned host buffers ------------------
    e2e = None
    if not args.no_e2e:
        dev_b = ctor(*ctor_args, channels=1, device=local_rank, fft_size=args.fft_size or None)
        e2e_steps = max(2, min(args.steps, 5))
        run_e2e = (lambda: dev_b.process_int16(x_host, out=y_host)) if i16 else (lambda: dev_b.process(x_host, out=y_host))
        run_e2e()                              # warm-up (allocates the staging buffers)
        barrier(); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            run_e2e()                          # returns when y_host is complete
        dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        barrier()
        e2e = {"value": samples_step_all / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(x_host.nbytes),
               "d2h_bytes_per_step": int(y_host.nbytes), "steps": e2e_steps, "ms_per_step": dt * 1e3,
               "api": f"{ctor.__name__}(...).{'process_int16' if i16 else 'process'}(pinned host array)"}
        got = y_host[channels - 1].astype(np.float64) / 32767 if i16 else y_host[channels - 1]
        assert float(np.sqrt(np.mean((got - want) ** 2))) <= (3e-5 if i16 else 1e-5), "e2e parity broken"

    if rank != 0:
        if dist:
            dist[1].destroy_process_group()
        return

    peak, peak_src = _peaks()
    traffic = None
    try:   # DRAM bytes of one launch from the committed `ncu --set full` capture of this workload
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload and channels == 1000 and not i16:
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
    alg_bytes = 2.0 * es * channels * n_out       # per launch, this rank: one read + one write per sample
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    plan = dev.plan
    flops_per_block = 2 * (5.0 * plan.fft_size * np.log2(plan.fft_size)) + 6 * plan.fft_size
    blocks = -(-n_out // plan.hop) * ((channels + 1) // 2)
    line = {
        "metric": "Msamples/s overlap-add FFT filter, chunk=%d" % chunk, "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 compute, int16 PCM I/O (separate mode, SURVEY 8(f) N3)" if i16 else "f32", "data": "synthetic",
        "config": {"workload": desc + (" [16-bit PCM in/out, conversions fused]" if i16 else ""), "channels_per_gpu": channels, "samples_per_channel": n_in,
                   "out_samples_per_channel": n_out, "chunk": chunk, "fft_size": plan.fft_size, "hop": plan.hop,
                   "mask": "real" if plan.mask_is_real else "complex", "n_taps": plan.n_taps,
                   "l2": "inputs larger than L2 (%.2f GB read + %.2f GB written per pass vs 126 MB L2)" % (
                       x_host.nbytes / 1e9, y_host.nbytes / 1e9),
                   "parallelism": f"channel-sharded x{world}, no data-path collective",
                   "parity_rms_vs_oracle": max(errs)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu)",
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "kernel": "fir_block_kernel",
                     "algorithmic_bytes_per_sample": 2 * es,
                     "fp32_tflops_nominal_radix2_count": flops_per_block * blocks / (ms_step * 1e-3) / 1e12},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
    }
    if cpu_leg:
        line["cpu_baseline"] = cpu_leg
    print(json.dumps(line), flush=True)
    if dist:
        dist[1].destroy_process_group()


if __name__ == "__main__":
    main()
