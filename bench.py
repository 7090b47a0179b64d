#!/usr/bin/env python
"""bench.py — Msamples/s of the FFT filter hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A *pass* is the hot path run once over one batch of synthetic input (default:
BASELINE configs[1] — low-cut 800 Hz, 1000 mono channels x 10 s @ 44.1 kHz,
chunk 4096, per GPU) = one kernel launch.  A *step* is ``passes_per_step``
passes (config key; ceil(1500 / K), so that the timed region of K steps lasts
over a second whatever K is and the clock samples come from inside it;
``--passes`` overrides).  Rank 0 prints ONE JSON line:

* ``value``      whole-job Msamples/s with inputs resident in HBM (device-timed,
                 CUDA events on the launching stream, max over ranks);
* ``e2e``        the same metric through the public device-class API with
                 pinned HOST buffers, H2D + kernel + D2H inside the timed region;
* ``roofline``   achieved algorithmic GB/s (8 B per output sample, SURVEY.md
                 §8(d) M2) of the fused FIR kernel vs the measured HBM peak;
* ``cpu_baseline`` the unmodified reference (``oracle/_ref``, kind "reference";
                 the numpy oracle port if the copy is absent, kind "port") on the
                 host cores, bounded sample, with the CPU quota it ran under and a
                 single-core Example1-loop figure;
* ``scatter_gather`` (N > 1) rank 0 holds all N x channels rows on its GPU:
                 adt_comm scatter -> kernel -> adt_comm gather (our own NCCL
                 communicator), per-phase device times, max over ranks;
* ``secondary``  device-timed fraction of the HBM roofline for the other BASELINE
                 configs (EQ, chunk sweep, 96 kHz stereo) and the biquad chain.

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
def _cpu_device(kind, ctor_args, fs, chunk):
    """One reference device object: the unmodified reference when oracle/_ref exists, else the oracle port."""
    import oracle.ref
    ref = oracle.ref.load()
    if ref is not None:
        ref.config.initialize(fs, chunk)
        if kind == "eq3fft":
            return ref.CreateEQ3BandFFT(*ctor_args)
        return (ref.CreateLowCutFilter if kind == "lowcut" else ref.CreateHighCutFilter)(*ctor_args)
    import oracle
    if kind == "eq3fft":
        return oracle.SlidingFftEq3(fs, chunk, *ctor_args)
    return oracle.SlidingFftFilter(fs, chunk, ctor_args[0], kind)


def _cpu_kind():
    import oracle.ref
    return "reference" if oracle.ref.available() else "port"


def _cpu_worker(args):
    kind, ctor_args, fs, chunk, n_samples, n_channels, seed = args
    rng = np.random.default_rng(seed)
    done = 0
    for _ in range(n_channels):
        x = rng.uniform(-1, 1, n_samples).astype(np.float32)
        dev = _cpu_device(kind, ctor_args, fs, chunk)          # one device object per channel (Example2.py:13-14)
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


def cpu_environment():
    """What the CPU arm actually had: affinity, cgroup quota (cpu.max), load average."""
    env = {"affinity_cpus": len(os.sched_getaffinity(0)), "os_cpu_count": os.cpu_count()}
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                env["cgroup_" + os.path.basename(path)] = f.read().strip()
        except OSError:
            pass
    quota = env.get("cgroup_cpu.max", "max").split()
    if len(quota) == 2 and quota[0] != "max":
        env["cgroup_cpus"] = float(quota[0]) / float(quota[1])
    try:
        env["loadavg"] = list(os.getloadavg())
    except OSError:
        pass
    return env


def cpu_single_core(wl, budget_s=3.0):
    """the Example1.py:15-18 loop, one process, one channel at a time."""
    kind, ctor_args, fs, chunk, _, seconds, _ = wl
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        done += _cpu_worker((kind, ctor_args, fs, chunk, fs * seconds, 1, 99))
    dt = time.perf_counter() - t0
    return {"value": done / dt / 1e6, "unit": "Msamples/s", "cores": 1,
            "what": f"Example1.py:15-18 loop, one process, {done // (fs * seconds)} channel(s) x {fs * seconds} samples"}


def passes_per_step(args):
    """Deterministic from --steps alone (both arms print the same config): K steps x R passes >= 1500 passes."""
    return args.passes if args.passes > 0 else max(1, -(-1500 // max(args.steps, 1)))


def workload_config(args, wl, channels, world):
    kind, ctor_args, fs, chunk, _, seconds, desc = wl
    return {"workload": desc + (" [16-bit PCM in/out, conversions fused]" if args.io == "i16" else ""),
            "sampling_rate": fs, "chunk": chunk, "channels_per_gpu": channels, "samples_per_channel": fs * seconds,
            "io": args.io, "passes_per_step": passes_per_step(args),
            "l2": "inputs larger than L2 (%.2f GB read + %.2f GB written per pass vs 126 MB L2)" % (
                channels * fs * seconds * (2 if args.io == "i16" else 4) / 1e9,
                channels * (-(-fs * seconds // chunk) * chunk) * (2 if args.io == "i16" else 4) / 1e9),
            "parallelism": f"channel-sharded x{world}, no data-path collective"}


def run_reference_arm(args, wl, rank, channels, world):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    kind, ctor_args, fs, chunk, _, seconds, desc = wl
    budget_s = 120.0                    # whole timed region, whatever K is
    env = cpu_environment()
    single = cpu_single_core(wl)
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
    kind_s = _cpu_kind()
    sample = (f"{per_core * cores} channels x {fs * seconds} samples per step over {cores} processes (one device object "
              f"per channel; channels per step sized so that {args.steps} steps take about {budget_s:.0f} s); "
              + ("the unmodified reference package (oracle/_ref copy of pyAudioDspTools), EffectFFTFilter.apply / "
                 "EffectEQ3BandFFT.apply per chunk" if kind_s == "reference" else
                 "numpy oracle port of the reference's apply (oracle/_ref absent)"))
    line = {
        "impl": "reference", "metric": "Msamples/s overlap-add FFT filter, chunk=%d" % chunk, "value": v,
        "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, channels, world),
        "cpu_arithmetic": "complex64/complex128 numpy pocketfft on float32 chunks (the reference's own types)",
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": kind_s, "sample": sample,
                         "environment": env, "single_core": single},
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
def _kernel_src_sha():
    """Hash of the sources of the hot kernel (arithmetic, kernel body, its instantiation parameters): a committed
    ncu traffic figure is only quoted while it matches."""
    import hashlib
    h = hashlib.sha256()
    for name in ("fft_core.cuh", "fir_kernel.cuh", "fir_k8192.cu"):
        with open(os.path.join(ROOT, "pyaudiodsptools_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def _closed_form(taps, chunk, x_row, n_out):
    """float64 closed form of the device stream: the parity yardstick inside the bench."""
    import oracle
    from scipy.signal import fftconvolve
    full = fftconvolve(np.asarray(x_row, dtype=np.float64), taps)
    want = np.zeros(n_out)
    d = oracle.stream_delay(chunk)
    want[d:] = full[: n_out - d]
    return want


def _fill_device_rows(ctx, dptr, base, rows, row_bytes):
    """Upload `base` (k rows) once and replicate it on the device until `rows` rows are filled."""
    k = base.shape[0]
    ctx.h2d(dptr, base)
    done = k
    while done < rows:
        n = min(done, rows - done)
        ctx.d2d(dptr + done * row_bytes, dptr, n * row_bytes)
        done += n
    ctx.sync()


def measure_secondary(adt, ctx, name, local_rank, peak, passes=10):
    """Device-timed fraction of the HBM roofline for one of the other BASELINE workloads (inputs resident)."""
    kind, ctor_args, fs, chunk, channels, seconds, desc = WORKLOADS[name]
    adt.config.initialize(fs, chunk)
    ctor = {"lowcut": adt.CreateLowCutFilter, "highcut": adt.CreateHighCutFilter, "eq3fft": adt.CreateEQ3BandFFT}[kind]
    dev = ctor(*ctor_args, channels=1, device=local_rank)
    n_in = fs * seconds
    n_out = dev.out_length(n_in)
    base = np.random.default_rng(77).uniform(-1, 1, (16, n_in)).astype(np.float32)
    dx, dy = ctx.malloc(channels * n_in * 4), ctx.malloc(channels * n_out * 4)
    try:
        _fill_device_rows(ctx, dx, base, channels, n_in * 4)
        run = lambda: dev.process_device(dx, n_in, n_in, dy, n_out, n_out, channels)
        for _ in range(3):
            run()
        ctx.sync()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(passes):
            run()
        e1.record()
        ms_burst = e0.elapsed_ms(e1) / passes          # a short burst, still at boost clocks
        # ... then ~0.6 s of back-to-back passes, like the headline: the sustained figure under the power cap
        passes = max(passes, int(600.0 / max(ms_burst, 1e-3)))
        e0.record()
        for _ in range(passes):
            run()
        e1.record()
        ms = e0.elapsed_ms(e1) / passes
        yrow = np.empty((1, n_out), np.float32)
        ctx.d2h(yrow, dy + (channels - 1) * n_out * 4)
        err = float(np.sqrt(np.mean((yrow[0] - _closed_form(dev.taps, chunk, base[(channels - 1) % 16], n_out)) ** 2)))
    finally:
        ctx.free(dx); ctx.free(dy)
    gbs = 8.0 * channels * n_out / (ms * 1e-3) / 1e9
    return {"workload": name, "what": desc, "ms_per_pass": ms, "Msamples_s": channels * n_out / (ms * 1e-3) / 1e6,
            "GBps": gbs, "frac": gbs / peak, "ms_per_pass_burst": ms_burst,
            "frac_burst": 8.0 * channels * n_out / (ms_burst * 1e-3) / 1e9 / peak,
            "fft_size": dev.plan.fft_size, "hop": dev.plan.hop,
            "segments": getattr(dev, "n_segments", 1), "mask": "real" if dev.plan.mask_is_real else "complex",
            "parity_rms_vs_oracle": err, "passes": passes}


def measure_biquad(adt, ctx, local_rank, peak, channels=1000, n=441000):
    """CreateEQ3Band low -> mid -> high chain, device resident, float32 (bit-exact mode)."""
    eq = adt.CreateEQ3Band(100, 2, 700, -4, 8000, 5, channels=channels, device=local_rank)
    if not hasattr(eq, "apply_device"):
        return None
    base = np.random.default_rng(78).uniform(-1, 1, (16, n)).astype(np.float32)
    dx, dy = ctx.malloc(channels * n * 4), ctx.malloc(channels * n * 4)
    try:
        _fill_device_rows(ctx, dx, base, channels, n * 4)
        eq.apply_device(dx, dy, n, n)
        ctx.sync()
        eq.reset()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        eq.apply_device(dx, dy, n, n)
        e1.record()
        ms = e0.elapsed_ms(e1)
        yrow = np.empty((1, n), np.float32)
        ctx.d2h(yrow, dy + (channels - 1) * n * 4)
        import oracle
        o = oracle.Eq3BandBiquad(100, 2, 700, -4, 8000, 5)
        m = 20000
        xr = base[(channels - 1) % 16, :m].copy()
        want = o.applyhighband(o.applymidband(o.applylowband(xr)))
        exact = bool(np.array_equal(yrow[0, :m], want))
    finally:
        ctx.free(dx); ctx.free(dy)
    gbs = 8.0 * channels * n / (ms * 1e-3) / 1e9
    return {"workload": "biquad_chain", "what": f"CreateEQ3Band low->mid->high, {channels} channels x {n} samples, one fused launch, "
            "float32 bit-exact mode", "ms_per_pass": ms, "Msamples_s": channels * n / (ms * 1e-3) / 1e6,
            "Msamples_s_per_band": 3 * channels * n / (ms * 1e-3) / 1e6, "GBps": gbs, "frac": gbs / peak,
            "bit_exact_vs_oracle_first_20000": exact}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lowcut", choices=sorted(WORKLOADS))
    ap.add_argument("--channels", type=int, default=0, help="override channels per GPU")
    ap.add_argument("--fft-size", type=int, default=0, help="override the planner's FFT size")
    ap.add_argument("--passes", type=int, default=0, help="passes per step (0 = ceil(1500 / steps))")
    ap.add_argument("--io", default="f32", choices=["f32", "i16"],
                    help="sample format at the HBM/PCIe boundary: f32 = the reference's float32 chunks (the "
                         "BASELINE metric); i16 = 16-bit PCM fused into load/store (SURVEY 8(f) N3, separate mode)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-scatter", action="store_true", help="skip the scatter -> kernel -> gather leg (N > 1)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
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
        run_reference_arm(args, wl, rank, channels, world)
        return

    cpu_leg = None
    if world == 1 and not args.no_cpu:
        # before any CUDA initialisation, so the worker processes can simply be forked
        import multiprocessing as mp
        cores = len(os.sched_getaffinity(0))
        env = cpu_environment()
        single = cpu_single_core(wl)
        per_core, done, dt = 25, 0, 0.0
        with mp.get_context("fork").Pool(cores) as pool:
            _, t1 = cpu_throughput(wl, 1, pool, cores)
            per_core = int(max(1, min(25, 5.0 / max(t1, 1e-3))))
            while dt < 10.0:
                d_, t_ = cpu_throughput(wl, per_core, pool, cores)
                done += d_
                dt += t_
        kind_s = _cpu_kind()
        cpu_leg = {"value": done / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind_s,
                   "sample": f"{done // (fs * seconds)} channels x {fs * seconds} samples in {dt:.1f} s over {cores} "
                             "processes (one device object per channel), " +
                             ("the unmodified reference package (oracle/_ref)" if kind_s == "reference" else
                              "numpy oracle port of the reference's 3-chunk FFT apply loop"),
                   "environment": env, "single_core": single}

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
    from pyaudiodsptools_b200 import _native
    ctx = _native.default_context(local_rank)
    # the thread that feeds the copy engines and the pinned staging buffers go next to the GPU's PCIe root
    numa = {"bound": False, "skipped": "--no-numa"} if args.no_numa else ctx.bind_host_to_gpu_numa()
    adt.config.initialize(fs, chunk)
    ctor = {"lowcut": adt.CreateLowCutFilter, "highcut": adt.CreateHighCutFilter, "eq3fft": adt.CreateEQ3BandFFT}[kind]
    dev = ctor(*ctor_args, channels=1, device=local_rank, fft_size=args.fft_size or None)
    n_in = fs * seconds
    n_out = dev.out_length(n_in)
    R = passes_per_step(args)

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

    def one_pass():
        if i16:
            dev.process_device_int16(dx, n_in, n_in, dy, n_out, n_out, channels)
        else:
            dev.process_device(dx, n_in, n_in, dy, n_out, n_out, channels)

    for _ in range(args.warmup):
        one_pass()
    ctx.sync()
    # parity spot check of what is being timed (row 0 and last row vs the oracle closed form)
    taps = dev.taps
    yrow = np.empty((1, n_out), io_dtype)
    errs = []
    for row in (0, channels - 1):
        ctx.d2h(yrow, dy + row * n_out * es)
        xin = x_host[row].astype(np.float32) / 32768 if i16 else x_host[row]
        want = _closed_form(taps, chunk, xin, n_out)
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
    for _ in range(args.steps * R):
        one_pass()
    e1.record()
    ctx.sync(); barrier()
    ms_total = max_over_ranks(e0.elapsed_ms(e1))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and clocks.get("samples", 0) < 5:
        # only with --passes 1 and very few steps (ncu runs): the timed region ends before NVML answers a handful
        # of queries, so repeat the same launches untimed for ~0.2 s purely to observe clocks / throttle reasons
        probe = ClockSampler(local_rank)
        probe.start()
        t_end = time.perf_counter() + 0.2
        while time.perf_counter() < t_end:
            for _ in range(10):
                one_pass()
            ctx.sync()
        pc = probe.stop()
        pc["how"] = "timed region too short for 5 NVML samples; probed for 0.2 s of the same launches right after it"
        pc["samples_in_timed_region"] = clocks.get("samples", 0)
        clocks = pc
    ms_step = ms_total / args.steps
    ms_pass = ms_step / R
    samples_pass_all = channels * n_out * world
    value = samples_pass_all / (ms_pass * 1e-3) / 1e6

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
        dt_local = (time.perf_counter() - t0) / e2e_steps
        dt = max_over_ranks(dt_local)
        barrier()
        e2e = {"value": samples_pass_all / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(x_host.nbytes),
               "d2h_bytes_per_step": int(y_host.nbytes), "steps": e2e_steps, "ms_per_step": dt * 1e3,
               "passes_per_step": 1,
               "GBps_h2d_per_rank": x_host.nbytes / dt / 1e9, "GBps_d2h_per_rank": y_host.nbytes / dt / 1e9,
               "GBps_host_memory_all_ranks": (x_host.nbytes + y_host.nbytes) * world / dt / 1e9,
               "numa": numa,
               "api": f"{ctor.__name__}(...).{'process_int16' if i16 else 'process'}(pinned host array)"}
        got = y_host[channels - 1].astype(np.float64) / 32767 if i16 else y_host[channels - 1]
        assert float(np.sqrt(np.mean((got - want) ** 2))) <= (3e-5 if i16 else 1e-5), "e2e parity broken"
        # the platform's transfer ceiling: the same bytes up and down at the same time on every rank, no kernel
        ctx.copy_roundtrip(dx, x_host, y_host, dy)
        barrier(); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.copy_roundtrip(dx, x_host, y_host, dy)
        dt_copy = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        barrier()
        e2e["transfer_only"] = {
            "ms_per_step": dt_copy * 1e3, "GBps_per_rank_each_way": x_host.nbytes / dt_copy / 1e9,
            "fraction_of_e2e_time": dt_copy / dt,
            "what": "concurrent cudaMemcpyAsync H2D + D2H of the step's bytes on all ranks at once, no kernel: the "
                    "ceiling the host / PCIe side of this box allows"}

    # ---- N > 1: the product's own multi-GPU path: scatter -> kernel -> gather over adt_comm (NCCL) ----------
    scatter_gather = None
    if dist and not args.no_scatter and not i16:
        from pyaudiodsptools_b200 import sharding
        comm = sharding.Communicator(ctx, rank, world)
        tot = world * channels
        d_full_in = ctx.malloc(tot * n_in * 4) if rank == 0 else 0
        d_full_out = ctx.malloc(tot * n_out * 4) if rank == 0 else 0
        if rank == 0:                      # every shard is a copy of rank 0's rows (synthetic)
            for r in range(world):
                ctx.d2d(d_full_in + r * channels * n_in * 4, dx, channels * n_in * 4)
        d_in, d_out = ctx.malloc(channels * n_in * 4), ctx.malloc(channels * n_out * 4)
        ev = [ctx.event() for _ in range(4)]
        sums, iters = [0.0, 0.0, 0.0, 0.0], 3
        for it in range(iters + 1):        # first iteration = warm-up (NCCL sets up its P2P channels)
            barrier(); ctx.sync()
            ev[0].record()
            comm.scatter_channels(d_full_in, d_in, tot, n_in, root=0)
            ev[1].record()
            dev.process_device(d_in, n_in, n_in, d_out, n_out, n_out, channels)
            ev[2].record()
            comm.gather_channels(d_out, d_full_out, tot, n_out, root=0)
            ev[3].record()
            ctx.sync()
            t = [ev[0].elapsed_ms(ev[1]), ev[1].elapsed_ms(ev[2]), ev[2].elapsed_ms(ev[3]), ev[0].elapsed_ms(ev[3])]
            t = [max_over_ranks(v) for v in t]
            if it:
                sums = [a + b for a, b in zip(sums, t)]
        ms_sc, ms_k, ms_ga, ms_all = (v / iters for v in sums)
        sg_err = None
        if rank == 0:
            ctx.d2h(yrow, d_full_out + (tot - 1) * n_out * 4)     # last row of the last rank's shard
            sg_err = float(np.sqrt(np.mean((yrow[0] - want) ** 2)))
            assert sg_err <= 1e-5, f"scatter/gather parity broken: {sg_err}"
            ctx.free(d_full_in); ctx.free(d_full_out)
        ctx.free(d_in); ctx.free(d_out)
        comm.close()
        scatter_gather = {
            "impl": "adt_comm (own NCCL communicator in libadt_b200.so: grouped ncclSend/ncclRecv, root = rank 0)",
            "channels_total": tot, "ms_scatter": ms_sc, "ms_kernel": ms_k, "ms_gather": ms_ga, "ms_total": ms_all,
            "GBps_root_egress_scatter": (world - 1) * channels * n_in * 4 / (ms_sc * 1e-3) / 1e9,
            "GBps_root_ingress_gather": (world - 1) * channels * n_out * 4 / (ms_ga * 1e-3) / 1e9,
            "GBps_per_link": (world - 1) * channels * n_in * 4 / (ms_sc * 1e-3) / 1e9 / (world - 1),
            "Msamples_s": tot * n_out / (ms_all * 1e-3) / 1e6, "iters": iters, "parity_rms_vs_oracle": sg_err,
            "note": "every row leaves and re-enters rank 0's HBM over NVLink: bounded by one GPU's link bandwidth "
                    "(770 GB/s per direction measured peer copy), not by the kernels"}

    # ---- the other BASELINE workloads and the biquad, device-timed --------------------
    peak, peak_src = _peaks()
    secondary = None
    if not args.no_secondary and not i16:
        ctx.free(dx); ctx.free(dy)
        dx = dy = 0
        secondary = []
        names = [n for n in ("eq", "highcut512", "highcut1024", "highcut4096", "highcut16384", "stereo96k", "lowcut")
                 if n != args.workload]
        for name in names:
            try:
                rec = measure_secondary(adt, ctx, name, local_rank, peak)
                rec["ms_per_pass"] = max_over_ranks(rec["ms_per_pass"])
                w_ = WORKLOADS[name]
                n_o = -(-w_[2] * w_[5] // w_[3]) * w_[3]
                rec["Msamples_s"] = w_[4] * n_o * world / (rec["ms_per_pass"] * 1e-3) / 1e6      # whole job
                rec["GBps"] = 8.0 * w_[4] * n_o / (rec["ms_per_pass"] * 1e-3) / 1e9              # per GPU
                rec["frac"] = rec["GBps"] / peak
                rec["n_gpus"] = world
            except Exception as e:      # a secondary line must never take the headline down
                rec = {"workload": name, "error": f"{type(e).__name__}: {e}"}
            secondary.append(rec)
        try:
            rec = measure_biquad(adt, ctx, local_rank, peak)
            if rec:
                secondary.append(rec)
        except Exception as e:
            secondary.append({"workload": "biquad_chain", "error": f"{type(e).__name__}: {e}"})
        adt.config.initialize(fs, chunk)

    if rank != 0:
        if dist:
            dist[1].destroy_process_group()
        return

    traffic, traffic_src = None, None
    try:   # DRAM bytes of one launch from the committed `ncu --set full` capture of this workload
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload and channels == 1000 and not i16:
            if tj.get("kernel_src_sha") == _kernel_src_sha():
                traffic = tj["traffic_bytes_per_launch"]
                traffic_src = f"profiles/r02_traffic.json (ncu --set full, kernel sources {tj['kernel_src_sha']})"
            else:
                traffic_src = (f"profiles/r02_traffic.json was captured for kernel sources {tj.get('kernel_src_sha')}, "
                               f"the kernel is now {_kernel_src_sha()}: not quoted")
    except Exception:
        pass
    alg_bytes = 2.0 * es * channels * n_out       # per launch, this rank: one read + one write per sample
    achieved = alg_bytes / (ms_pass * 1e-3) / 1e9
    plan = dev.plan
    flops_per_block = 2 * (5.0 * plan.fft_size * np.log2(plan.fft_size)) + 6 * plan.fft_size
    blocks = -(-n_out // plan.hop) * ((channels + 1) // 2)
    line = {
        "metric": "Msamples/s overlap-add FFT filter, chunk=%d" % chunk, "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 compute, int16 PCM I/O (separate mode, SURVEY 8(f) N3)" if i16 else "f32", "data": "synthetic",
        "config": workload_config(args, wl, channels, world),
        "plan": {"fft_size": plan.fft_size, "hop": plan.hop, "mask": "real" if plan.mask_is_real else "complex",
                 "n_taps": plan.n_taps, "out_samples_per_channel": n_out, "ms_per_pass": ms_pass,
                 "parity_rms_vs_oracle": max(errs)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu)",
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "kernel": "fir_block_kernel",
                     "algorithmic_bytes_per_sample": 2 * es,
                     "fp32_tflops_nominal_radix2_count": flops_per_block * blocks / (ms_pass * 1e-3) / 1e12},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
    }
    if scatter_gather:
        line["scatter_gather"] = scatter_gather
    if secondary is not None:
        line["secondary"] = secondary
    if cpu_leg:
        line["cpu_baseline"] = cpu_leg
    print(json.dumps(line), flush=True)
    if dist:
        dist[1].destroy_process_group()


if __name__ == "__main__":
    main()
