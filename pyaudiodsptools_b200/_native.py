"""ctypes binding of libadt_b200.so (include/adt_b200.h).

No PyTorch, no cupy: device memory, streams and events all go through the C
ABI.  Importing this module never needs a GPU; *using* it does, and fails
loudly (AdtError) when there is none — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADT_LIB_PATH") or os.path.join(_HERE, "libadt_b200.so")   # override: A/B of builds
NCCL_UNIQUE_ID_BYTES = 128


class AdtError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"adt_b200 error {status}: {message}")
        self.status = status


class FirDesc(C.Structure):
    _fields_ = [("fft_size", C.c_int32), ("hop", C.c_int32), ("n0", C.c_int32), ("back", C.c_int32),
                ("mask_is_real", C.c_int32), ("chunk", C.c_int32), ("n_channels", C.c_int32),
                ("reserved", C.c_int32)]


_P = C.c_void_p
_F = C.POINTER(C.c_float)
# name -> (restype, argtypes); every symbol include/adt_b200.h declares
PROTOTYPES = {
    "adt_version": (C.c_char_p, []),
    "adt_status_string": (C.c_char_p, [C.c_int]),
    "adt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "adt_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "adt_ctx_destroy": (C.c_int, [_P]),
    "adt_last_error": (C.c_char_p, [_P]),
    "adt_ctx_sync": (C.c_int, [_P]),
    "adt_ctx_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "adt_ctx_device_name": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "adt_ctx_pci_bus_id": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "adt_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "adt_free": (C.c_int, [_P, _P]),
    "adt_malloc_host": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "adt_free_host": (C.c_int, [_P, _P]),
    "adt_memset": (C.c_int, [_P, _P, C.c_int, C.c_size_t]),
    "adt_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "adt_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "adt_memcpy_d2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "adt_copy_roundtrip_host": (C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, C.c_size_t]),
    "adt_event_create": (C.c_int, [_P, C.POINTER(_P)]),
    "adt_event_destroy": (C.c_int, [_P]),
    "adt_event_record": (C.c_int, [_P]),
    "adt_event_elapsed_ms": (C.c_int, [_P, _P, C.POINTER(C.c_float)]),
    "adt_fir_create": (C.c_int, [_P, C.POINTER(FirDesc), _P, C.POINTER(_P)]),
    "adt_fir_create_segmented": (C.c_int, [_P, C.c_int32, C.POINTER(FirDesc), C.POINTER(_P), C.POINTER(_P)]),
    "adt_fir_destroy": (C.c_int, [_P]),
    "adt_fir_process_dev": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32]),
    "adt_fir_process_host": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32]),
    "adt_fir_process_dev_i16": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32]),
    "adt_fir_process_host_i16": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32]),
    "adt_fir_apply_host": (C.c_int, [_P, _P, _P]),
    "adt_fir_apply_dev": (C.c_int, [_P, _P, _P]),
    "adt_fir_reset": (C.c_int, [_P]),
    "adt_fir_set_epilogue": (C.c_int, [_P, C.c_int, _P]),
    "adt_shape_apply_dev": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int64]),
    "adt_shape_apply_host": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int64]),
    "adt_delay_create": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "adt_delay_destroy": (C.c_int, [_P]),
    "adt_delay_apply_dev": (C.c_int, [_P, _P, _P, C.c_int64]),
    "adt_delay_apply_host": (C.c_int, [_P, _P, _P, C.c_int64]),
    "adt_biquad_create": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int32, C.c_int32, C.POINTER(_P)]),
    "adt_biquad_destroy": (C.c_int, [_P]),
    "adt_biquad_reset": (C.c_int, [_P]),
    "adt_biquad_apply_dev": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64]),
    "adt_biquad_apply_host": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64]),
    "adt_biquad_chain_apply_dev": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64]),
    "adt_biquad_chain_apply_host": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64]),
    "adt_comm_unique_id": (C.c_int, [C.c_char_p]),
    "adt_comm_create": (C.c_int, [_P, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "adt_comm_destroy": (C.c_int, [_P]),
    "adt_comm_scatter_rows": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32]),
    "adt_comm_gather_rows": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32]),
    "adt_comm_scatterv_rows": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int64), C.c_int64, C.c_int32]),
    "adt_comm_gatherv_rows": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int64), C.c_int64, C.c_int32]),
    "adt_comm_broadcast": (C.c_int, [_P, _P, C.c_size_t, C.c_int32]),
    "adt_comm_barrier": (C.c_int, [_P]),
}

_lib = None
_lock = threading.Lock()


def load():
    """dlopen the in-tree library and attach prototypes (idempotent)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise AdtError(-1, f"{LIB_PATH} not built; run `make -C pyaudiodsptools_b200/csrc` "
                                   "or `python -c 'import __graft_entry__ as g; g.build()'`")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def device_count() -> int:
    n = C.c_int(0)
    load().adt_device_count(C.byref(n))
    return n.value


def _ptr(a):
    return None if a is None else (a if isinstance(a, int) else a.ctypes.data)


class Context:
    """One device + one stream (adt_ctx).  All calls check status and raise."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _P()
        rc = self.lib.adt_ctx_create(device, C.byref(h))
        if rc != 0:
            msg = self.lib.adt_status_string(rc).decode()
            raise AdtError(rc, f"cannot create a context on CUDA device {device}: {msg} "
                               "(this library has no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise AdtError(rc, self.lib.adt_last_error(self.h).decode() or self.lib.adt_status_string(rc).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.adt_ctx_destroy(self.h)
            self.h = None

    # -- memory ------------------------------------------------------------
    def malloc(self, nbytes: int) -> int:
        p = _P()
        self.check(self.lib.adt_malloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, dptr: int):
        self.check(self.lib.adt_free(self.h, dptr))

    def pinned_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """numpy array backed by page-locked host memory (freed with the array)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = _P()
        self.check(self.lib.adt_malloc_host(self.h, max(n, 1), C.byref(p)))
        buf = (C.c_byte * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        import weakref
        weakref.finalize(buf, _free_pinned, self.lib, self.h, p.value)
        return arr

    def memset(self, dptr, value, nbytes):
        self.check(self.lib.adt_memset(self.h, dptr, value, nbytes))

    def h2d(self, dptr: int, host: np.ndarray):
        host = np.ascontiguousarray(host)
        self.check(self.lib.adt_memcpy_h2d(self.h, dptr, host.ctypes.data, host.nbytes))

    def d2h(self, host: np.ndarray, dptr: int):
        assert host.flags["C_CONTIGUOUS"]
        self.check(self.lib.adt_memcpy_d2h(self.h, host.ctypes.data, dptr, host.nbytes))

    def d2d(self, dst: int, src: int, nbytes: int):
        """Asynchronous device-to-device copy on the context stream."""
        self.check(self.lib.adt_memcpy_d2d(self.h, dst, src, nbytes))

    def copy_roundtrip(self, dst_dev: int, src_host: np.ndarray, dst_host: np.ndarray, src_dev: int):
        """H2D of src_host and D2H into dst_host at the same time (two copy streams); blocks until both are done."""
        self.check(self.lib.adt_copy_roundtrip_host(self.h, dst_dev, src_host.ctypes.data, src_host.nbytes,
                                                    dst_host.ctypes.data, src_dev, dst_host.nbytes))

    def sync(self):
        self.check(self.lib.adt_ctx_sync(self.h))

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        self.check(self.lib.adt_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def device_name(self) -> str:
        buf = C.create_string_buffer(256)
        self.check(self.lib.adt_ctx_device_name(self.h, buf, 256))
        return buf.value.decode()

    def pci_bus_id(self) -> str:
        buf = C.create_string_buffer(32)
        self.check(self.lib.adt_ctx_pci_bus_id(self.h, buf, 32))
        return buf.value.decode().lower()

    def bind_host_to_gpu_numa(self):
        """Pin the calling process to the CPUs of the GPU's NUMA node, so that page-locked staging buffers
        allocated afterwards (first touch / local allocation policy) and the thread that feeds the copy
        engines sit next to the GPU's PCIe root.  Returns a dict describing what was done (for bench.py);
        never raises: on boxes without NUMA information it is a no-op."""
        info = {"pci": None, "numa_node": None, "cpus": None, "bound": False}
        try:
            bus = self.pci_bus_id()
            info["pci"] = bus
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
            info["numa_node"] = node
            if node < 0:
                return info
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = _parse_cpulist(f.read().strip())
            allowed = os.sched_getaffinity(0) & cpus
            info["cpus"] = len(allowed)
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["bound"] = True
                try:   # prefer the node for every later allocation of this process (MPOL_PREFERRED = 1)
                    libc = C.CDLL(None, use_errno=True)
                    mask = C.c_ulong(1 << node)
                    rc = libc.syscall(238, 1, C.byref(mask), C.c_ulong(8 * C.sizeof(C.c_ulong)))
                    info["mempolicy"] = "preferred" if rc == 0 else f"errno {C.get_errno()}"
                except Exception as e:  # pragma: no cover
                    info["mempolicy"] = f"unavailable: {e}"
        except Exception as e:
            info["error"] = str(e)
        return info

    # -- events ------------------------------------------------------------
    def event(self):
        return Event(self)


def _parse_cpulist(text):
    cpus = set()
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _free_pinned(lib, ctx_h, p):
    try:
        lib.adt_free_host(ctx_h, p)
    except Exception:
        pass


class Event:
    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = _P()
        ctx.check(ctx.lib.adt_event_create(ctx.h, C.byref(h)))
        self.h = h

    def record(self):
        self.ctx.check(self.ctx.lib.adt_event_record(self.h))

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float(0)
        self.ctx.check(self.ctx.lib.adt_event_elapsed_ms(self.h, stop.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self.h:
                self.ctx.lib.adt_event_destroy(self.h)
        except Exception:
            pass


_contexts = {}


def default_context(device: int = 0) -> Context:
    """Process-wide context per device (what the device classes use)."""
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx
