"""ctypes binding of ``libmaicos_b200.so`` -- the C ABI declared in ``include/maicos_b200.h``.

The library is the only compute path of this package.  There is **no CPU fallback**: if the
shared object is missing, or no CUDA device is visible, every compute call raises.
"""

from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path

import numpy as np

_LIB_PATH = Path(os.environ.get("MAICOS_B200_LIB") or Path(__file__).resolve().parent / "_lib" / "libmaicos_b200.so")

EBUSY = -5  # MB200_EBUSY: no free slot for a deferred batch
c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_i64_p = ctypes.POINTER(ctypes.c_int64)

#: every symbol ``include/maicos_b200.h`` declares -> (restype, argtypes)
SIGNATURES = {
    "mb200_abi_version": (ctypes.c_int, []),
    "mb200_last_error": (ctypes.c_char_p, []),
    "mb200_device_count": (ctypes.c_int, []),
    "mb200_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "mb200_ctx_destroy": (None, [ctypes.c_void_p]),
    "mb200_ctx_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "mb200_ctx_stream": (ctypes.c_uint64, [ctypes.c_void_p]),
    "mb200_ctx_stats": (ctypes.c_int, [ctypes.c_void_p, c_i64_p]),
    "mb200_ctx_set_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "mb200_ctx_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "mb200_probe_peaks": (ctypes.c_int, [ctypes.c_void_p, c_double_p]),
    "mb200_ctx_set_precision": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "mb200_ctx_set_hist_mode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "mb200_prefetch_frames": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "mb200_prefetch_frames_shared": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]),
    "mb200_ctx_prefetch_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "mb200_xtc_frame": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]),
    "mb200_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "mb200_host_free": (None, [ctypes.c_void_p]),
    "mb200_host_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "mb200_plans_ahead": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]),
    "mb200_plan_describe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_int32, ctypes.c_void_p]),
    "mb200_running_stats": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_sfactor_maxn": (ctypes.c_int, [c_double_p, ctypes.c_double, c_i32_p]),
    "mb200_structure_factor": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_double_p,
         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_structure_factor_block": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_double_p,
         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
         ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_saxs_frames": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
         ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_int, ctypes.c_int32, ctypes.c_int32,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_groups_create": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)],
    ),
    "mb200_groups_destroy": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "mb200_saxs_frames_groups": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_int, ctypes.c_int32,
         ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_saxs_frames_submit": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_int, ctypes.c_int32,
         ctypes.c_int32, ctypes.c_void_p],
    ),
    "mb200_saxs_frames_collect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mb200_saxs_frames_cells": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
         ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p],
    ),
    "mb200_dev_alloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "mb200_dev_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "mb200_dev_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "mb200_dev_write": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "mb200_compound_prologue": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_topology_create": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.POINTER(ctypes.c_void_p)],
    ),
    "mb200_topology_destroy": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "mb200_compound_prologue_topo": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_compound_prologue_window": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
         ctypes.c_void_p],
    ),
    "mb200_velocity_weights": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
         ctypes.c_void_p],
    ),
    "mb200_dielectric_prepare": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
         ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_dipsf_frames": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
         ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
         ctypes.c_void_p],
    ),
    "mb200_profile_hist": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
         ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
         ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb200_profile_hist_window": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
         ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_int, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
}

_lib = None
_lock = threading.Lock()
_contexts: dict[int, "Context"] = {}


class MB200Error(RuntimeError):
    """Raised when a C-ABI call fails (message from ``mb200_last_error``)."""


def library_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library and bind every declared symbol (no compute, no GPU needed)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not _LIB_PATH.exists():
            raise MB200Error(
                f"{_LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C maicos_b200/csrc`. maicos_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(str(_LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        msg = load().mb200_last_error().decode("utf8", "replace")
        raise MB200Error(f"libmaicos_b200 error {rc}: {msg}")


def device_count() -> int:
    return int(load().mb200_device_count())


class Context:
    """One per (process, device); owns the stream, workspaces and the q-plan cache."""

    def __init__(self, device: int):
        lib = load()
        handle = ctypes.c_void_p()
        _check(lib.mb200_ctx_create(int(device), ctypes.byref(handle)))
        self._h = handle
        self.device = int(device)
        mode = _default_precision or os.environ.get("MAICOS_B200_PRECISION", "")
        if mode:  # e.g. MAICOS_B200_PRECISION=fp64 for the CLI; maicos_b200.set_precision() overrides
            self.set_precision(mode)
        mode = os.environ.get("MAICOS_B200_HIST", "")
        if mode:
            self.set_hist_mode(mode)

    @property
    def handle(self):
        return self._h

    def synchronize(self) -> None:
        _check(load().mb200_ctx_synchronize(self._h))

    def set_timing(self, enable: bool) -> None:
        _check(load().mb200_ctx_set_timing(self._h, int(bool(enable))))

    def stats(self) -> dict:
        out = (ctypes.c_int64 * 8)()
        _check(load().mb200_ctx_stats(self._h, out))
        keys = ["launches", "n_valid_q", "ctas", "dmma_steps", "atom_q", "contract_ns", "k_splits", "warp_tiles"]
        return dict(zip(keys, [int(v) for v in out]))

    def set_precision(self, mode: str) -> None:
        """``"i8x5"`` (default: split precision on the tcgen05 tensor cores) or ``"fp64"`` (FP64 DMMA)."""
        _check(load().mb200_ctx_set_precision(self._h, {"fp64": 0, "i8x5": 1, "i8x5-smem": 2}[mode]))

    def set_hist_mode(self, mode: str) -> None:
        """Weighted-histogram kernel: ``"chunks"`` (default: four non-returning shared atomics per element),
        ``"atomics"`` (three returning atomics) or ``"sorted"`` (counting-sort tiles); bitwise identical results."""
        _check(load().mb200_ctx_set_hist_mode(self._h, {"chunks": 0, "atomics": 1, "sorted": 2}[mode]))

    def probe_peaks(self) -> dict:
        """Measured ceilings of this device, in this process: ``i8_tops``, ``fp64_dmma_tflops``, ``hbm_copy_gbs``,
        ``fp64_fma_tflops``."""
        out = (ctypes.c_double * 4)()
        _check(load().mb200_probe_peaks(self._h, out))
        return {"i8_tops": float(out[0]), "fp64_dmma_tflops": float(out[1]), "hbm_copy_gbs": float(out[2]),
                "fp64_fma_tflops": float(out[3])}

    def launch_count(self) -> int:
        return int(load().mb200_ctx_launch_count(self._h))

    def close(self) -> None:
        if self._h is not None and self._h.value:
            load().mb200_ctx_destroy(self._h)
            self._h = None


def default_device() -> int:
    """LOCAL_RANK under torchrun (one process per GPU), else MAICOS_B200_DEVICE, else 0."""
    for var in ("MAICOS_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v is not None and v != "":
            return int(v)
    return 0


_tls = threading.local()
_default_precision = ""


def set_default_precision(mode: str) -> None:
    """Arithmetic mode of every context of this process: those that exist and those created later."""
    global _default_precision
    if mode not in ("fp64", "i8x5", "i8x5-smem"):
        raise KeyError(mode)
    _default_precision = mode
    with _lock:
        existing = list(_contexts.values())
    if not existing:
        existing = [context()]
    for ctx in existing:
        ctx.set_precision(mode)


def set_thread_device(device: int | None) -> None:
    """Device that :func:`context` hands out on the calling thread when none is named (one worker thread per GPU in
    single-process multi-GPU runs, ``maicos_b200.parallel.use_devices``)."""
    _tls.device = device


def context(device: int | None = None) -> Context:
    """Cached context for ``device``; raises :class:`MB200Error` when no GPU is visible."""
    if device is None:
        device = getattr(_tls, "device", None)
    if device is None:
        device = default_device()
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def ptr(a: np.ndarray | None):
    """Raw data pointer of a numpy array (None -> NULL)."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data)


def check(rc: int) -> None:
    _check(rc)


class _PinnedBlock:
    """One page-locked allocation (freed explicitly when its slot grows)."""

    def __init__(self, nbytes: int):
        p = ctypes.c_void_p()
        _check(load().mb200_host_alloc(nbytes, ctypes.byref(p)))
        self.ptr, self.nbytes = p, nbytes
        self.buf = (ctypes.c_char * nbytes).from_address(p.value)

    def free(self):
        if self.ptr:
            load().mb200_host_free(self.ptr)
            self.ptr = None


_pinned_slots: dict[str, _PinnedBlock] = {}
_pinned_free: list[_PinnedBlock] = []      # released blocks kept for the next run (cudaMallocHost is slow)
_PINNED_POOL_BYTES = 2 << 30


def _pinned_acquire(nbytes: int) -> _PinnedBlock:
    best = None
    for blk in _pinned_free:
        if blk.nbytes >= nbytes and (best is None or blk.nbytes < best.nbytes):
            best = blk
    if best is not None and best.nbytes <= 4 * nbytes + (1 << 20):
        _pinned_free.remove(best)
        return best
    return _PinnedBlock(nbytes + nbytes // 4)


def pinned_empty(shape, dtype, slot: str = "default") -> np.ndarray:
    """Page-locked staging array backed by the grow-only block of ``slot``.

    A slot's previous array is invalidated by the next call with the same slot (staging buffers are
    reused batch after batch).  Without a CUDA device (host-logic tests) an ordinary array is
    returned; it is only a buffer, never a compute fallback.
    """
    dtype = np.dtype(dtype)
    count = int(np.prod(tuple(int(x) for x in shape), dtype=np.int64))
    nbytes = max(count * dtype.itemsize, 1)
    if device_count() <= 0:
        return np.empty(shape, dtype=dtype)
    block = _pinned_slots.get(slot)
    if block is None or block.nbytes < nbytes:
        if block is not None:
            _pinned_release_block(block)
        block = _pinned_slots[slot] = _pinned_acquire(nbytes)
    return np.frombuffer(block.buf, dtype=dtype, count=count).reshape(shape)


def _pinned_release_block(block: _PinnedBlock) -> None:
    if sum(b.nbytes for b in _pinned_free) + block.nbytes <= _PINNED_POOL_BYTES:
        _pinned_free.append(block)
    else:
        block.free()


def release_pinned(prefix: str) -> None:
    """Return the staging blocks whose slot name starts with ``prefix`` to the pool (end of a run)."""
    for name in [k for k in _pinned_slots if k.startswith(prefix)]:
        _pinned_release_block(_pinned_slots.pop(name))


class DeviceBuffer:
    """Grow-only device allocation owned by the host layer (intermediates that stay on the GPU)."""

    def __init__(self, ctx: Context):
        self.ctx, self.ptr, self.nbytes = ctx, ctypes.c_void_p(), 0

    def ensure(self, nbytes: int) -> int:
        if nbytes > self.nbytes:
            self.free()
            _check(load().mb200_dev_alloc(self.ctx.handle, int(nbytes + nbytes // 8), ctypes.byref(self.ptr)))
            self.nbytes = int(nbytes + nbytes // 8)
        return self.ptr.value

    def read(self, shape, dtype=np.float64) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        _check(load().mb200_dev_read(self.ctx.handle, self.ptr, ptr(out), out.nbytes))
        return out

    def write(self, array: np.ndarray) -> int:
        """Upload a host array (grows the buffer as needed); returns the device pointer."""
        array = np.ascontiguousarray(array)
        self.ensure(array.nbytes)
        _check(load().mb200_dev_write(self.ctx.handle, self.ptr, ptr(array), array.nbytes))
        return self.ptr.value

    def free(self) -> None:
        if self.ptr:
            _check(load().mb200_dev_free(self.ctx.handle, self.ptr))
        self.ptr, self.nbytes = ctypes.c_void_p(), 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _default_copy_threads() -> int:
    """Staging-copy threads of this process: up to six, but no more than half of this rank's share of the host cores (one
    process per GPU: eight ranks with six copy threads each oversubscribe a 32-core host and slow each other down)."""
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1))
    share = max(1, (os.cpu_count() or 1) // ranks)
    return int(max(1, min(6, share // 2)))


def host_copy(dst: np.ndarray, src: np.ndarray, n_threads: int = 0) -> None:
    """``dst[...] = src`` for C-contiguous arrays of the same dtype and size, copied by a few threads (GIL released)."""
    if n_threads <= 0:
        n_threads = _default_copy_threads()
    if dst.nbytes != src.nbytes or dst.dtype != src.dtype or not (dst.flags.c_contiguous and src.flags.c_contiguous):
        np.copyto(dst, src)
        return
    _check(load().mb200_host_copy(dst.ctypes.data, src.ctypes.data, dst.nbytes, int(n_threads)))
