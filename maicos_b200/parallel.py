"""One process per GPU: frame / q-block sharding and the single end-of-run allreduce.

The reference has no parallel layer at all (frames are processed serially,
``/root/reference/src/maicos/core/base.py:90-101``).  Here the data-parallel axes of the problem
are used directly (``SURVEY.md`` section 8(e)):

* ``mode="frames"`` -- rank ``r`` analyses frames ``r, r + W, r + 2W, ...`` of the selected
  trajectory slice with no communication inside the loop; at the end ONE ``all_reduce(SUM)`` of a
  packed float64 buffer merges the per-rank frame statistics (Chan's parallel Welford merge:
  every rank contributes ``(n, mean, M2)`` in its own slot plus the running sums and its slots of
  ``timeseries / frames / times``).
* ``mode="qblocks"`` -- every rank sees every frame but evaluates only its share of the q-plan's
  tile groups (``mb200_*_frames(block=rank, n_blocks=world)``); the per-frame binned partial sums are
  allreduced before they enter the frame statistics.  For a single huge frame.

``torch.distributed`` is plumbing only: NCCL over NVLink on GPU boxes, gloo for the CPU tests.
"""

from __future__ import annotations

import contextlib
import os

import numpy as np

_nccl_buffers: dict = {}
_state = {"initialised": False, "rank": 0, "world": 1, "mode": "frames", "device": None, "backend": None, "devices": None}


def init(mode: str = "frames", backend: str | None = None) -> None:
    """Join the process group described by the torchrun environment (no-op for one process)."""
    if mode not in ("frames", "qblocks"):
        raise ValueError("mode must be 'frames' or 'qblocks'")
    _state["mode"] = mode
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        _state.update(initialised=True, rank=0, world=1)
        return
    import torch
    import torch.distributed as dist

    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        _state["device"] = torch.device("cuda", local)
    else:
        _state["device"] = torch.device("cpu")
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            dist.init_process_group(backend=backend, device_id=_state["device"])
        else:
            dist.init_process_group(backend=backend)
    _state.update(initialised=True, rank=dist.get_rank(), world=dist.get_world_size(), backend=backend)


def use_devices(devices=None) -> list[int] | None:
    """Single-process multi-GPU: let every batched analysis of this process deal its batches of frames round robin to
    these CUDA devices (one worker thread and one library context per device, no ``torchrun``, no collective: the
    per-frame results come back to the one set of running statistics in frame order, so the numbers are those of a
    single-GPU run).  ``devices``: a list of device indices, an integer count, ``"all"`` / ``None`` for every visible
    device, ``[]`` / ``0`` to go back to one device.  Returns the list in use."""
    from . import _cabi

    if devices is None or devices == "all":
        devices = list(range(_cabi.device_count()))
    elif isinstance(devices, int):
        devices = list(range(devices))
    devices = [int(d) for d in devices]
    n = _cabi.device_count()
    if any(d < 0 or d >= max(n, 1) for d in devices):
        raise ValueError(f"device index out of range (have {n} CUDA devices)")
    _state["devices"] = devices if len(devices) > 0 else None
    return _state["devices"]


def devices() -> list[int] | None:
    """Devices of the single-process multi-GPU mode (``None``: the process default device only).  The environment
    variable ``MAICOS_B200_DEVICES`` (``all`` or a comma-separated list) selects them for the CLI; it is ignored in
    processes started by ``torchrun`` (one process per GPU)."""
    if _state["devices"] is None and not _state.get("devices_env_read"):
        _state["devices_env_read"] = True
        spec = os.environ.get("MAICOS_B200_DEVICES", "").strip()
        if spec and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
            use_devices("all" if spec.lower() == "all" else [int(x) for x in spec.split(",") if x.strip()])
    return _state["devices"]


def set_mode(mode: str) -> None:
    """Switch the sharding axis of an initialised process group (``"frames"`` or ``"qblocks"``)."""
    if mode not in ("frames", "qblocks"):
        raise ValueError("mode must be 'frames' or 'qblocks'")
    _state["mode"] = mode


@contextlib.contextmanager
def local_only():
    """Inside the block this rank behaves like a single-process run (no sharding, no collectives) -- for work that one
    rank does on its own while the group exists, e.g. a reference result to compare a sharded one with."""
    keep = (_state["rank"], _state["world"])
    _state["rank"], _state["world"] = 0, 1
    try:
        yield
    finally:
        _state["rank"], _state["world"] = keep


def shutdown() -> None:
    if _state["world"] > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()
    _nccl_buffers.clear()
    _state.update(initialised=False, rank=0, world=1, device=None, backend=None)


def rank() -> int:
    return _state["rank"]


def world() -> int:
    return _state["world"]


def mode() -> str:
    return _state["mode"]


def frame_sharded() -> bool:
    return _state["world"] > 1 and _state["mode"] == "frames"


def q_sharded() -> bool:
    return _state["world"] > 1 and _state["mode"] == "qblocks"


def q_block() -> tuple[int, int]:
    """``(block, n_blocks)`` argument of the C ABI for this rank."""
    return (_state["rank"], _state["world"]) if q_sharded() else (0, 1)


def allreduce_sum(buf: np.ndarray) -> np.ndarray:
    """In-place sum of a float64 host buffer over all ranks (one collective)."""
    if _state["world"] <= 1:
        return buf
    import torch
    import torch.distributed as dist

    assert buf.dtype == np.float64 and buf.flags.c_contiguous
    t = torch.from_numpy(buf.reshape(-1))
    if _state["backend"] == "nccl":
        # a persistent device tensor and a persistent pinned mirror per size: no allocation, no pageable staging per call
        key = t.numel()
        cached = _nccl_buffers.get(key)
        if cached is None:
            if len(_nccl_buffers) > 16:
                _nccl_buffers.clear()
            cached = _nccl_buffers[key] = (torch.empty(key, dtype=torch.float64, device=_state["device"]),
                                           torch.empty(key, dtype=torch.float64).pin_memory())
        d, h = cached
        h.copy_(t)
        d.copy_(h, non_blocking=True)
        dist.all_reduce(d, op=dist.ReduceOp.SUM)
        h.copy_(d, non_blocking=True)
        torch.cuda.current_stream(_state["device"]).synchronize()
        t.copy_(h)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return buf


def barrier() -> None:
    if _state["world"] > 1:
        import torch.distributed as dist

        dist.barrier()


def share_layout(layout):
    """First non-``None`` ``layout`` (a small picklable object) among the ranks.

    A rank whose share of the frames is empty (more ranks than frames) has no observables to derive the packed
    buffer of :func:`merge_welford` from; it adopts the layout of a rank that has.
    """
    if _state["world"] <= 1:
        return layout
    import torch.distributed as dist

    gathered = [None] * _state["world"]
    dist.all_gather_object(gathered, layout)
    for item in gathered:
        if item is not None:
            return item
    return None


def merge_welford(n_local: int, means: dict, sems: dict, sums: dict, extra: dict[str, np.ndarray]):
    """Merge per-rank running statistics with ONE allreduce.

    ``means / sems / sums`` hold, per observable, the rank-local running mean, standard error of the
    mean (``sqrt(var_pop / n)``) and sum.  ``extra`` arrays (timeseries, frames, times) are zero
    outside the rank's own slots and are summed.  Returns ``(n_total, means, sems, sums, extra)``.

    Layout of the packed buffer: ``[W slots of (n, mean..., M2...)] [sums...] [extra...]`` -- writing
    each rank's triple into its own slot turns the allreduce into an allgather of the triples.
    """
    W, r = _state["world"], _state["rank"]
    keys = sorted(means.keys())
    shapes = {k: np.shape(means[k]) for k in keys}
    sizes = {k: int(np.prod(shapes[k], dtype=np.int64)) for k in keys}
    total = sum(sizes.values())
    slot = 1 + 2 * total
    ex_keys = sorted(extra.keys())
    ex_total = sum(extra[k].size for k in ex_keys)
    buf = np.zeros(W * slot + total + ex_total, dtype=np.float64)
    base = r * slot
    buf[base] = n_local
    o = base + 1
    for k in keys:
        m = np.asarray(means[k], dtype=np.float64).reshape(-1)
        s = np.asarray(sems[k], dtype=np.float64).reshape(-1)
        buf[o:o + sizes[k]] = m
        buf[o + total:o + total + sizes[k]] = s * s * float(n_local) * float(n_local)  # M2 = n * var_pop
        o += sizes[k]
    o = W * slot
    for k in keys:
        buf[o:o + sizes[k]] = np.asarray(sums[k], dtype=np.float64).reshape(-1)
        o += sizes[k]
    for k in ex_keys:
        buf[o:o + extra[k].size] = np.asarray(extra[k], dtype=np.float64).reshape(-1)
        o += extra[k].size
    allreduce_sum(buf)

    ns = buf[0:W * slot:slot]
    n_tot = int(round(ns.sum()))
    mean_all = np.zeros(total)
    for w in range(W):
        mean_all += ns[w] * buf[w * slot + 1:w * slot + 1 + total]
    mean_all /= max(n_tot, 1)
    m2_all = np.zeros(total)
    for w in range(W):
        d = buf[w * slot + 1:w * slot + 1 + total] - mean_all
        m2_all += buf[w * slot + 1 + total:w * slot + 1 + 2 * total] + ns[w] * d * d
    sem_all = np.sqrt(np.maximum(m2_all, 0.0) / max(n_tot, 1) / max(n_tot, 1))
    out_means, out_sems, out_sums, out_extra = {}, {}, {}, {}
    o = 0
    so = W * slot
    for k in keys:
        sl = slice(o, o + sizes[k])
        out_means[k] = mean_all[sl].reshape(shapes[k]) if shapes[k] else float(mean_all[o])
        out_sems[k] = sem_all[sl].reshape(shapes[k]) if shapes[k] else float(sem_all[o])
        v = buf[so + o:so + o + sizes[k]]
        out_sums[k] = v.reshape(shapes[k]).copy() if shapes[k] else float(v[0])
        o += sizes[k]
    o = W * slot + total
    for k in ex_keys:
        out_extra[k] = buf[o:o + extra[k].size].reshape(extra[k].shape).copy()
        o += extra[k].size
    return n_tot, out_means, out_sems, out_sums, out_extra
