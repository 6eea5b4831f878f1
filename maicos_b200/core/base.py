"""Frame loop, running statistics and profile base of the drop-in analysis classes.

Keeps the protocol of ``maicos.core.base`` (``/root/reference/src/maicos/core/base.py``):

* ``AnalysisBase(atomgroup, unwrap, pack, refgroup, jitter, concfreq, wrap_compound)``,
  ``.run(start, stop, step, frames, verbose, progressbar_kwargs) -> self`` (``base.py:517-536``),
  hooks ``_prepare / _single_frame / _conclude / save``, attributes ``results, _obs, sums, means,
  sems, n_frames, frames, times, timeseries, corrtime, _index, _frame_index`` (``base.py:138-171``),
  the per-frame order *unwrap -> refgroup centring -> wrap -> jitter -> _single_frame -> statistics*
  (``base.py:417-499``), ``concfreq`` (``496-499``) and ``savetxt`` (``538-599``);
* ``ProfileBase`` (``base.py:722-857``): weighted + count histogram per frame, normalisation
  ``none / volume / number``.

What is different, by design: the loop *stages* frames and hands a whole batch to one C-ABI call
(``_stage_frame`` / ``_compute_staged``), so the GPU sees hundreds of frames per launch sequence;
the per-frame observables come back as small arrays and enter the statistics in frame order with
the reference's own recurrences (``lib/math.py:325-459``), so ``means / sems / sums`` are what a
serial run produces.  ``_single_frame()`` stays callable on its own (the reference's tests call
it directly) and is then a batch of one.  With ``maicos_b200.parallel`` initialised, frames are
sharded over ranks and merged by one allreduce at the end.
"""

from __future__ import annotations

import collections
import itertools
import logging
import time
import warnings
from collections.abc import Callable
from datetime import datetime

import numpy as np

from .. import __version__, parallel
from ..lib.math import center_cluster, new_mean, new_variance
from ..lib.util import (atomgroup_header, correlation_analysis, get_center, get_cli_input, get_module_input_str,
                        is_orthorhombic)

_COMPATIBLE = (np.ndarray, float, int, list, np.float32, np.float64, np.int32, np.int64)
_WRAP_COMPOUNDS = ["atoms", "group", "residues", "segments", "molecules", "fragments"]


class Results(dict):
    """Dictionary with attribute access (the ``MDAnalysis.analysis.base.Results`` behaviour used here)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as err:
            raise AttributeError(f"'Results' object has no attribute '{key}'") from err

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        try:
            del self[key]
        except KeyError as err:
            raise AttributeError(f"'Results' object has no attribute '{key}'") from err

    def copy(self):
        return Results(self)


def _is_arange(atomgroup, idx) -> bool:
    """``idx`` is ``arange(idx[0], idx[-1] + 1)`` (checked once per atom group object)."""
    cached = getattr(atomgroup, "_mb200_contiguous", None)
    if cached is None or cached[0] != len(idx):
        ok = bool(np.array_equal(idx, np.arange(int(idx[0]), int(idx[0]) + len(idx))))
        cached = (len(idx), ok)
        try:
            atomgroup._mb200_contiguous = cached
        except AttributeError:
            pass
    return cached[1]


def positions_into(atomgroup, out: np.ndarray) -> np.ndarray:
    """Copy the current positions of ``atomgroup`` into ``out`` (one pass, no temporary)."""
    try:
        src = atomgroup.universe.trajectory.ts.positions
        idx = atomgroup.indices
        n = len(idx)
        if n and int(idx[-1]) - int(idx[0]) == n - 1 and (getattr(atomgroup, "_all", False) or _is_arange(atomgroup, idx)):
            block = src[int(idx[0]):int(idx[0]) + n]  # a contiguous run of atoms: one block copy
            if n * 12 >= (1 << 22) and block.dtype == out.dtype:
                from .. import _cabi

                _cabi.host_copy(out, block)
            else:
                np.copyto(out, block)
        else:
            np.take(src, idx, axis=0, out=out)
    except (AttributeError, TypeError, ValueError):
        out[...] = atomgroup.positions
    return out


def gather_rows(rows: list) -> np.ndarray:
    """``[F, ...]`` array of staged rows ``(block, row)``: a view when they are consecutive rows of one block (the usual
    case: one staging buffer, or consecutive frames of a page-locked trajectory), else a stacked copy."""
    buf = rows[0][0]
    base = buf.ctypes.data
    if all(r[0].ctypes.data == base and r[1] == rows[0][1] + i for i, r in enumerate(rows)):
        return buf[rows[0][1]:rows[0][1] + len(rows)]
    return np.stack([b[r] for b, r in rows])


_COMPOUND_ATTR = {"residues": "resindices", "molecules": "molnums", "segments": "segindices",
                  "fragments": "fragindices"}


def compound_runs(atomgroup, compound: str):
    """``compound_ptr`` (int64 ``[n_c + 1]``) when the device prologue can own the compounds, else ``None``.

    Requirements: the in-memory shim (its wrap / unwrap semantics are what ``mb200_compound_prologue``
    restates bit for bit -- with a real MDAnalysis universe the host path keeps MDAnalysis' own),
    compounds are contiguous runs of the atom group in ascending id order (the order
    ``AtomGroup.center(compound=...)`` returns), complete with respect to the universe, at most 64 atoms.
    """
    from ..mdshim import Universe as ShimUniverse

    if compound not in _COMPOUND_ATTR or not isinstance(atomgroup.universe, ShimUniverse):
        return None
    cache = atomgroup.__dict__.setdefault("_mb200_compound_runs", {})  # topology does not change during a run
    if compound not in cache:
        cache[compound] = _compound_runs(atomgroup, compound)
    return cache[compound]


def _compound_runs(atomgroup, compound: str):
    ids = np.asarray(getattr(atomgroup, _COMPOUND_ATTR[compound]))
    if len(ids) == 0:
        return None
    starts = np.concatenate([[0], np.flatnonzero(np.diff(ids)) + 1])
    run_ids = ids[starts]
    if np.any(np.diff(run_ids) <= 0):
        return None
    ptr = np.concatenate([starts, [len(ids)]]).astype(np.int64)
    if np.max(np.diff(ptr)) > 64:
        return None
    all_ids = np.asarray(getattr(atomgroup.universe.atoms, _COMPOUND_ATTR[compound]))
    if np.count_nonzero(np.isin(all_ids, run_ids)) != len(ids):
        return None
    return ptr


def center_weights_for(atomgroup, bin_method: str):
    """Per-atom weights of ``get_center`` (``lib/util.py:668-679``): masses, |charges| or None."""
    if bin_method == "cog":
        return None
    if bin_method == "com":
        return np.ascontiguousarray(atomgroup.masses, dtype=np.float64)
    if bin_method == "coc":
        return np.ascontiguousarray(np.abs(atomgroup.charges), dtype=np.float64)
    raise ValueError(f"'{bin_method}' is an unknown binning method. Use 'cog', 'com' or 'coc'.")


class CompoundPrologue:
    """Device-side unwrap / wrap / centre / dipole weights of an atom group's compounds."""

    def __init__(self, atomgroup, compound: str, bin_method: str, make_whole: bool, wrap_center: bool,
                 weight_mode: int = 0, pdim: int = 0, uv_mode: int = 0, uv_dim: int = 0):
        from .. import _cabi

        self.ptr = compound_runs(atomgroup, compound)
        if self.ptr is None:
            raise ValueError("compounds are not eligible for the device prologue")
        self.n_atoms, self.n_c = atomgroup.n_atoms, len(self.ptr) - 1
        self.cw = center_weights_for(atomgroup, bin_method)
        has = atomgroup.universe._attrs
        self.masses = np.ascontiguousarray(atomgroup.masses, dtype=np.float64) if "masses" in has else None
        self.charges = np.ascontiguousarray(atomgroup.charges, dtype=np.float64) if "charges" in has else None
        if weight_mode and (self.masses is None or self.charges is None):
            raise ValueError("dipole weights need masses and charges")
        self.make_whole, self.wrap_center = bool(make_whole), bool(wrap_center)
        self.weight_mode, self.pdim = int(weight_mode), int(pdim)
        self.uv_mode, self.uv_dim = int(uv_mode), int(uv_dim)  # unit vector of the projection: Cartesian / cylinder / sphere
        self._dev = {}  # device -> (context, topology handle, centres buffer, weights buffer), created on first use
        self._on_device()  # fail early (ValueError / MB200Error) on the constructing thread's device

    def _on_device(self):
        """Per-device state of the calling thread's device: compound_ptr, centre weights, masses and charges go to a
        device once per analysis, not with every batch."""
        import ctypes

        from .. import _cabi

        ctx = _cabi.context()
        st = self._dev.get(ctx.device)
        if st is None:
            topo = ctypes.c_void_p()
            _cabi.check(_cabi.load().mb200_topology_create(
                ctx.handle, _cabi.ptr(self.ptr), self.n_c, self.n_atoms, _cabi.ptr(self.cw), _cabi.ptr(self.masses),
                _cabi.ptr(self.charges), ctypes.byref(topo)))
            st = self._dev[ctx.device] = (ctx, topo, _cabi.DeviceBuffer(ctx), _cabi.DeviceBuffer(ctx))
        return st

    def run(self, positions: np.ndarray, boxes: np.ndarray, window=None):
        """positions float32 [F, n_atoms, 3] (pinned), boxes float32 [F, 3] -> device pointers (centres, weights).

        ``window = (n_total, first)``: ``positions`` holds whole frames ``[F, n_total, 3]`` and the group is atoms
        ``first .. first + n_atoms`` of each (``AnalysisBase._frame_window``)."""
        from .. import _cabi

        ctx, topo, centers, weights = self._on_device()
        F = positions.shape[0]
        n_total, first = window if window is not None else (self.n_atoms, 0)
        if positions.shape[1] != n_total:
            raise ValueError(f"frames of {positions.shape[1]} atoms, expected {n_total}")
        c_ptr = centers.ensure(F * self.n_c * 24)
        n_w = {0: 0, 1: 1, 2: 1, 3: 1, 4: 3}[self.weight_mode]
        w_ptr = weights.ensure(F * self.n_c * 8 * n_w) if n_w else None
        boxes = np.ascontiguousarray(boxes, dtype=np.float32)
        _cabi.check(_cabi.load().mb200_compound_prologue_window(
            ctx.handle, _cabi.ptr(positions), 0, F, int(n_total), int(first), _cabi.ptr(boxes), topo, int(self.make_whole),
            int(self.wrap_center), self.weight_mode, self.pdim, self.uv_mode, self.uv_dim, c_ptr, w_ptr))
        return c_ptr, w_ptr

    def close(self) -> None:
        from .. import _cabi

        dev, self._dev = getattr(self, "_dev", {}), {}
        for ctx, topo, centers, weights in dev.values():
            centers.free()
            weights.free()
            if topo:
                _cabi.load().mb200_topology_destroy(ctx.handle, topo)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VelocityWeights:
    """Device-side ``velocity_weights`` / ``temperature_weights`` (``lib/weights.py:101-127, 195-225``) from the raw
    float32 velocities of a batch: mode 0 ``v[:, vdim]`` per atom, 1 mass-weighted compound mean, 2 kinetic temperature."""

    def __init__(self, atomgroup, mode: int, vdim: int, prologue: CompoundPrologue | None = None):
        self.mode, self.vdim = int(mode), int(vdim)
        self.n_atoms = atomgroup.n_atoms
        self.prologue = prologue          # compounds: the prologue's topology (compound_ptr + masses)
        self.n_c = prologue.n_c if prologue is not None else self.n_atoms
        self.masses = np.ascontiguousarray(atomgroup.masses, dtype=np.float64) if (prologue is None and mode == 2) else None
        self._dev = {}  # device -> (context, own topology handle or None, weights buffer)
        self._on_device()

    def _on_device(self):
        import ctypes

        from .. import _cabi

        ctx = _cabi.context()
        st = self._dev.get(ctx.device)
        if st is None:
            own = None
            if self.prologue is None:     # every atom its own compound
                own = ctypes.c_void_p()
                _cabi.check(_cabi.load().mb200_topology_create(ctx.handle, None, self.n_atoms, self.n_atoms, None,
                                                               _cabi.ptr(self.masses), None, ctypes.byref(own)))
            st = self._dev[ctx.device] = (ctx, own, _cabi.DeviceBuffer(ctx))
        return st

    def run(self, velocities: np.ndarray) -> int:
        """velocities float32 [F, n_atoms, 3] (pinned) -> device pointer of the weights float64 [F, n_c]."""
        from .. import _cabi

        ctx, own, weights = self._on_device()
        topo = own if own is not None else self.prologue._on_device()[1]
        F = velocities.shape[0]
        w_ptr = weights.ensure(F * self.n_c * 8)
        _cabi.check(_cabi.load().mb200_velocity_weights(ctx.handle, _cabi.ptr(velocities), 0, F, topo, self.mode, self.vdim, w_ptr))
        return w_ptr

    def close(self) -> None:
        from .. import _cabi

        dev, self._dev = getattr(self, "_dev", {}), {}
        for ctx, own, weights in dev.values():
            weights.free()
            if own:
                _cabi.load().mb200_topology_destroy(ctx.handle, own)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


#: numbers the runs of AnalysisCollection (tags of shared frame announcements, ``mb200_prefetch_frames_shared``)
_collection_runs = itertools.count(1)


class _Runner:
    """Shared ``run`` machinery of :class:`AnalysisBase` and :class:`AnalysisCollection`."""

    def _run(self, analysis_instances, start=None, stop=None, step=None, frames=None, verbose=None,
             progressbar_kwargs=None):
        self._run_locals = locals()
        logging.getLogger(__name__).setLevel(logging.INFO if verbose else logging.WARNING)
        if frames is not None and not all(opt is None for opt in [start, stop, step]):
            raise ValueError("start/stop/step cannot be combined with frames")
        share_run = next(_collection_runs) if len(analysis_instances) > 1 else 0
        for a in analysis_instances:
            a._run_locals = self._run_locals
            a._share_run = share_run
            a._setup_frames(a._trajectory, start=start, stop=stop, step=step, frames=frames)
        for a in analysis_instances:
            a._call_prepare()
        if len(analysis_instances) > 1:
            # members of a collection stage the same number of frames per device call, so that those reading the frames
            # where the trajectory holds them announce identical blocks and share one host -> device copy
            batched = [a for a in analysis_instances if getattr(type(a)._single_frame, "_from_batch", False)]
            for a in batched:
                a._batch_size = min(b._batch_size for b in batched)

        if progressbar_kwargs:
            logging.getLogger(__name__).info("progressbar_kwargs are accepted for API compatibility and ignored "
                                             "(frames are staged in batches; there is no per-frame progress bar).")
        first = analysis_instances[0]
        local = first._local_frame_slots()
        trajectory = first._trajectory
        try:
            for slot in local:
                ts = trajectory[int(first.frames_selected[slot])]
                # every analysis starts from the untouched frame; frames of the in-memory shim are copy-on-write views
                # of the trajectory, so "untouched" is the view itself and nothing is copied unless an analysis transforms
                shared = len(analysis_instances) > 1 and not ts.positions.flags.writeable
                ts_original = ts.positions if shared else (ts.copy().positions if len(analysis_instances) > 1 else None)
                for a in analysis_instances:
                    if shared:
                        ts.positions = ts_original
                    elif ts_original is not None:
                        ts.positions[...] = ts_original
                    a._enqueue_frame(ts, int(slot))
            for a in analysis_instances:
                a._flush_staged()
                a._merge_ranks()
            for a in analysis_instances:
                a._call_conclude()
        finally:
            for a in analysis_instances:
                a._release_run_resources()
        return self


class AnalysisBase(_Runner):
    """Base class of multi-frame analyses (protocol of ``maicos.core.AnalysisBase``).

    Parameters
    ----------
    atomgroup : AtomGroup
        atoms to analyse (MDAnalysis AtomGroup or the in-memory shim).
    unwrap : bool
        make compounds whole before the analysis (``atoms.unwrap(compound=wrap_compound)``).
    pack : bool
        wrap compounds into the primary cell.
    refgroup : AtomGroup or None
        centre the cell on this group's (periodic) centre of mass every frame.
    jitter : float
        magnitude of a uniform random displacement added to every coordinate.
    concfreq : int
        call ``_conclude`` (and ``save``) every ``concfreq`` frames; 0 = only at the end.
    wrap_compound : str
        compound kept whole when wrapping.
    """

    #: frames staged per device call by classes that implement ``_stage_frame``/``_compute_staged``
    _batch_size = 1
    #: True when ``pack`` with ``wrap_compound="atoms"`` is applied by the kernel instead of the host
    _device_pack = False

    def __init__(self, atomgroup, unwrap: bool, pack: bool, refgroup, jitter: float, concfreq: int,
                 wrap_compound: str) -> None:
        self.atomgroup = atomgroup
        if self.atomgroup.n_atoms == 0:
            raise ValueError("The provided `atomgroup` does not contain any atoms.")
        self._universe = atomgroup.universe
        self._trajectory = self._universe.trajectory
        self.refgroup = refgroup
        self.unwrap = unwrap
        self.pack = pack
        self.jitter = jitter
        self.concfreq = concfreq
        if wrap_compound not in _WRAP_COMPOUNDS:
            raise ValueError(
                "Unrecognized `wrap_compound` definition "
                f"{wrap_compound}: \nPlease use "
                "one of 'atoms', 'group', 'residues', "
                "'segments', 'molecules', or 'fragments'."
            )
        self.wrap_compound = wrap_compound
        if self.unwrap and self._universe.dimensions is None:
            raise ValueError("Universe does not have `dimensions` and can't be unwrapped!")
        if self.pack and self._universe.dimensions is None:
            raise ValueError("Universe does not have `dimensions` and can't be packed!")
        if self.unwrap and self.wrap_compound == "atoms":
            logging.warning(
                "Unwrapping in combination with the `wrap_compound='atoms` is superfluous. "
                "`unwrap` will be set to `False`."
            )
            self.unwrap = False
        if self.refgroup is not None:
            if self.refgroup.n_atoms == 0:
                raise ValueError("The provided `refgroup` does not contain any atoms.")
            if not self.pack:
                raise ValueError(
                    "Disabling `pack` with a `refgroup` is not allowed. Shifting "
                    "atoms probably outside of the primary cell withput packing them "
                    "back may lead to sever problems during the analysis!"
                )
        self.module_has_save = callable(getattr(self.__class__, "save", None))
        self.results = Results()
        self._obs = Results()
        self._staged: list = []
        self._pools = None            # device -> single-worker executor of the device calls
        self._inflight_q = collections.deque()
        self._batch_device = None     # device of the batch being announced / submitted
        self._batch_tag = 0           # content tag of the batch being announced (shared frames of a collection)
        self._pending = {}            # device -> (staged batch, future of its submission) waiting to be collected
        self._stage_bufs = {}         # (tag, parity, rows, shape, dtype) -> pinned staging array of this run
        self._window_state = None     # see ``_frame_window``
        self._share_run = 0           # number of the collection run this instance is part of, 0 = on its own
        self._stage_parity = 0
        self.timings = {"stage_s": 0.0, "wait_device_s": 0.0, "device_call_s": 0.0}

    # ---- cell ------------------------------------------------------------------------------
    @property
    def box_lengths(self) -> np.ndarray:
        """Lengths of the simulation cell vectors (float64)."""
        return self._universe.dimensions[:3].astype(np.float64)

    @property
    def box_center(self) -> np.ndarray:
        """Centre of the simulation cell."""
        return self.box_lengths / 2

    # ---- frame selection -------------------------------------------------------------------
    def _setup_frames(self, trajectory, start=None, stop=None, step=None, frames=None) -> None:
        n_total = len(trajectory)
        if frames is not None:
            sel = np.asarray(frames)
            if sel.dtype == bool:
                sel = np.nonzero(sel)[0]
            sel = np.asarray(sel, dtype=np.int64)
        else:
            sel = np.arange(n_total, dtype=np.int64)[slice(start, stop, step)]
        self.start, self.stop, self.step = start, stop, step
        self.frames_selected = sel
        self.n_frames = len(sel)
        self.frames = np.zeros(self.n_frames, dtype=int)
        self.times = np.zeros(self.n_frames)
        self._sliced_trajectory = trajectory[sel]

    def _local_frame_slots(self) -> np.ndarray:
        """Slots (positions in the selected-frame list) this rank analyses."""
        slots = np.arange(self.n_frames)
        if parallel.frame_sharded():
            # fewer frames than ranks: the surplus ranks analyse nothing and join the final merge with n = 0
            # (parallel.init(mode='qblocks') is the mode that spreads a single huge frame)
            slots = slots[parallel.rank()::parallel.world()]
        return slots

    # ---- hooks ---------------------------------------------------------------------------------
    def _prepare(self) -> None:
        """Set things up before the frame loop."""

    def _single_frame(self):
        """Analyse the current frame: fill ``self._obs`` and return the correlation scalar."""
        raise NotImplementedError("Only implemented in child classes")

    def _single_frame_from_batch(self):
        """``_single_frame`` of the batched classes: a batch of one frame through the device call."""
        (obs, value), = self._compute_staged([self._stage_frame()])
        self._obs.update(obs)
        return value

    _single_frame_from_batch._from_batch = True

    def _stage_frame(self):
        """Capture what the device needs from the current (already transformed) frame."""
        return None

    def _slot_prefix(self) -> str:
        return f"{type(self).__name__}@{id(self):x}:"

    def _stage_positions(self, tag: str = "pos") -> tuple[np.ndarray, int]:
        """``(block, row)`` holding the float32 positions of ``self.atomgroup`` in the current frame.

        Normally a row of the pinned batch buffer filled with one host copy.  When the trajectory itself lives in
        page-locked memory (``trajectory.pinned_row``), the atom group is the whole system and nothing transformed the
        frame, the row IS the trajectory's: no host copy, the device reads the frame where the reader put it."""
        ag = self.atomgroup
        pinned_row = getattr(self._trajectory, "pinned_row", None)
        window = self._frame_window()
        if pinned_row is not None and (getattr(ag, "_all", False) or window is not None):
            ts = self._universe.trajectory.ts
            where = pinned_row(ts.frame)
            if where is not None and where[0][where[1]].ctypes.data == ts.positions.ctypes.data:
                return where
        if window is not None:  # a frame that is not where the trajectory holds it: the whole frame is staged all the same
            buf, row = self._staging_rows((window[0], 3), np.float32, tag)
            np.copyto(buf[row], self._universe.trajectory.ts.positions)
            return buf, row
        buf, row = self._staging_rows((ag.n_atoms, 3), np.float32, tag)
        positions_into(ag, buf[row])
        return buf, row

    def _gather_staged(self, rows: list, tag: str = "gather") -> np.ndarray:
        """``[F, ...]`` array of staged rows for the device call: the rows themselves when they are consecutive (a view,
        no copy), else gathered into a page-locked buffer (strided frame selections of a pinned trajectory)."""
        from .. import _cabi

        buf, base = rows[0][0], rows[0][0].ctypes.data
        if all(r[0].ctypes.data == base and r[1] == rows[0][1] + i for i, r in enumerate(rows)):
            return buf[rows[0][1]:rows[0][1] + len(rows)]
        # three buffers in turn: with deferred batches (mb200_saxs_frames_submit) the copy of batch k out of its buffer may
        # still be queued on the device when batch k + 1 is gathered
        self._gather_turn = (getattr(self, "_gather_turn", -1) + 1) % 3
        out = _cabi.pinned_empty((len(rows), *buf.shape[1:]), buf.dtype, slot=f"{self._slot_prefix()}{tag}:{self._gather_turn}")
        for i, (b, r) in enumerate(rows):
            _cabi.host_copy(out[i], b[r])
        return out

    def _announce_rows(self, rows: list) -> None:
        """Announce consecutive staged rows to the library (``mb200_prefetch_frames``): their host -> device copy then
        runs on the copy stream while the device call before this batch is still computing."""
        from .. import _cabi

        if not rows:
            return
        buf, base = rows[0][0], rows[0][0].ctypes.data
        if not all(r[0].ctypes.data == base and r[1] == rows[0][1] + i for i, r in enumerate(rows)):
            return  # not one run of rows: the call gathers and copies by itself
        block = buf[rows[0][1]:rows[0][1] + len(rows)]
        ctx = _cabi.context(self._batch_device)  # the device this batch is about to be submitted to
        # in a collection the analyses that read the frames where the trajectory holds them announce the same block with
        # the same tag (run, first frame of the batch): one host -> device copy serves them all
        _cabi.check(_cabi.load().mb200_prefetch_frames_shared(ctx.handle, block.ctypes.data, block.nbytes, self._batch_tag))

    def _staging_rows(self, shape, dtype, tag: str) -> tuple[np.ndarray, int]:
        """``(pinned batch buffer [batch, *shape], row)`` for the frame being staged.

        Two buffers alternate per batch: one is read by the in-flight device call while the host
        fills the other, so frames go trajectory -> pinned memory -> HBM with one host copy.
        """
        rows = max(1, self._batch_size)
        key = (tag, self._stage_parity, rows, shape, dtype)
        buf = self._stage_bufs.get(key)
        if buf is None:  # one look-up per (buffer, run) instead of one per frame
            from .. import _cabi

            buf = _cabi.pinned_empty((rows, *shape), dtype, slot=f"{self._slot_prefix()}{tag}:{self._stage_parity}")
            for k in [k for k in self._stage_bufs if k[:2] == key[:2]]:  # the slot's previous array is invalid now
                del self._stage_bufs[k]
            self._stage_bufs[key] = buf
        return buf, len(self._staged) % rows

    def _compute_staged(self, staged: list) -> list:
        """Turn staged frames into ``[(obs dict, correlation scalar), ...]`` with one device call."""
        raise NotImplementedError

    def _announce_staged(self, staged: list) -> None:
        """Hook called when a batch is complete, before it is queued: classes whose device call reads the pinned
        staging buffer directly announce it (``mb200_prefetch_frames``) so that its host -> device copy overlaps with
        the device call still running."""

    def _conclude(self) -> None:
        """Finalise ``results`` from ``sums / means / sems``."""

    # ---- driver --------------------------------------------------------------------------------
    def _call_prepare(self) -> None:
        if self.refgroup is not None:
            if not hasattr(self.refgroup, "masses") or np.sum(self.refgroup.masses) == 0:
                logging.warning("No masses available in refgroup, falling back to center of geometry")
                self.ref_weights = np.ones(self.refgroup.n_atoms)
            else:
                self.ref_weights = self.refgroup.masses
        self._window_state = False  # decided on first use (``_frame_window``), after ``_prepare``
        self._prepare()
        # a streaming reader that decodes into a ring of page-locked frame buffers (mdshim.TRRTrajectory(pinned_ring=k)):
        # rows handed to the device must not be overwritten while a batch is in flight
        ring = int(getattr(self._trajectory, "ring_len", 0) or 0)
        if ring:
            self._batch_size = int(max(1, min(self._batch_size, ring // (len(self._run_devices()) + 2 + int(self._defers_batches())))))
        if logging.getLogger().isEnabledFor(logging.INFO):  # the headers scan every atom: only when shown
            if self.refgroup is not None:
                logging.info(
                    f"Coordinates are relative to the center of mass of reference atomgroup "
                    f"{atomgroup_header(self.refgroup)}."
                )
            else:
                logging.info("Coordinates are relative to the center of the simulation box.")
            logging.info(f"Considered atomgroup {atomgroup_header(self.atomgroup)}.")
        if hasattr(self, "n_bins"):
            logging.info(f"Using {self.n_bins} bins.")
        self.timeseries = np.zeros(self.n_frames)
        self._n_local = 0
        self._n_flushed = 0
        self._staged = []
        self._inflight_q = collections.deque()
        self._pending = {}
        self._stage_bufs = {}
        #: host-side time split of the last run: staging, waiting for the device, inside the C-ABI call
        self.timings = {"stage_s": 0.0, "wait_device_s": 0.0, "device_call_s": 0.0}
        for name in ("sums", "means", "sems"):
            if hasattr(self, name):
                delattr(self, name)
        logging.info(f"Analysing {self.n_frames} trajectory frames.")
        logging.debug(f"Module input: {get_module_input_str(self)}")

    def _host_transforms(self) -> bool:
        """True when ``_transform_frame`` may rewrite the frame on the host (same conditions, asked once per run)."""
        device_compounds = getattr(self, "_prologue", None) is not None
        if (self.unwrap and not device_compounds) or self.refgroup is not None or self.jitter != 0.0:
            return True
        return bool(self.pack and self._universe.dimensions is not None and not device_compounds and not (
            self._device_pack and self.wrap_compound == "atoms" and is_orthorhombic(self._universe.dimensions)))

    def _frame_window(self):
        """``(n_total, first)`` when the device calls of this run read the atom group out of WHOLE frames, else ``None``.

        A group that is not the whole universe normally has its atoms gathered into the staging buffer frame by frame
        (one host pass over the selection).  When the trajectory lives in page-locked memory and nothing transforms the
        frames on the host, classes that can address their atoms inside a whole frame (``_window_capable``: ``"range"`` --
        a contiguous selection, as the water of a slab system; ``"indices"`` -- any selection, Saxs gathers by index on
        the device anyway, the profiles of single atoms copy them out of the frame in front of the histogram) stage the
        trajectory's own rows instead: no host copy, and in a collection the frames reach the device once for all
        members.  ``first`` is ``None`` for a selection that is not one contiguous range."""
        if self._window_state is False:
            self._window_state = None
            mode = getattr(self, "_window_capable", None)
            mode = mode() if callable(mode) else mode
            ag = self.atomgroup
            if (mode and not getattr(ag, "_all", False) and getattr(self._trajectory, "has_pinned_frames", False)
                    and hasattr(ag, "indices") and not self._host_transforms()):
                idx = np.asarray(ag.indices)
                n_total = int(self._universe.atoms.n_atoms)
                if len(idx) and int(idx[-1]) - int(idx[0]) + 1 == len(idx) and (len(idx) < 2 or bool(np.all(np.diff(idx) == 1))):
                    self._window_state = (n_total, int(idx[0]))
                elif mode == "indices":
                    self._window_state = (n_total, None)
        return self._window_state

    def _transform_frame(self, ts) -> None:
        """unwrap -> refgroup centring -> wrap -> jitter, in the reference's order (base.py:436-451)."""
        device_compounds = getattr(self, "_prologue", None) is not None
        if self.unwrap and not device_compounds:
            self._universe.atoms.unwrap(compound=self.wrap_compound)
        if self.refgroup is not None:
            com_refgroup = center_cluster(self.refgroup, self.ref_weights)
            self._universe.atoms.translate(self.box_center - com_refgroup)
        if self.pack and self._universe.dimensions is not None and not device_compounds and not (
            self._device_pack and self.wrap_compound == "atoms" and is_orthorhombic(self._universe.dimensions)
        ):
            self._universe.atoms.wrap(compound=self.wrap_compound)
        if self.jitter != 0.0:
            pos = ts.writable_positions() if hasattr(ts, "writable_positions") else ts.positions
            pos += np.random.random(size=(len(pos), 3)) * self.jitter

    def _begin_frame(self, ts, slot: int) -> None:
        self._frame_index = slot
        self._ts = ts
        self.frames[slot] = ts.frame
        self.times[slot] = ts.time
        self._transform_frame(ts)

    def _enqueue_frame(self, ts, slot: int) -> None:
        self._begin_frame(ts, slot)
        # batched unless a subclass replaced `_single_frame` with its own per-frame code
        if getattr(type(self)._single_frame, "_from_batch", False):
            t0 = time.perf_counter()
            self._staged.append((slot, self._stage_frame()))
            self.timings["stage_s"] += time.perf_counter() - t0
            if len(self._staged) >= self._flush_threshold():
                self._flush_staged(wait=False)
        else:
            self._obs = Results()
            value = self._single_frame()
            self._accumulate(slot, value, self._obs)

    def _flush_threshold(self) -> int:
        """Frames per device call: the first two batches of a run are a quarter and a half of ``_batch_size`` so
        that the device starts after a few staged frames instead of a whole batch (shorter pipeline fill)."""
        b = max(1, self._batch_size)
        k = getattr(self, "_n_flushed", 2)
        return b if k >= 2 else max(1, b >> (2 - k))

    def _run_devices(self) -> list:
        """CUDA devices this run deals its batches to: ``parallel.use_devices`` or ``[None]`` (the default device)."""
        devs = parallel.devices()
        return list(devs) if devs else [None]

    def _flush_staged(self, wait: bool = True) -> None:
        """Hand the staged frames to the device.

        The device call runs on a worker thread (the C ABI releases the GIL) so that the host can stage the
        next batch meanwhile.  A new batch is queued *behind* the one in flight before the host waits for the
        latter, so the worker goes from one device call straight into the next; batches complete in order,
        and their statistics are accumulated on the calling thread, in frame order, while the device works.
        With ``parallel.use_devices`` batches go round robin to one worker per device; results still enter the
        statistics in frame order.
        """
        if self._staged:
            staged, self._staged = self._staged, []
            devs = self._run_devices()
            k = self._n_flushed = getattr(self, "_n_flushed", 0) + 1
            device = devs[(k - 1) % len(devs)]
            if self._pools is None:
                self._pools = {}
            pool = self._pools.get(device)
            if pool is None:
                from concurrent.futures import ThreadPoolExecutor

                from .. import _cabi

                pool = self._pools[device] = ThreadPoolExecutor(
                    max_workers=1, thread_name_prefix=f"mb200-batch-{device}", initializer=_cabi.set_thread_device,
                    initargs=(device,))
            self._batch_device = device
            self._batch_tag = ((self._share_run << 32) | (int(staged[0][0]) + 1)) if self._share_run else 0
            self._announce_staged([st for _, st in staged])
            deferred = self._defers_batches()
            if deferred:
                # submit / collect (mb200_saxs_frames_submit): the worker queues this batch behind the one that is running and
                # only then waits for the older one -- the device never idles between two batches
                ticket = pool.submit(self._submit_batch, staged)
                older = self._pending.get(device)
                if older is not None:
                    self._inflight_q.append(pool.submit(self._collect_batch, *older))
                self._pending[device] = (staged, ticket)
            else:
                self._inflight_q.append(pool.submit(self._device_batch, staged))
            # the next batch is staged into the next pinned buffer: one more buffer than batches in flight
            self._stage_parity = (self._stage_parity + 1) % (len(devs) + 1 + int(deferred))
            while len(self._inflight_q) > len(devs):
                self._finish_batch(self._inflight_q.popleft())
        if wait:
            self._drain()

    def _drain(self) -> None:
        # batches still waiting on the device: collected in the order they were submitted (= frame order)
        for device, older in sorted(self._pending.items(), key=lambda kv: kv[1][0][0][0]):
            self._inflight_q.append(self._pools[device].submit(self._collect_batch, *older))
        self._pending = {}
        while self._inflight_q:
            self._finish_batch(self._inflight_q.popleft())

    def _defers_batches(self) -> bool:
        """True when the class splits its device call into ``_submit_staged`` / ``_collect_staged``."""
        return False

    def _device_batch(self, staged: list):
        t0 = time.perf_counter()
        out = self._compute_staged([s for _, s in staged])
        self.timings["device_call_s"] += time.perf_counter() - t0
        return staged, out

    def _submit_batch(self, staged: list):
        t0 = time.perf_counter()
        state = self._submit_staged([s for _, s in staged])
        self.timings["device_call_s"] += time.perf_counter() - t0
        return state

    def _collect_batch(self, staged: list, ticket):
        state = ticket.result()  # re-raises what the submission raised
        t0 = time.perf_counter()
        out = self._collect_staged([s for _, s in staged], state)
        self.timings["device_call_s"] += time.perf_counter() - t0
        return staged, out

    def _finish_batch(self, fut) -> None:
        if fut is None:
            return
        t0 = time.perf_counter()
        staged, out = fut.result()  # re-raises worker exceptions
        self.timings["wait_device_s"] += time.perf_counter() - t0
        start = 0
        if not hasattr(self, "means") and staged:  # the first frame of the run initialises the statistics
            (slot, _), (obs, value) = staged[0], out[0]
            obs = Results(obs)
            self._accumulate(slot, value, obs)
            self._frame_index, self._obs = slot, obs
            start = 1
        if start < len(staged) and not self._accumulate_batch(staged[start:], out[start:]):
            for (slot, _), (obs, value) in zip(staged[start:], out[start:]):
                obs = Results(obs)
                self._accumulate(slot, value, obs)
                self._frame_index, self._obs = slot, obs

    def _accumulate_batch(self, staged: list, out: list) -> bool:
        """The statistics recurrences of ``_accumulate`` for a whole batch in one host call per observable
        (``mb200_running_stats``: the same operations, frame after frame -- bitwise the per-frame result).  Returns False
        (nothing done) when the batch needs the per-frame path: ``concfreq`` checkpoints, observables that are not float64 /
        int64 arrays or float scalars, or shapes that changed."""
        from .. import _cabi

        if self.concfreq or not out:
            return False
        keys = list(self.means.keys())
        plan = []
        for key in keys:
            first = out[0][0].get(key)
            m = self.means[key]
            if first is None or any(key not in o for o, _ in out):
                return False
            if isinstance(first, np.ndarray) and isinstance(m, np.ndarray) and first.dtype in (np.float64, np.int64) \
                    and first.shape == m.shape and m.dtype in (np.float64, np.int64) \
                    and np.asarray(self.sums[key]).dtype == first.dtype:
                if any(not isinstance(o[key], np.ndarray) or o[key].shape != first.shape or o[key].dtype != first.dtype
                       for o, _ in out):
                    return False
                plan.append((key, "array", first.dtype))
            elif type(first) in (float, np.float64) and type(m) in (float, np.float64):
                if any(type(o[key]) not in (float, np.float64) for o, _ in out):
                    return False
                plan.append((key, "scalar", np.dtype(np.float64)))
            else:
                return False
        if any(len(o) != len(keys) for o, _ in out):
            return False
        lib, n0, F = _cabi.load(), self._n_local, len(out)
        for key, kind, dtype in plan:
            x = np.ascontiguousarray(np.stack([o[key] for o, _ in out]) if kind == "array"
                                     else np.array([o[key] for o, _ in out], dtype=np.float64))
            size = int(x.size // F)
            mean = np.ascontiguousarray(self.means[key], dtype=np.float64).reshape(-1).copy()
            sem = np.ascontiguousarray(self.sems[key], dtype=np.float64).reshape(-1).copy()
            total = np.ascontiguousarray(self.sums[key]).reshape(-1).copy()
            is_int = dtype == np.int64
            _cabi.check(lib.mb200_running_stats(
                n0, F, size, int(kind == "scalar"), None if is_int else _cabi.ptr(x), _cabi.ptr(x) if is_int else None,
                _cabi.ptr(mean), _cabi.ptr(sem), None if is_int else _cabi.ptr(total), _cabi.ptr(total) if is_int else None))
            if kind == "array":
                shape = np.shape(self.means[key])
                self.means[key], self.sems[key], self.sums[key] = mean.reshape(shape), sem.reshape(shape), total.reshape(shape)
            else:
                self.means[key], self.sems[key], self.sums[key] = np.float64(mean[0]), np.float64(sem[0]), np.float64(total[0])
        for (slot, _), (obs, value) in zip(staged, out):
            self.timeseries[slot] = np.nan if value is None else value
        self._n_local += F
        self._index = self._n_local
        self._frame_index, self._obs = staged[-1][0], Results(out[-1][0])
        return True

    def _call_single_frame(self, ts, current_frame_index) -> None:
        """Serial entry point with the reference's signature (one frame, statistics updated)."""
        self._begin_frame(ts, current_frame_index)
        self._obs = Results()
        value = self._single_frame()
        self._accumulate(current_frame_index, value, self._obs)

    def _accumulate(self, slot: int, value, obs) -> None:
        """Running mean / standard error / sum of every observable (base.py:453-499)."""
        self.timeseries[slot] = np.nan if value is None else value
        self._n_local += 1
        self._index = n = self._n_local
        if n == 1 or not hasattr(self, "means"):
            for key in obs:
                if type(obs[key]) is list:
                    obs[key] = np.array(obs[key])
                if not isinstance(obs[key], _COMPATIBLE):
                    raise TypeError(f"Obervable {key} has uncompatible type.")
            self.sums = obs.copy()
            self.means = obs.copy()
            self.sems = Results({key: np.zeros(np.shape(obs[key])) for key in obs})
        else:
            for key in obs:
                x = obs[key]
                if type(x) is list:
                    x = obs[key] = np.array(x)
                old_mean = self.means[key]
                old_var = self.sems[key] ** 2 * (n - 1)
                self.means[key] = new_mean(old_mean, x, n)
                self.sems[key] = np.sqrt(new_variance(old_var, old_mean, self.means[key], x, n) / n)
                self.sums[key] = self.sums[key] + x
        if self.concfreq and parallel.frame_sharded():
            # frames are dealt round robin, so after n local frames the ranks together hold the first n * W frames of the
            # selection: every `concfreq // W` local frames (while all ranks still have frames) the partial statistics
            # are merged and concluded -- what a serial run shows at the same global frame count (base.py:496-499)
            every = max(1, self.concfreq // parallel.world())
            if n % every == 0 and n <= (self.n_frames // parallel.world()) // every * every:
                self._sharded_checkpoint()
        elif self.concfreq and n % self.concfreq == 0 and slot > 0:
            self._conclude()
            if self.module_has_save:
                self.save()

    def _sharded_checkpoint(self) -> None:
        """``concfreq`` under frame sharding: Welford merge of the ranks' partial statistics (one allreduce),
        ``_conclude`` + ``save`` on the merged state, then back to the rank-local state."""
        keep = {k: getattr(self, k) for k in ("means", "sems", "sums", "timeseries", "frames", "times", "_index")}
        keep = {k: (Results({kk: np.copy(vv) for kk, vv in v.items()}) if isinstance(v, dict) else np.copy(v))
                for k, v in keep.items()}
        self._merge_ranks()
        self._conclude()
        if self.module_has_save and parallel.rank() == 0:
            self.save()
        for k, v in keep.items():
            setattr(self, k, v if isinstance(v, dict) or np.ndim(v) else v.item())

    def _merge_ranks(self) -> None:
        """Frame-sharded runs: merge the per-rank statistics with one allreduce."""
        if not parallel.frame_sharded():
            return
        mine = np.zeros(self.n_frames, dtype=bool)
        mine[self._local_frame_slots()[:self._n_local]] = True  # the frames analysed so far (all of them at the end)
        # ranks without frames (more ranks than frames) take the observables' shapes from a rank that has some and
        # enter the merge with n = 0
        have = hasattr(self, "sums")
        layout = parallel.share_layout(
            {k: (np.shape(v), np.asarray(v).dtype.str) for k, v in self.sums.items()} if have else None)
        if layout is None:
            raise ValueError("No frame was analysed on any rank.")
        if not have:
            zeros = {k: np.zeros(shape, dtype=np.dtype(dt)) if shape else np.dtype(dt).type(0) for k, (shape, dt) in layout.items()}
            self.sums = Results({k: np.copy(v) for k, v in zeros.items()})
            self.means = Results({k: np.asarray(v, dtype=np.float64).copy() for k, v in zeros.items()})
            self.sems = Results({k: np.asarray(v, dtype=np.float64).copy() for k, v in zeros.items()})
        dtypes = {k: np.asarray(v).dtype for k, v in self.sums.items()}
        extra = {
            "timeseries": np.where(mine, self.timeseries, 0.0),
            "frames": np.where(mine, self.frames, 0).astype(np.float64),
            "times": np.where(mine, self.times, 0.0),
        }
        n, means, sems, sums, extra = parallel.merge_welford(self._n_local, self.means, self.sems, self.sums, extra)
        for k in sums:  # integer observables (bin counts) stay integers: float64 holds them exactly
            if np.issubdtype(dtypes[k], np.integer):
                sums[k] = np.rint(sums[k]).astype(dtypes[k])
        self.means, self.sems, self.sums = Results(means), Results(sems), Results(sums)
        self.timeseries = extra["timeseries"]
        self.frames = np.rint(extra["frames"]).astype(int)
        self.times = extra["times"]
        self._index = n

    def _call_conclude(self) -> None:
        with warnings.catch_warnings():
            if parallel.rank() != 0:
                warnings.simplefilter("ignore")
            self.corrtime = correlation_analysis(self.timeseries)
        self._conclude()
        if self.concfreq and self.module_has_save and parallel.rank() == 0:
            self.save()

    def _release_run_resources(self) -> None:
        """End of a run, also when a frame raised: wait for the batch in flight, stop the worker thread and hand the
        pinned staging blocks (keyed by ``id(self)``) back, so that no later instance can be given a slot that a device
        call still reads."""
        for device, older in list(getattr(self, "_pending", {}).items()):  # submitted, never collected: free the ticket
            try:
                self._pools[device].submit(self._collect_batch, *older).result()
            except Exception:
                pass
        self._pending = {}
        while self._inflight_q:
            try:
                self._inflight_q.popleft().result()
            except Exception:  # the original exception is already propagating
                pass
        self._staged = []
        pools, self._pools = self._pools, None
        for pool in (pools or {}).values():
            pool.shutdown(wait=True)
        from .. import _cabi

        _cabi.release_pinned(self._slot_prefix())  # staging buffers live for one run
        self._stage_bufs = {}
        release_selection = getattr(self, "_release_selection", None)
        if release_selection is not None:
            release_selection()

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None, progressbar_kwargs=None):
        """Iterate over the trajectory (same arguments as ``maicos.core.AnalysisBase.run``)."""
        return _Runner._run(self, analysis_instances=(self,), start=start, stop=stop, step=step, frames=frames,
                            verbose=verbose, progressbar_kwargs=progressbar_kwargs)

    # ---- output --------------------------------------------------------------------------------
    def savetxt(self, fname: str, X: np.ndarray, columns: list[str] | None = None) -> None:
        """``numpy.savetxt`` with the self-describing header of MAICoS output files."""
        fname = str(fname)
        now = datetime.now().strftime("%a, %b %d %Y at %H:%M:%S ")
        name = self.__class__.__name__
        messages = []
        for cls in self.__class__.mro()[-3::-1]:
            if hasattr(cls, "OUTPUT") and cls.OUTPUT not in messages:
                messages.append(cls.OUTPUT)
        groups = f"  (grp) {atomgroup_header(self.atomgroup)}\n"
        if getattr(self, "refgroup", None) is not None:
            groups += f"  (ref) {atomgroup_header(self.refgroup)}\n"
        header = (
            f"This file was generated by {name} on {now}\n\n"
            f"{name} is part of MAICoS v{__version__}\n\n"
            f"Command line:    {get_cli_input()}\n"
            f"Module input:    {get_module_input_str(self)}\n\n"
            f"Statistics over {self._index} frames\n\n"
            f"Considered atomgroups:\n"
            f"{groups}\n"
            + "\n".join(messages)
            + "\n\n"
        )
        if columns is not None:
            header += "|".join([f"{c:^23}" for c in columns])[3:]
        if not fname.endswith(".dat"):
            fname += ".dat"
        np.savetxt(fname, X, header=header, fmt="% .14e ", encoding="utf8")


class AnalysisCollection(_Runner):
    """Run several analyses over the same trajectory in one loop (``base.py:602-718``).

    Every instance sees the original, untransformed frame.
    """

    def __init__(self, *analysis_instances: AnalysisBase) -> None:
        for a in analysis_instances:
            if a._trajectory is not analysis_instances[0]._trajectory:
                raise ValueError("`analysis_instances` do not have the same trajectory.")
            if not isinstance(a, AnalysisBase):
                raise TypeError(f"Analysis object {a} is not a child of `AnalysisBase`.")
        self._analysis_instances = analysis_instances

    def run(self, start=None, stop=None, step=None, frames=None, verbose=None, progressbar_kwargs=None):
        return _Runner._run(self, analysis_instances=self._analysis_instances, start=start, stop=stop, step=step,
                            frames=frames, verbose=verbose, progressbar_kwargs=progressbar_kwargs)

    def save(self) -> None:
        for a in self._analysis_instances:
            if not a.module_has_save:
                warnings.warn(f"`{a}` has no save() method. Analysis results of this instance can not be written.", stacklevel=2)
            else:
                a.save()


class ProfileBase:
    """Weighted profile + bin counts per frame (``base.py:722-857``).

    The two numpy histogram passes of the reference (``base.py:815-818``) are one fused CUDA
    pass here (``mb200_profile_hist``), batched over the staged frames.
    """

    #: set by modules whose weights depend on topology only (masses, charges, counts)
    _static_weights = False
    def _window_capable(self) -> str:
        """A selection is read out of whole frames of a page-locked trajectory (``AnalysisBase._frame_window``): single
        atoms by index, compounds (device prologue) when they form one contiguous range."""
        return "indices" if self.grouping == "atoms" else "range"

    def __init__(self, atomgroup, grouping: str, bin_method: str, output: str, weighting_function: Callable,
                 weighting_function_kwargs: None | dict, normalization: str) -> None:
        self.atomgroup = atomgroup
        self.grouping = grouping.lower()
        self.bin_method = bin_method.lower()
        self.output = output
        self.normalization = normalization.lower()
        kwargs = {} if weighting_function_kwargs is None else weighting_function_kwargs
        self.weighting_function = lambda ag: weighting_function(ag, grouping, **kwargs)
        if not hasattr(self, "results"):
            self.results = Results()
        if not hasattr(self, "_obs"):
            self._obs = Results()

    def _prepare(self):
        normalizations = ["none", "volume", "number"]
        if self.normalization not in normalizations:
            raise ValueError(
                f"Normalization '{self.normalization}' not supported. Use {', '.join(normalizations)}."
            )
        groupings = ["atoms", "segments", "residues", "molecules", "fragments"]
        if self.grouping not in groupings:
            raise ValueError(
                f"'{self.grouping}' is not a valid option for grouping. Use {', '.join(groupings)}."
            )
        logging.info(f"Atoms grouped by {self.grouping}.")
        if not hasattr(self, "unwrap"):
            self.unwrap = True
        self._cached_weights = None
        self._dev_static_w = None
        # frames per device call: bounded by ~64 MB of staged positions per batch
        n_pos = self.atomgroup.n_atoms
        self._batch_size = int(max(1, min(64, (128 << 20) // max(1, n_pos * 12))))
        # compounds: unwrap / wrap / centres / dipole weights on the device when they are eligible
        self._prologue = None
        self._velocity = None
        mode = getattr(self, "_device_weight_mode", None)          # (weight mode, pdim[, uv mode, uv dim]) of Diporder*
        vmode = getattr(self, "_device_velocity_mode", None)       # (mode, vdim) of Velocity* / Temperature*
        from ..mdshim import Universe as ShimUniverse

        has_vel = vmode is not None and isinstance(self.atomgroup.universe, ShimUniverse) and (
            getattr(self.atomgroup.universe.trajectory, "velocities", None) is not None)
        if (self.grouping != "atoms" and getattr(self, "refgroup", None) is None and not getattr(self, "jitter", 0.0)
                and (self._static_weights or mode is not None or has_vel)
                and is_orthorhombic(self.atomgroup.universe.dimensions)
                and compound_runs(self.atomgroup, self.grouping) is not None):
            wm, pdim, uvm, uvd = (tuple(mode) + (0, 0))[:4] if mode is not None else (0, 0, 0, 0)
            try:
                self._prologue = CompoundPrologue(
                    self.atomgroup, self.grouping, self.bin_method, make_whole=bool(getattr(self, "unwrap", False)),
                    wrap_center=bool(getattr(self, "unwrap", False) or getattr(self, "pack", False)),
                    weight_mode=wm, pdim=pdim, uv_mode=uvm, uv_dim=uvd)
            except ValueError:
                self._prologue = None
        if has_vel and (self.grouping == "atoms" or self._prologue is not None):
            try:
                self._velocity = VelocityWeights(self.atomgroup, vmode[0] if self.grouping == "atoms" else 1, vmode[1],
                                                 prologue=self._prologue)
            except ValueError:
                self._velocity = None
        # atoms grouping with pack=True: the reference wraps every atom of the universe on the host in every frame
        # (base.py:446-448); when nothing else needs the wrapped coordinates (weights from the topology or the velocities)
        # the float32 wrap runs on the device in front of the histogram (``pack_boxes`` of mb200_profile_hist_window)
        dims = self.atomgroup.universe.dimensions
        self._device_pack = bool(
            self.grouping == "atoms" and getattr(self, "pack", False) and not getattr(self, "jitter", 0.0)
            and (self._static_weights or self._velocity is not None)
            and getattr(type(self)._single_frame, "_from_batch", False)
            and dims is not None and is_orthorhombic(dims))
        win = self._frame_window()  # whole frames of a page-locked trajectory travel: size the batches by them
        if win is not None:
            self._batch_size = int(max(1, min(64, (128 << 20) // max(1, win[0] * 12))))

    # geometry classes provide these two
    def _hist_geometry(self) -> dict:
        """``dict(geometry, dim, edges, center, zrange)`` of the current frame for the kernel."""
        raise NotImplementedError("Only implemented in child classes.")

    def _compute_histogram(self, positions: np.ndarray, weights: np.ndarray | None = None) -> np.ndarray:
        """Histogram of ``positions`` in the current frame's bins (CUDA), ``weights=None`` -> counts."""
        from ..lib._hist import profile_hist

        g = self._hist_geometry()
        pos = np.ascontiguousarray(positions)
        if pos.dtype not in (np.float32, np.float64):
            pos = pos.astype(np.float64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        hw, hc = profile_hist(g["geometry"], pos[None], g["dim"], None if w is None else w[None],
                              g["edges"][None], None if g["center"] is None else g["center"][None],
                              None if g["zrange"] is None else g["zrange"][None])
        if weights is None:
            return hc[0] if g["geometry"] != 1 else hc[0].astype(float)
        return hw[0]

    def _profile_positions(self) -> np.ndarray:
        if self.grouping == "atoms":
            return self.atomgroup.positions
        return get_center(self.atomgroup, bin_method=self.bin_method, compound=self.grouping)

    def _profile_weights(self) -> np.ndarray:
        if self._static_weights:
            if self._cached_weights is None:
                self._cached_weights = np.ascontiguousarray(self.weighting_function(self.atomgroup), dtype=np.float64)
            return self._cached_weights
        return np.ascontiguousarray(self.weighting_function(self.atomgroup), dtype=np.float64)

    def _stage_profile(self) -> dict:
        """Positions, weights and bin geometry of the current frame (after the geometry base ran).

        Positions (and per-frame weights) go straight into the pinned batch buffer of the frame's row.
        """
        st = dict(self._hist_geometry())
        if self._velocity is not None:  # raw float32 velocities; the weights are formed on the device
            vbuf, vrow = self._staging_rows((self.atomgroup.n_atoms, 3), np.float32, "vel")
            vel = self._universe.trajectory.ts.velocities
            if getattr(self.atomgroup, "_all", False):
                np.copyto(vbuf[vrow], vel)
            else:
                np.take(vel, self.atomgroup.indices, axis=0, out=vbuf[vrow])
            st["velocities"] = (vbuf, vrow)
        if self._prologue is not None:  # raw atoms; centres and weights are formed on the device
            st["positions"] = self._stage_positions()
            st["box"] = np.array(self._universe.dimensions[:3], dtype=np.float32)
            st["weights"] = self._profile_weights() if self._static_weights else None
            st["obs"] = dict(self._obs)
            return st
        if self.grouping == "atoms":
            buf, row = self._stage_positions()
            if self._device_pack:
                st["box"] = np.array(self._universe.dimensions[:3], dtype=np.float32)
        else:
            centres = get_center(self.atomgroup, bin_method=self.bin_method, compound=self.grouping)
            buf, row = self._staging_rows(centres.shape, np.float64, "pos")
            buf[row] = centres
        st["positions"] = (buf, row)
        if self._static_weights:
            st["weights"] = self._profile_weights()
        elif self._velocity is not None:
            st["weights"] = None
        else:
            w = self._profile_weights()
            wbuf, wrow = self._staging_rows(w.shape, np.float64, "w")
            wbuf[wrow] = w
            st["weights"] = (wbuf, wrow)
        st["obs"] = dict(self._obs)
        return st

    _gather_rows = staticmethod(gather_rows)

    def _announce_profile_rows(self, staged: list) -> None:
        """float32 atom positions of a batch (atoms grouping or the device prologue) are prefetched under the previous call."""
        if staged and (self._prologue is not None or self.grouping == "atoms"):
            self._announce_rows([s["positions"] for s in staged])

    def _selection_on_device(self):
        """Handle of the atom group's indices into whole frames on the calling thread's device (uploaded once per run)."""
        import ctypes

        from .. import _cabi

        ctx = _cabi.context()
        handles = self.__dict__.setdefault("_selection_handles", {})
        if ctx.device not in handles:
            idx = np.ascontiguousarray(self.atomgroup.indices, dtype=np.int32)
            ptr = np.array([0, len(idx)], dtype=np.int64)
            handle = ctypes.c_void_p()
            _cabi.check(_cabi.load().mb200_groups_create(ctx.handle, _cabi.ptr(idx), _cabi.ptr(ptr), 1,
                                                         int(self._universe.atoms.n_atoms), ctypes.byref(handle)))
            handles[ctx.device] = (ctx, handle)
        return handles[ctx.device][1]

    def _release_selection(self) -> None:
        from .. import _cabi

        for ctx, handle in self.__dict__.pop("_selection_handles", {}).values():
            _cabi.load().mb200_groups_destroy(ctx.handle, handle)

    def _static_weights_on_device(self, weights: np.ndarray) -> int:
        """Topology-only weights (masses, charges, ones) are uploaded once per run, not with every batch."""
        from .. import _cabi

        ctx = _cabi.context()
        if self._dev_static_w is None:
            self._dev_static_w = {}
        cached = self._dev_static_w.get(ctx.device)
        if cached is None or cached[0] is not weights:
            buf = _cabi.DeviceBuffer(ctx)
            buf.write(np.ascontiguousarray(weights, dtype=np.float64))
            cached = self._dev_static_w[ctx.device] = (weights, buf)
        return cached[1].ptr.value

    def _compute_profiles(self, staged: list) -> list:
        from ..lib._hist import profile_hist

        g0 = staged[0]
        pos = self._gather_staged([s["positions"] for s in staged])
        edges = np.stack([s["edges"] for s in staged])
        center = None if g0["center"] is None else np.stack([s["center"] for s in staged])
        zrange = None if g0["zrange"] is None else np.stack([s["zrange"] for s in staged])
        vel_w = None
        if self._velocity is not None:
            vel_w = (self._velocity.run(self._gather_staged([s["velocities"] for s in staged], tag="velg")), True)
        if self._prologue is not None:
            c_ptr, w_ptr = self._prologue.run(pos, np.stack([s["box"] for s in staged]), window=self._frame_window())
            dev_w = (w_ptr, True) if w_ptr else vel_w
            if self._static_weights and not w_ptr:
                dev_w = (self._static_weights_on_device(g0["weights"]), False)
            hw, hc = profile_hist(g0["geometry"], None, g0["dim"], None, edges, center, zrange,
                                  device_positions=(c_ptr, len(staged), self._prologue.n_c), device_weights=dev_w)
        else:
            win = self._frame_window() if self.grouping == "atoms" else None  # else: centres formed on the host
            sel = None
            if win is not None and win[1] is None:  # not one range: gathered by index on the device
                sel, win = (self._selection_on_device(), self.atomgroup.n_atoms), None
            win = None if win is None else (win[1], self.atomgroup.n_atoms)  # atoms [first, first + n) of whole frames
            boxes = np.stack([s["box"] for s in staged]) if (self._device_pack and self.grouping == "atoms") else None
            if self._static_weights:
                hw, hc = profile_hist(g0["geometry"], pos, g0["dim"], None, edges, center, zrange, window=win, pack_boxes=boxes,
                                      selection=sel, device_weights=(self._static_weights_on_device(g0["weights"]), False))
            elif vel_w is not None:
                hw, hc = profile_hist(g0["geometry"], pos, g0["dim"], None, edges, center, zrange, device_weights=vel_w, window=win,
                                      pack_boxes=boxes, selection=sel)
            else:
                w = self._gather_rows([s["weights"] for s in staged])
                hw, hc = profile_hist(g0["geometry"], pos, g0["dim"], w, edges, center, zrange, window=win, selection=sel)
        out = []
        for f, s in enumerate(staged):
            obs = s["obs"]
            profile = hw[f]
            obs["bincount"] = hc[f].astype(float) if g0["geometry"] == 1 else hc[f]
            if self.normalization == "volume":
                profile = profile / obs["bin_volume"]
            obs["profile"] = profile
            out.append(obs)
        return out

    def _conclude(self) -> None:
        if self.normalization == "number":
            with np.errstate(divide="ignore", invalid="ignore"):
                self.results.profile = self.sums.profile / self.sums.bincount
        else:
            self.results.profile = self.means.profile
        self.results.dprofile = self.sems.profile

    def save(self) -> None:
        """Write ``positions, profile, error`` columns to ``self.output``."""
        AnalysisBase.savetxt(
            self, self.output,
            np.vstack((self.results.bin_pos, self.results.profile, self.results.dprofile)).T,
            columns=["positions [Å]", "profile", "error"],
        )
