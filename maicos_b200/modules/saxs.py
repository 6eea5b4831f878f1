"""Small-angle X-ray scattering: ``Saxs`` with the reference's API on the CUDA structure-factor path.

Drop-in for ``maicos.Saxs`` (``/root/reference/src/maicos/modules/saxs.py``): same constructor
(99-114), ``results`` layout (``scattering_vectors, structure_factors, scattering_intensities,
dstructure_factors, dscattering_intensities`` and ``miller_indices`` for ``bin_spectrum=False``;
292-300) and output file (303-350).  The reference's semantics are kept, quirks included:
one structure factor *per element group* weighted by ``f_element(q)^2`` (no cross terms, 158-164,
207-210), ``bincount`` taken from the last group (233), errors not divided by the bin count
(252-253).

The frame body (175-241) -- wrap to ``[-L/2, L/2]``, S(q) on the reciprocal lattice, form factors,
three ``|q|`` histograms -- is one fused device call per batch of frames: ``mb200_saxs_frames``.
"""

from __future__ import annotations

import logging

import numpy as np

from .. import _cabi, parallel
from ..core.base import AnalysisBase, gather_rows
from ..lib import tables
from ..lib.math import atomic_form_factor
from ..lib.util import cell_diagonal, is_orthorhombic


def _object_ids(objects: np.ndarray) -> np.ndarray:
    """``id()`` of every entry of an object array.  In CPython ``id`` is the object's address and an object array stores
    exactly those addresses, so the ids are the array's own buffer read as integers (microseconds instead of one Python
    call per atom); any other interpreter / layout takes the generic route."""
    try:
        import ctypes
        import sys

        arr = np.ascontiguousarray(objects)
        if sys.implementation.name == "cpython" and arr.dtype == object and arr.itemsize == ctypes.sizeof(ctypes.c_void_p) and len(arr):
            raw = (ctypes.c_byte * arr.nbytes).from_address(arr.ctypes.data)
            ids = np.frombuffer(raw, dtype=np.intp).astype(np.int64)
            if ids[0] == id(arr[0]) and ids[-1] == id(arr[-1]):
                return ids
    except (TypeError, ValueError, AttributeError):
        pass
    return np.fromiter(map(id, objects), dtype=np.int64, count=len(objects))


def _element_groups(atomgroup):
    """``(sorted unique element names, group id per atom)`` -- what ``np.unique(elements, return_inverse=True)`` gives.

    Topology arrays hold a handful of distinct ``str`` objects referenced a million times, so the atoms are first
    grouped by object identity (integers) and only the few representatives are compared as strings; the result is
    kept on the universe for further analyses of the same selection.
    """
    universe = atomgroup.universe
    cache = universe.__dict__.setdefault("_mb200_element_groups", {}) if hasattr(universe, "__dict__") else {}
    ix = getattr(atomgroup, "ix", None)
    if ix is None:
        key = None
    elif getattr(atomgroup, "_all", False):  # the whole universe: no need to fingerprint a million indices
        key = ("all", len(ix), id(universe))
    else:
        key = (len(ix), hash(np.asarray(ix).tobytes()), id(universe))
    if key is not None and key in cache:
        return cache[key]
    elements = np.asarray(atomgroup.elements)
    if elements.dtype == object and len(elements) > 4096:
        ids = _object_ids(elements)
        # a handful of distinct objects: peel them off one comparison pass at a time (np.unique sorts all atoms twice)
        inv = np.full(len(ids), -1, dtype=np.int64)
        first = []
        while len(first) < 64:
            todo = np.flatnonzero(inv < 0) if first else None
            if todo is not None and len(todo) == 0:
                break
            i0 = 0 if todo is None else int(todo[0])
            inv[ids == ids[i0]] = len(first)
            first.append(i0)
        if (inv < 0).any():  # many distinct string objects after all
            _, first, inv = np.unique(ids, return_index=True, return_inverse=True)
        names = np.array([str(elements[i]) for i in first])
        unique, remap = np.unique(names, return_inverse=True)
        inverse = remap[inv]
    else:
        unique, inverse = np.unique(elements.astype(str), return_inverse=True)
    if ix is not None:
        cache[key] = (unique, inverse)
    return unique, inverse


def _destroy_groups(device: int, handle) -> None:
    try:
        _cabi.load().mb200_groups_destroy(_cabi.context(device).handle, handle)
    except Exception:  # interpreter shutdown: the context is gone with the process
        pass


class _GroupList:
    """``Saxs.groups`` of the reference (one AtomGroup per element, saxs.py:158-164), built when somebody looks: the
    device path works on the index arrays."""

    def __init__(self, atomgroup, index):
        self._atomgroup, self._index, self._made = atomgroup, index, {}

    def __len__(self):
        return len(self._index)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        i = range(len(self))[i]
        if i not in self._made:
            self._made[i] = self._atomgroup[self._index[i]]
        return self._made[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _OnesList(_GroupList):
    """``Saxs.weights``: ones per group (the form factors are applied after the kernel, saxs.py:161-163)."""

    def __init__(self, index):
        self._index, self._made = index, {}

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return np.ones(len(self._index[range(len(self))[i]]))


class Saxs(AnalysisBase):
    r"""Small angle X-ray scattering intensities :math:`I(q) = \sum_e f_e(q)^2 S_e(q)`.

    Parameters
    ----------
    atomgroup : AtomGroup
        atoms to analyse; needs ``elements``.
    unwrap, pack, refgroup, jitter, concfreq
        see :class:`maicos_b200.core.AnalysisBase`.
    bin_spectrum : bool
        bin by :math:`|q|` (default) or keep every Miller vector (needs a constant box).
    qmin, qmax, dq : float
        range and bin width of :math:`|q|` in 1/Å.
    thetamin, thetamax : float
        accepted angle between :math:`q` and the z axis in degrees.
    output : str
        output file name.

    Attributes
    ----------
    results.scattering_vectors, results.structure_factors, results.scattering_intensities,
    results.dstructure_factors, results.dscattering_intensities, results.miller_indices
    """

    _batch_size = 32
    _device_pack = True
    #: any selection is gathered by index on the device: with a page-locked trajectory whole frames are staged as they are
    _window_capable = "indices"

    def __init__(self, atomgroup, unwrap: bool = False, pack: bool = True, refgroup=None, jitter: float = 0.0,
                 concfreq: int = 0, bin_spectrum: bool = True, qmin: float = 0, qmax: float = 6, dq: float = 0.1,
                 thetamin: float = 0, thetamax: float = 180, output: str = "sq.dat") -> None:
        self._locals = locals()
        super().__init__(atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
                         wrap_compound="atoms")
        self.bin_spectrum = bin_spectrum
        self.qmin = qmin
        self.qmax = qmax
        self.dq = dq
        self.thetamin = thetamin
        self.thetamax = thetamax
        self.output = output

    def _prepare(self) -> None:
        logging.info("Analysis of small angle X-ray scattering intensities (SAXS).")
        self.thetamin = min(self.thetamin, self.thetamax)
        self.thetamax = max(self.thetamin, self.thetamax)
        if self.thetamin < 0 or self.thetamin > 180:
            raise ValueError(f"thetamin ({self.thetamin}°) has to between 0 and 180°.")
        if self.thetamax < 0 or self.thetamax > 180:
            raise ValueError(f"thetamax ({self.thetamax}°) has to between 0 and 180°.")
        if self.thetamin > self.thetamax:
            raise ValueError(f"thetamin ({self.thetamin}°) larger than thetamax ({self.thetamax}°).")
        self.thetamin *= np.pi / 180
        self.thetamax *= np.pi / 180

        # one group per element; weights are ones, form factors are applied after the kernel
        unique, inverse = _element_groups(self.atomgroup)
        universe = self.atomgroup.universe
        cache = universe.__dict__.setdefault("_mb200_saxs_groups", {}) if hasattr(universe, "__dict__") else {}
        window = self._frame_window()  # whole frames staged: the groups index the universe's atoms, not the selection's
        key = (id(inverse), window is not None)
        prepared = cache.get(key)  # `inverse` lives in the element-group cache: same selection, same object
        if prepared is None or prepared[0] is not inverse:
            index = [np.flatnonzero(inverse == g).astype(np.int32) for g in range(len(unique))]
            ptr = np.concatenate([[0], np.cumsum([len(i) for i in index])]).astype(np.int64)
            flat = np.ascontiguousarray(np.concatenate(index), dtype=np.int32)
            if window is not None:
                flat = np.ascontiguousarray(np.asarray(self.atomgroup.indices)[flat], dtype=np.int32)
            cm = np.ascontiguousarray(np.stack([tables.cm_packed(str(el)) for el in unique]))
            prepared = cache[key] = (inverse, index, ptr, flat, cm, {})
        _, self._group_index, self._group_ptr, self._group_flat, self._cm, self._groups_handles = prepared
        self._n_staged_atoms = window[0] if window is not None else self.atomgroup.n_atoms
        self.elements = list(unique)
        self.groups = _GroupList(self.atomgroup, self._group_index)
        self.weights = _OnesList(self._group_index)

        # frames per device call: bounded by ~128 MB of staged positions (and by the table workspace)
        self._batch_size = int(max(1, min(type(self)._batch_size, (128 << 20) // max(1, self._n_staged_atoms * 12))))
        if self.bin_spectrum:
            self.n_bins = int(np.ceil((self.qmax - self.qmin) / self.dq))
        else:
            self.box = cell_diagonal(self._universe.dimensions)
            self.scattering_vector_factors = 2 * np.pi / self.box
            self.max_n = np.ceil(self.qmax / self.scattering_vector_factors).astype(int)
            self._cells = None
            # two float64 cubes per frame come back from a batch: at most ~256 MB of them
            self._batch_size = int(max(1, min(self._batch_size, (256 << 20) // max(1, 16 * int(np.prod(self.max_n))))))

    # ---- frame body ------------------------------------------------------------------------
    def _stage_frame(self):
        box = np.array(cell_diagonal(self._ts.dimensions), dtype=np.float32)  # saxs.py:176
        if not self.bin_spectrum and not np.all(box == self.box):
            raise ValueError(
                f"Dimensions in frame {self._frame_index} are different from "
                "initial dimenions. Can not use `bin_spectrum=False`."
            )
        buf, row = self._stage_positions()
        return buf, row, box

    def _announce_staged(self, staged: list) -> None:
        if self.bin_spectrum:
            self._announce_rows([(st[0], st[1]) for st in staged])
            # NPT: every frame has its own cell and needs its own q-plan; they are built here, on the staging thread, while
            # the previous batch computes (constant cell: the one plan is cached after the first batch)
            boxes = np.array([st[2] for st in staged], dtype=np.float64)
            if len(boxes) > 1 and not np.all(boxes == boxes[0]):
                ctx = _cabi.context(self._batch_device)
                _cabi.check(_cabi.load().mb200_plans_ahead(
                    ctx.handle, _cabi.ptr(boxes), len(boxes), float(self.qmin), float(self.qmax), float(self.thetamin),
                    float(self.thetamax), self.n_bins, _cabi.ptr(self._cm), len(self._group_index)))

    def _device_packs(self) -> bool:
        # the kernel's pack is the orthorhombic ``x -= floor(x / L) L``; a triclinic cell is packed by the host
        # (``atoms.wrap`` in ``_transform_frame``) before the frame is staged
        return bool(self.pack and self._universe.dimensions is not None and is_orthorhombic(self._universe.dimensions))

    def _defers_batches(self) -> bool:
        # MAICOS_B200_DEFER=0: one synchronous call per batch (for A/B measurements)
        import os

        return bool(self.bin_spectrum) and os.environ.get("MAICOS_B200_DEFER", "1") != "0"

    def _submit_staged(self, staged: list):
        """Queue a batch (``mb200_saxs_frames_submit``) and return at once; ``_collect_staged`` waits for it."""
        import ctypes

        F = len(staged)
        pos = self._gather_staged([(st[0], st[1]) for st in staged])  # a view of the staging buffer / pinned trajectory
        boxes = np.array([st[2] for st in staged], dtype=np.float32)
        block, n_blocks = parallel.q_block()
        ctx = _cabi.context()
        ticket = ctypes.c_int32(-1)
        rc = _cabi.load().mb200_saxs_frames_submit(
            ctx.handle, _cabi.ptr(pos), 0, F, _cabi.ptr(boxes), self._groups_on(ctx), _cabi.ptr(self._cm),
            float(self.qmin), float(self.qmax), float(self.thetamin), float(self.thetamax), self.n_bins,
            int(self._device_packs()), block, n_blocks, ctypes.byref(ticket))
        if rc == _cabi.EBUSY:  # every slot of the context is taken (many analyses share it): the synchronous call
            return ("done", self._compute_staged(staged))
        _cabi.check(rc)
        return ("ticket", ctx, int(ticket.value), pos)  # `pos` stays referenced until the batch is collected

    def _collect_staged(self, staged: list, state) -> list:
        if state[0] == "done":
            return state[1]
        _, ctx, ticket, _pos = state
        F, G = len(staged), len(self.groups)
        s = np.zeros((F, G, self.n_bins))
        i = np.zeros((F, G, self.n_bins))
        c = np.zeros((F, G, self.n_bins), dtype=np.int64)
        _cabi.check(_cabi.load().mb200_saxs_frames_collect(ctx.handle, ticket, _cabi.ptr(s), _cabi.ptr(i), _cabi.ptr(c)))
        return self._binned_observables(s, i, c)

    def _compute_staged(self, staged: list) -> list:
        if not self.bin_spectrum:
            return self._compute_unbinned(staged)
        F, G = len(staged), len(self.groups)
        pos = self._gather_staged([(st[0], st[1]) for st in staged])  # a view of the staging buffer / pinned trajectory
        boxes = np.array([st[2] for st in staged], dtype=np.float32)
        s = np.zeros((F, G, self.n_bins))
        i = np.zeros((F, G, self.n_bins))
        c = np.zeros((F, G, self.n_bins), dtype=np.int64)
        block, n_blocks = parallel.q_block()
        ctx = _cabi.context()
        handle = self._groups_on(ctx)
        _cabi.check(
            _cabi.load().mb200_saxs_frames_groups(
                ctx.handle, _cabi.ptr(pos), 0, F, _cabi.ptr(boxes), handle, _cabi.ptr(self._cm),
                float(self.qmin), float(self.qmax), float(self.thetamin), float(self.thetamax), self.n_bins,
                int(self._device_packs()), block, n_blocks, _cabi.ptr(s), _cabi.ptr(i), _cabi.ptr(c),
            )
        )
        return self._binned_observables(s, i, c)

    def _binned_observables(self, s: np.ndarray, i: np.ndarray, c: np.ndarray) -> list:
        """Per-frame observables of a batch from the binned per-group sums ``[F, G, n_bins]``."""
        F, G = s.shape[0], s.shape[1]
        if parallel.q_sharded():
            packed = np.concatenate([s.reshape(-1), i.reshape(-1), c.reshape(-1).astype(np.float64)])
            parallel.allreduce_sum(packed)
            m = s.size
            s = packed[:m].reshape(s.shape)
            i = packed[m:2 * m].reshape(s.shape)
            c = np.rint(packed[2 * m:]).astype(np.int64).reshape(s.shape)
        # per-frame observables, formed for the whole batch at once: this runs on the worker thread between two device
        # calls, every microsecond of it is device idle time (tools/e2e_timeline.py)
        sf = s.sum(axis=1) if G > 1 else s[:, 0].copy()
        si = i.sum(axis=1) if G > 1 else i[:, 0].copy()
        bc = np.ascontiguousarray(c[:, G - 1])
        last = s[:, G - 1, -1].tolist()
        return [({"structure_factors": sf[f], "scattering_intensities": si[f], "bincount": bc[f]}, last[f]) for f in range(F)]

    def _groups_on(self, ctx):
        """Handle of the element groups' index lists on ``ctx``'s device (uploaded once per selection and device)."""
        handle = self._groups_handles.get(ctx.device)
        if handle is None:
            import ctypes
            import weakref

            handle = ctypes.c_void_p()
            _cabi.check(_cabi.load().mb200_groups_create(ctx.handle, _cabi.ptr(self._group_flat), _cabi.ptr(self._group_ptr),
                                                         len(self._group_index), self._n_staged_atoms, ctypes.byref(handle)))
            # kept with the universe's prepared selection (further analyses of the same selection reuse it), destroyed
            # with the universe
            self._groups_handles[ctx.device] = handle
            weakref.finalize(self._universe, _destroy_groups, ctx.device, handle)
        return handle

    def _compute_unbinned(self, staged: list) -> list:
        """``bin_spectrum=False``: raw cubes summed over element groups (saxs.py:188-189, 238-239), a batch of frames per
        device call (``mb200_saxs_frames_cells``): the device returns S of the accepted q vectors only, the cubes are
        filled here."""
        lib = _cabi.load()
        F, G = len(staged), len(self.groups)
        box = np.ascontiguousarray(self.box, dtype=np.float32)
        if self._cells is None:  # constant cell: one q-plan per run
            desc = np.zeros(16, dtype=np.int64)
            dims = np.array(box, dtype=np.float64)  # named: the pointer must outlive the call
            _cabi.check(lib.mb200_plan_describe(_cabi.ptr(dims), float(self.qmin), float(self.qmax),
                                                float(self.thetamin), float(self.thetamax), 0, _cabi.ptr(desc)))
            if tuple(desc[1:4]) != tuple(self.max_n):
                # the kernel's cube uses the float32 q factor (_cmath.pyx:93-96), `max_n` the float64 one (saxs.py:150-153);
                # where they differ the reference fails on the cube shapes (saxs.py:238)
                raise ValueError(f"operands could not be broadcast together with shapes {tuple(self.max_n)} {tuple(desc[1:4])}")
            nv = int(desc[0])
            self._cells = (np.zeros(nv, dtype=np.int32), np.zeros(nv), None)
        cells, q, ff2 = self._cells
        nv = len(cells)
        pos = self._gather_staged([(st[0], st[1]) for st in staged])
        s = np.zeros((F, G, nv))
        ctx = _cabi.context()
        _cabi.check(
            lib.mb200_saxs_frames_cells(
                ctx.handle, _cabi.ptr(pos), 0, F, _cabi.ptr(box), self._groups_on(ctx), float(self.qmin), float(self.qmax),
                float(self.thetamin), float(self.thetamax), int(self._device_packs()), nv, _cabi.ptr(cells), _cabi.ptr(q),
                _cabi.ptr(s),
            )
        )
        if ff2 is None:
            ff2 = np.stack([atomic_form_factor(q, str(el)) ** 2 for el in self.elements]) if nv else np.zeros((G, 0))
            self._cells = (cells, q, ff2)
        n_cells = int(np.prod(self.max_n))
        at_end = nv > 0 and int(cells[-1]) == n_cells - 1  # the correlation scalar is the last cell of the last cube
        out = []
        for f in range(F):
            sf = np.zeros(n_cells)
            si = np.zeros(n_cells)
            acc_s = np.zeros(nv)
            acc_i = np.zeros(nv)
            for g in range(G):  # same order of additions as the reference's loop over groups
                acc_s += s[f, g]
                acc_i += ff2[g] * s[f, g]
            sf[cells] = acc_s
            si[cells] = acc_i
            last = float(s[f, G - 1, -1]) if at_end else 0.0
            out.append(({"structure_factors": sf.reshape(self.max_n), "scattering_intensities": si.reshape(self.max_n)}, last))
        return out

    _single_frame = AnalysisBase._single_frame_from_batch

    def _conclude(self) -> None:
        if self.bin_spectrum:
            q = np.arange(self.qmin, self.qmax, self.dq) + 0.5 * self.dq
            with np.errstate(divide="ignore", invalid="ignore"):
                sf = self.sums.structure_factors / self.sums.bincount
                si = self.sums.scattering_intensities / self.sums.bincount
            dsf = self.sems.structure_factors
            dsi = self.sems.scattering_intensities
        else:
            miller = np.indices(tuple(self.max_n)).reshape(3, -1).T  # == np.array(list(np.ndindex(max_n))), saxs.py:256
            q = np.linalg.norm(miller * self.scattering_vector_factors[np.newaxis, :], axis=1)
            order = np.argsort(q)
            q, miller = q[order], miller[order]
            sf = self.means.structure_factors.flatten()[order]
            si = self.means.scattering_intensities.flatten()[order]
        keep = np.invert(np.isnan(sf))
        n = self.atomgroup.n_atoms
        self.results.scattering_vectors = q[keep]
        self.results.structure_factors = sf[keep] / n
        self.results.scattering_intensities = si[keep] / n
        if self.bin_spectrum:
            self.results.dstructure_factors = dsf[keep] / n
            self.results.dscattering_intensities = dsi[keep] / n
        else:
            self.results.miller_indices = miller[keep]

    def save(self) -> None:
        """Write q, S(q), I(q) (and errors or Miller indices) to ``self.output``."""
        r = self.results
        if self.bin_spectrum:
            self.savetxt(
                self.output,
                np.vstack([r.scattering_vectors, r.structure_factors, r.scattering_intensities,
                           r.dstructure_factors, r.dscattering_intensities]).T,
                columns=["q (1/Å)", "S(q) (arb. units)", "I(q) (arb. units)", "ΔS(q)", "ΔI(q)"],
            )
        else:
            out = np.hstack([r.scattering_vectors[:, np.newaxis], r.miller_indices,
                             r.structure_factors[:, np.newaxis], r.scattering_intensities[:, np.newaxis]])
            boxinfo = "box_x = {:.3f} Å, box_y = {:.3f} Å, box_z = {:.3f} Å\n".format(*self.box)
            self.savetxt(self.output, out, columns=[boxinfo, "q (1/Å)", "q_i", "q_j", "q_k", "S(q) (arb. units)",
                                                   "I(q) (arb. units)"])
