"""Structure factor of the dipolar orientation field: ``DiporderStructureFactor``.

Drop-in for ``maicos.DiporderStructureFactor``
(``/root/reference/src/maicos/modules/diporderstructurefactor.py``): per frame, the centres of the
compounds carry the three Cartesian components of their unit dipole as weights; the structure
factor of each component is binned by ``|q|`` and averaged *per frame* over the q-vectors of a bin,
the three components are summed and divided by the number of compounds (93-155).  The reference
calls the kernel three times on identical positions; ``mb200_dipsf_frames`` shares the y / z phase
tables between the three weight sets.
"""

from __future__ import annotations

import logging

import numpy as np

from .. import _cabi, parallel
from ..core.base import AnalysisBase, CompoundPrologue, compound_runs, gather_rows
from ..lib.util import cell_diagonal, get_center, is_orthorhombic


class DiporderStructureFactor(AnalysisBase):
    r"""Structure factor :math:`S(q)` of :math:`\hat\mu`-weighted compound centres.

    Parameters are those of the reference: ``atomgroup, bin_method="com", grouping="molecules",
    refgroup=None, unwrap=True, pack=True, jitter=0.0, concfreq=0, qmin=0, qmax=6, dq=0.01,
    output="sq.dat"``.  Results: ``results.scattering_vectors``, ``results.structure_factors``.
    """

    _batch_size = 16
    #: a contiguous selection is read out of whole frames of a page-locked trajectory (``AnalysisBase._frame_window``)
    _window_capable = "range"

    def __init__(self, atomgroup, bin_method: str = "com", grouping: str = "molecules", refgroup=None,
                 unwrap: bool = True, pack: bool = True, jitter: float = 0.0, concfreq: int = 0, qmin: float = 0,
                 qmax: float = 6, dq: float = 0.01, output: str = "sq.dat") -> None:
        self._locals = locals()
        super().__init__(atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter,
                         wrap_compound=grouping, concfreq=concfreq)
        self.bin_method = str(bin_method).lower()
        self.qmin = qmin
        self.qmax = qmax
        self.dq = dq
        self.output = output

    def _prepare(self) -> None:
        logging.info("Analysis of the structure factor of dipoles.")
        self.n_bins = int(np.ceil((self.qmax - self.qmin) / self.dq))
        # centres and unit dipoles on the device when the compounds are eligible (contiguous, complete)
        self._prologue = None
        if (self.refgroup is None and not self.jitter and self.wrap_compound != "atoms"
                and is_orthorhombic(self._universe.dimensions)  # the device prologue wraps orthorhombic cells
                and compound_runs(self.atomgroup, self.wrap_compound) is not None):
            try:
                self._prologue = CompoundPrologue(self.atomgroup, self.wrap_compound, self.bin_method,
                                                  make_whole=self.unwrap, wrap_center=self.unwrap or self.pack,
                                                  weight_mode=4)
            except ValueError:
                self._prologue = None
        window = self._frame_window() if self._prologue is not None else None  # whole frames of a page-locked trajectory
        n_staged = window[0] if window is not None else self.atomgroup.n_atoms
        self._batch_size = int(max(1, min(type(self)._batch_size, (128 << 20) // max(1, n_staged * 12))))

    def _stage_frame(self):
        box = np.double(np.array(cell_diagonal(self._ts.dimensions), dtype=np.float32))  # diporderstructurefactor.py:94
        if self._prologue is not None:
            buf, row = self._stage_positions()
            return buf, row, box
        positions = np.ascontiguousarray(
            get_center(atomgroup=self.atomgroup, bin_method=self.bin_method, compound=self.wrap_compound),
            dtype=np.float64,
        )
        # diporder_weights(order_parameter="cos_theta") with the unit vector e_pdim is the pdim-th
        # component of mu / |mu| (weights.py:165-169); the dipoles are computed once, not three times
        dipoles = self.atomgroup.dipole_vector(compound=self.wrap_compound)
        unit = dipoles / np.linalg.norm(dipoles, axis=1)[:, np.newaxis]
        return positions, np.ascontiguousarray(unit.T, dtype=np.float64), box

    def _announce_staged(self, staged: list) -> None:
        if self._prologue is not None:
            self._announce_rows([(st[0], st[1]) for st in staged])
        boxes = np.array([st[2] for st in staged], dtype=np.float64)  # NPT: the frames' q-plans are built ahead (see Saxs)
        if len(boxes) > 1 and not np.all(boxes == boxes[0]):
            ctx = _cabi.context(self._batch_device)
            _cabi.check(_cabi.load().mb200_plans_ahead(ctx.handle, _cabi.ptr(boxes), len(boxes), float(self.qmin), float(self.qmax),
                                                       0.0, float(np.pi), self.n_bins, None, 0))

    def _compute_staged(self, staged: list) -> list:
        F = len(staged)
        boxes = np.stack([s[2] for s in staged])
        on_device = 0
        if self._prologue is not None:
            raw = self._gather_staged([(st[0], st[1]) for st in staged])
            pos, w = self._prologue.run(raw, boxes.astype(np.float32), window=self._frame_window())
            n, on_device = self._prologue.n_c, 1
        else:
            n = staged[0][0].shape[0]
            pos_h = np.stack([s[0] for s in staged])  # keep the arrays alive across the call
            w_h = np.stack([s[1] for s in staged])
            pos, w = _cabi.ptr(pos_h), _cabi.ptr(w_h)
        sb = np.zeros((F, 3, self.n_bins))
        cb = np.zeros((F, 3, self.n_bins), dtype=np.int64)
        block, n_blocks = parallel.q_block()
        ctx = _cabi.context()
        _cabi.check(
            _cabi.load().mb200_dipsf_frames(
                ctx.handle, pos, w, on_device, F, n, _cabi.ptr(boxes), float(self.qmin), float(self.qmax),
                self.n_bins, block, n_blocks, _cabi.ptr(sb), _cabi.ptr(cb),
            )
        )
        if parallel.q_sharded():
            packed = np.concatenate([sb.reshape(-1), cb.reshape(-1).astype(np.float64)])
            parallel.allreduce_sum(packed)
            sb = packed[:sb.size].reshape(sb.shape)
            cb = np.rint(packed[sb.size:]).astype(np.int64).reshape(sb.shape)
        # whole batch at once (this runs between two device calls): sum over pdim of nan_to_num(S / count), same order
        with np.errstate(invalid="ignore", divide="ignore"):
            ratio = np.nan_to_num(sb / cb)
        total = np.zeros((F, self.n_bins))
        for pdim in range(3):
            total += ratio[:, pdim]
        total /= n
        return [({"structure_factors": total[f]}, float(total[f, -1])) for f in range(F)]

    _single_frame = AnalysisBase._single_frame_from_batch

    def _conclude(self) -> None:
        q = np.arange(self.qmin, self.qmax, self.dq) + 0.5 * self.dq
        keep = np.where(self.means.structure_factors != 0)[0]
        self.results.scattering_vectors = q[keep]
        self.results.structure_factors = self.means.structure_factors[keep]

    def save(self) -> None:
        """Write q and S(q) to ``self.output``."""
        self.savetxt(self.output, np.vstack([self.results.scattering_vectors, self.results.structure_factors]).T,
                     columns=["q (1/Å)", "S(q) (arb. units)"])
