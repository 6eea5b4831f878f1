"""Device binning behind the dielectric profile modules.

The reference's frame bodies (``/root/reference/src/maicos/modules/dielectricplanar.py:152-241``,
``dielectriccylinder.py:131-199``, ``dielectricsphere.py:109-141``) are ``np.digitize`` + ``np.bincount`` pairs over
all atoms, once on the atoms' own coordinate and once per "virtual cut" direction on a (cuts x bins) table that is
cumulated along the cut axis and averaged.  Here

* ``charge_per_bin``  = ``bincount(digitize(coordinate, edges[1:-1]), charges)`` through ``mb200_profile_hist`` with
  ``MB200_BIN_DIGITIZE`` (exact fixed-point accumulation, order independent);
* ``virtual_cut_moment`` uses ``mean_c cumsum_c curQ[c, :] = (1 / n_c) sum_c (n_c - c) curQ[c, :]``: the table
  collapses to ONE weighted histogram over the compound's centre of charge with per-atom weights
  ``q (n_c - cut bin)``; ``mb200_dielectric_prepare`` forms centre and weights on the device for compounds that are
  contiguous runs of atoms (the numpy prologue below is the host-side restatement for arbitrary orderings).
"""

from __future__ import annotations

import numpy as np

from .. import _cabi
from . import _hist


def charge_per_bin(geometry: int, positions: np.ndarray, charges: np.ndarray, edges: np.ndarray, dim: int = 2,
                   center: np.ndarray | None = None) -> np.ndarray:
    """``np.bincount(np.digitize(coordinate, edges[1:-1]), weights=charges, minlength=n_bins)`` on the device.

    ``positions`` float32 / float64 ``[n, 3]``; the coordinate is ``x[dim]`` (planar), the radial distance from the
    axis through ``center`` (cylinder, ``transform_cylinder``) or from ``center`` (sphere, ``transform_sphere``).
    """
    hw, _ = _hist.profile_hist(geometry, positions[None], dim, np.asarray(charges, dtype=np.float64)[None],
                               np.asarray(edges, dtype=np.float64)[None],
                               center=None if center is None else np.asarray(center, dtype=np.float64)[None], digitize=True)
    return hw[0]


class VirtualCuts:
    """Per-frame ``sum_c (n_c - c) curQ[c, :]`` for up to two cut directions (see the module docstring)."""

    def __init__(self, atomgroup, compound: str, geometry: int, dim: int):
        self.ag, self.compound, self.geometry, self.dim = atomgroup, compound, int(geometry), int(dim)
        self.charges = np.ascontiguousarray(atomgroup.charges, dtype=np.float64)
        raw = getattr(atomgroup, "_get_compound_indices", None) or atomgroup._compound_ids  # MDAnalysis / shim
        ids = np.unique(raw(compound), return_inverse=True)[1]
        self.inverse_ix = ids
        self.n = len(ids)
        self.n_c = int(ids.max()) + 1 if self.n else 0
        # device prologue: compounds must be contiguous runs of atoms in the group's order
        self.contiguous = bool(self.n) and bool(np.all(np.diff(ids) >= 0))
        self.ptr = np.concatenate([[0], np.cumsum(np.bincount(ids, minlength=self.n_c))]).astype(np.int64)
        self.ctx = _cabi.context()
        self.coord = _cabi.DeviceBuffer(self.ctx)
        self.weights = _cabi.DeviceBuffer(self.ctx)

    def host_prologue(self, positions: np.ndarray, center, cut_dims, cut_bins, cut_steps):
        """NumPy restatement of ``k_dielectric_prepare`` (any atom order): ``(testpos [n], weights [n_cut, n])``."""
        pos = np.asarray(positions, dtype=np.float64)
        aq = np.abs(self.charges)
        if self.geometry == 0:
            coord = pos[:, self.dim]
        else:
            odims = np.roll(np.arange(3), -self.dim)[1:]
            d = pos[:, odims] - np.asarray(center, dtype=np.float64)[odims]
            coord = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        num = np.bincount(self.inverse_ix, weights=aq * coord, minlength=self.n_c)
        den = np.bincount(self.inverse_ix, weights=aq, minlength=self.n_c)
        with np.errstate(invalid="ignore", divide="ignore"):
            testpos = (num / den)[self.inverse_ix]
        w = np.empty((len(cut_dims), self.n))
        for j, (cd, nb, st) in enumerate(zip(cut_dims, cut_bins, cut_steps)):
            cbin = np.digitize(pos[:, cd], (np.arange(nb) * st)[1:])
            w[j] = self.charges * (nb - cbin)
        return testpos, w

    def moments(self, positions: np.ndarray, edges: np.ndarray, center, cut_dims, cut_bins, cut_steps) -> np.ndarray:
        """``[n_cut, n_bins]``: for cut direction j the histogram over the compound centres of ``q (n_j - cut bin)``.

        ``positions`` float32 ``[n, 3]`` of the atomgroup's atoms (current frame).
        """
        n_cut, n_bins = len(cut_dims), len(edges) - 1
        edges2 = np.asarray(edges, dtype=np.float64)[None]
        out = np.zeros((n_cut, n_bins))
        if self.n == 0:
            return out
        if self.contiguous:
            pos = np.ascontiguousarray(positions, dtype=np.float32)
            c_ptr = self.coord.ensure(self.n * 8)
            w_ptr = self.weights.ensure(n_cut * self.n * 8)
            cd = np.ascontiguousarray(cut_dims, dtype=np.int32)
            cb = np.ascontiguousarray(cut_bins, dtype=np.int32)
            cs = np.ascontiguousarray(cut_steps, dtype=np.float64)
            cen = None if center is None else np.ascontiguousarray(center, dtype=np.float64)
            _cabi.check(_cabi.load().mb200_dielectric_prepare(
                self.ctx.handle, _cabi.ptr(pos), 0, 1, self.n, _cabi.ptr(self.ptr), self.n_c, _cabi.ptr(self.charges),
                self.geometry, self.dim, _cabi.ptr(cen), n_cut, _cabi.ptr(cd), _cabi.ptr(cb), _cabi.ptr(cs), c_ptr, w_ptr))
            for j in range(n_cut):
                hw, _ = _hist.profile_hist(_hist.SCALAR, None, 0, None, edges2, device_positions=(c_ptr, 1, self.n),
                                           device_weights=(w_ptr + j * self.n * 8, False), digitize=True)
                out[j] = hw[0]
        else:
            testpos, w = self.host_prologue(positions, center, cut_dims, cut_bins, cut_steps)
            for j in range(n_cut):
                hw, _ = _hist.profile_hist(_hist.SCALAR, testpos[None], 0, w[j][None], edges2, digitize=True)
                out[j] = hw[0]
        return out
