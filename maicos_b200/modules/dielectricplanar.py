"""Planar dielectric profiles.

Drop-in for ``maicos.DielectricPlanar`` (``/root/reference/src/maicos/modules/dielectricplanar.py``): same constructor
signature and defaults, per-frame observables (``M, M_perp, M_perp_2, M_par, m_perp, mM_perp, mm_perp, cmM_perp,
cM_perp, m_par, mM_par, mm_par, cmM_par, cM_par``), ``results.eps_par / deps_par / eps_par_self / eps_par_coll /
eps_perp / deps_perp / eps_perp_self / eps_perp_coll`` and the two output files.  The per-atom work of the frame body
(the charge histogram and the virtual-cut tables, dielectricplanar.py:170-213) runs on the device
(:mod:`maicos_b200.lib._dielectric`); the O(n_bins) algebra stays on the host.
"""

from __future__ import annotations

import logging

import numpy as np
import scipy.constants

from ..core import PlanarBase
from ..lib import _dielectric, _hist
from ..lib.math import symmetrize
from ..lib.util import charge_neutral, get_compound


def _prefactor(temperature: float) -> float:
    """1 / (eps0 kB T) in units of Å / e² (dielectricplanar.py:246-251)."""
    pref = 1 / scipy.constants.epsilon_0
    pref /= scipy.constants.Boltzmann * temperature
    pref /= scipy.constants.angstrom / (scipy.constants.elementary_charge) ** 2
    return pref


@charge_neutral(filter="error")
class DielectricPlanar(PlanarBase):
    """Parallel and inverse perpendicular components of the planar dielectric tensor.

    Parameters follow the reference: ``atomgroup, dim=2, zmin=None, zmax=None, bin_width=0.5, refgroup=None,
    is_3d=False, sym=False, unwrap=True, pack=True, temperature=300, output_prefix="eps", concfreq=0, jitter=0.0,
    vcutwidth=0.1``.  For correlation analysis the norm of the parallel total dipole moment is used.
    """

    def __init__(self, atomgroup, dim: int = 2, zmin: float | None = None, zmax: float | None = None,
                 bin_width: float = 0.5, refgroup=None, is_3d: bool = False, sym: bool = False, unwrap: bool = True,
                 pack: bool = True, temperature: float = 300, output_prefix: str = "eps", concfreq: int = 0,
                 jitter: float = 0.0, vcutwidth: float = 0.1) -> None:
        self._locals = locals()
        wrap_compound = get_compound(atomgroup)
        if zmin is not None or zmax is not None:
            logging.warning(
                "Setting `zmin` and `zmax` might cut off molecules. This will lead to severe artifacts in the "
                "dielectric profiles."
            )
        super().__init__(atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, dim=dim,
                         zmin=zmin, zmax=zmax, bin_width=bin_width, wrap_compound=wrap_compound, concfreq=concfreq)
        self.is_3d = is_3d
        self.sym = sym
        self.temperature = temperature
        self.output_prefix = output_prefix
        self.concfreq = concfreq
        self.vcutwidth = vcutwidth

    def _prepare(self) -> None:
        logging.info("Analysis of the parallel and inverse perpendicular components of the planar dielectric tensor.")
        super()._prepare()
        self.comp = get_compound(self.atomgroup)
        self._cuts = _dielectric.VirtualCuts(self.atomgroup, self.comp, geometry=0, dim=self.dim)
        self.inverse_ix = self._cuts.inverse_ix
        self._charges = np.ascontiguousarray(self.atomgroup.charges, dtype=np.float64)

    def _single_frame(self) -> float:
        super()._single_frame()
        o = self._obs
        # total polarisation of the box (dielectricplanar.py:155-162)
        o.M = np.dot(self._universe.atoms.charges, self._universe.atoms.positions)
        o.M_perp = o.M[self.dim]
        o.M_perp_2 = o.M[self.dim] ** 2
        o.M_par = o.M[self.odims]

        positions = self.atomgroup.positions  # float32 [n, 3]
        # perpendicular component from the polarisation density (168-185): charge per z bin, integrated
        curQ = _dielectric.charge_per_bin(_hist.PLANAR, positions, self._charges, o.bin_edges, dim=self.dim)
        o.m_perp = -np.cumsum(curQ / o.bin_area)
        o.mM_perp = o.m_perp * o.M_perp
        o.mm_perp = o.m_perp**2 * o.bin_volume
        o.cmM_perp = o.m_perp * (o.M_perp - o.m_perp * o.bin_volume)
        o.cM_perp = o.M_perp - o.m_perp * o.bin_volume

        # parallel component by virtual cuts (187-225): charges sit at their molecule's centre of charge along z
        dims = self._ts.dimensions
        cut_bins, cut_steps, areas = [], [], []
        for j, direction in enumerate(self.odims):
            Lx = dims[direction]
            areas.append(dims[self.odims[1 - j]] * o.bin_width)
            vbinsx = np.ceil(Lx / self.vcutwidth).astype(int)
            cut_bins.append(int(vbinsx))
            cut_steps.append(float(Lx / vbinsx))
        tab = self._cuts.moments(positions, o.bin_edges, None, [int(d) for d in self.odims], cut_bins, cut_steps)
        o.m_par = np.zeros((self.n_bins, 2))
        for j in range(2):
            # -cumsum(curQx / Ax, axis=0).mean(axis=0) = -(1 / (Ax n_x)) sum_x (n_x - x) curQx[x, :]
            o.m_par[:, j] = -tab[j] / areas[j] / cut_bins[j]

        bin_volume = o.bin_volume[0]  # all bins have the same volume in planar geometry
        o.mM_par = np.dot(o.m_par, o.M_par)
        o.mm_par = (o.m_par * o.m_par).sum(axis=1) * bin_volume
        o.cmM_par = (o.m_par * (o.M_par - o.m_par * bin_volume)).sum(axis=1)
        o.cM_par = o.M_par - o.m_par * bin_volume
        return np.linalg.norm(o.M_par)

    def _conclude(self) -> None:
        super()._conclude()
        m, s, r = self.means, self.sems, self.results
        self._pref = _prefactor(self.temperature)
        r.V = m.bin_volume.sum()

        # perpendicular component (dielectricplanar.py:255-289)
        cov_perp = m.mM_perp - m.m_perp * m.M_perp
        dcov_perp = np.sqrt(s.mM_perp**2 + (m.M_perp * s.m_perp) ** 2 + (m.m_perp * s.M_perp) ** 2)
        var_perp = m.M_perp_2 - m.M_perp**2
        cov_perp_self = m.mm_perp - (m.m_perp**2 * m.bin_volume[0])
        cov_perp_coll = m.cmM_perp - m.m_perp * m.cM_perp
        if not self.is_3d:
            r.eps_perp = -self._pref * cov_perp
            r.eps_perp_self = -self._pref * cov_perp_self
            r.eps_perp_coll = -self._pref * cov_perp_coll
            r.deps_perp = self._pref * dcov_perp
        else:
            r.eps_perp = -cov_perp / (self._pref**-1 + var_perp / r.V)
            r.deps_perp = self._pref * dcov_perp
            r.eps_perp_self = (-self._pref * cov_perp_self) / (1 + self._pref / r.V * var_perp)
            r.eps_perp_coll = (-self._pref * cov_perp_coll) / (1 + self._pref / r.V * var_perp)

        # parallel component (291-319)
        cov_par = 0.5 * (m.mM_par - np.dot(m.m_par, m.M_par.T))
        dcov_par = 0.5 * np.sqrt(s.mM_par**2 + np.dot(s.m_par**2, (m.M_par**2).T) + np.dot(m.m_par**2, (s.M_par**2).T))
        cov_par_self = 0.5 * (m.mm_par - np.dot(m.m_par, m.m_par.sum(axis=0)))
        cov_par_coll = 0.5 * (m.cmM_par - (m.m_par * m.cM_par).sum(axis=1))
        r.eps_par = self._pref * cov_par
        r.deps_par = self._pref * dcov_par
        r.eps_par_self = self._pref * cov_par_self
        r.eps_par_coll = self._pref * cov_par_coll

        if self.sym:
            for key in ("eps_perp", "deps_perp", "eps_perp_self", "eps_perp_coll", "eps_par", "deps_par",
                        "eps_par_self", "eps_par_coll"):
                symmetrize(r[key], axis=0, inplace=True)

    def save(self) -> None:
        r = self.results
        columns = ["position [Å]", "ε^-1_⟂ - 1", "Δε^-1_⟂", "self ε^-1_⟂ - 1", "coll. ε^-1_⟂ - 1"]
        outdata_perp = np.vstack([r.bin_pos, r.eps_perp, r.deps_perp, r.eps_perp_self, r.eps_perp_coll]).T
        self.savetxt("{}{}".format(self.output_prefix, "_perp"), outdata_perp, columns=columns)
        columns = ["position [Å]", "ε_∥ - 1", "Δε_∥", "self ε_∥ - 1", "coll ε_∥ - 1"]
        outdata_par = np.vstack([r.bin_pos, r.eps_par, r.deps_par, r.eps_par_self, r.eps_par_coll]).T
        self.savetxt("{}{}".format(self.output_prefix, "_par"), outdata_par, columns=columns)
