"""Cylindrical dielectric profiles.

Drop-in for ``maicos.DielectricCylinder`` (``/root/reference/src/maicos/modules/dielectriccylinder.py``): same
constructor signature and defaults, observables ``m_r, m_r_tot, M_r, mM_r, m_z, M_z, mM_z``, results ``eps_z / deps_z /
eps_r / deps_r`` and the two output files.  Charge histograms over the radial coordinate and the axial virtual cuts
(dielectriccylinder.py:134-190) run on the device (:mod:`maicos_b200.lib._dielectric`).
"""

from __future__ import annotations

import logging

import numpy as np

from ..core import CylinderBase
from ..lib import _dielectric, _hist
from ..lib.util import charge_neutral, get_compound
from .dielectricplanar import _prefactor


@charge_neutral(filter="error")
class DielectricCylinder(CylinderBase):
    """Axial and inverse radial components of the cylindrical dielectric tensor.

    Parameters follow the reference: ``atomgroup, bin_width=0.1, temperature=300, single=False,
    output_prefix="eps_cyl", refgroup=None, concfreq=0, jitter=0.0, dim=2, rmin=0, rmax=None, zmin=None, zmax=None,
    vcutwidth=0.1, unwrap=True, pack=True``.  The total axial dipole moment is the correlation observable.
    """

    def __init__(self, atomgroup, bin_width: float = 0.1, temperature: float = 300, single: bool = False,
                 output_prefix: str = "eps_cyl", refgroup=None, concfreq: int = 0, jitter: float = 0.0, dim: int = 2,
                 rmin: float = 0, rmax: float | None = None, zmin: float | None = None, zmax: float | None = None,
                 vcutwidth: float = 0.1, unwrap: bool = True, pack: bool = True) -> None:
        self._locals = locals()
        self.comp = get_compound(atomgroup)
        if zmin is not None or zmax is not None or rmin != 0 or rmax is not None:
            logging.warning(
                "Setting `rmin` and `rmax` (as well as `zmin` and `zmax`) might cut off molecules. This will lead to "
                "severe artifacts in the dielectric profiles."
            )
        super().__init__(atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
                         dim=dim, zmin=zmin, zmax=zmax, bin_width=bin_width, rmin=rmin, rmax=rmax,
                         wrap_compound=self.comp)
        self.output_prefix = output_prefix
        self.temperature = temperature
        self.single = single
        self.vcutwidth = vcutwidth

    def _prepare(self) -> None:
        logging.info("Analysis of the axial and inverse radial components of the cylindrical dielectric tensor.")
        super()._prepare()
        self._cuts = _dielectric.VirtualCuts(self.atomgroup, self.comp, geometry=1, dim=self.dim)
        self.inverse_ix = self._cuts.inverse_ix
        self._charges = np.ascontiguousarray(self.atomgroup.charges, dtype=np.float64)
        self._all_charges = np.ascontiguousarray(self._universe.atoms.charges, dtype=np.float64)

    def _single_frame(self) -> float:
        super()._single_frame()
        o = self._obs
        center = self.box_center
        all_pos = self._universe.atoms.positions
        positions = self.atomgroup.positions
        # charge enclosed per radial bin, for the group and for the whole system (dielectriccylinder.py:134-158)
        curQ_r = _dielectric.charge_per_bin(_hist.CYLINDER, positions, self._charges, o.bin_edges, self.dim, center)
        o.m_r = -np.cumsum(curQ_r) / 2 / np.pi / o.L / o.bin_pos
        curQ_r_tot = _dielectric.charge_per_bin(_hist.CYLINDER, all_pos, self._all_charges, o.bin_edges, self.dim, center)
        o.m_r_tot = -np.cumsum(curQ_r_tot) / 2 / np.pi / o.L / o.bin_pos
        # not the radial dipole moment of the system, but the nomenclature of the other dielectric modules (160-164)
        o.M_r = np.sum(o.m_r_tot * o.bin_width)
        o.mM_r = o.m_r * o.M_r

        # axial component by virtual cuts along z; charges sit at their compound's |q|-weighted mean radius (166-190)
        nbinsz = np.ceil(o.L / self.vcutwidth).astype(int)
        tab = self._cuts.moments(positions, o.bin_edges, center, [self.dim], [int(nbinsz)], [float(o.L / nbinsz)])
        # -(cumsum(curQz, axis=0) / bin_area).mean(axis=0)
        o.m_z = -tab[0] / o.bin_area / int(nbinsz)
        # the system's dipole moment in z direction (193-195)
        o.M_z = np.dot(self._all_charges, np.asarray(all_pos[:, self.dim], dtype=np.float64))
        o.mM_z = o.m_z * o.M_z
        return o.M_z

    def _conclude(self) -> None:
        super()._conclude()
        m, s, r = self.means, self.sems, self.results
        self._pref = _prefactor(self.temperature)
        if not self.single:
            cov_z = m.mM_z - m.m_z * m.M_z
            cov_r = m.mM_r - m.m_r * m.M_r
            dcov_z = np.sqrt(s.mM_z**2 + s.m_z**2 * m.M_z**2 + m.m_z**2 * s.M_z**2)
            dcov_r = np.sqrt(s.mM_r**2 + s.m_r**2 * m.M_r**2 + m.m_r**2 * s.M_r**2)
        else:  # <M> = 0 for a single line of water molecules
            cov_z, cov_r, dcov_z, dcov_r = m.mM_z, m.mM_r, s.mM_z, s.mM_r
        r.eps_z = self._pref * cov_z
        r.deps_z = self._pref * dcov_z
        r.eps_r = -(2 * np.pi * self._obs.L * self._pref * r.bin_pos * cov_r)
        r.deps_r = 2 * np.pi * self._obs.L * self._pref * r.bin_pos * dcov_r

    def save(self) -> None:
        r = self.results
        outdata_z = np.array([r.bin_pos, r.eps_z, r.deps_z]).T
        outdata_r = np.array([r.bin_pos, r.eps_r, r.deps_r]).T
        self.savetxt("{}{}".format(self.output_prefix, "_z.dat"), outdata_z, columns=["positions [Å]", "ε_z - 1", "Δε_z"])
        self.savetxt("{}{}".format(self.output_prefix, "_r.dat"), outdata_r,
                     columns=["positions [Å]", "ε^-1_r - 1", "Δε^-1_r"])
