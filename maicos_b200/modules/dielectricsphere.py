"""Spherical dielectric profiles.

Drop-in for ``maicos.DielectricSphere`` (``/root/reference/src/maicos/modules/dielectricsphere.py``): same constructor
signature and defaults, observables ``m_r, m_r_tot, M_r, mM_r``, results ``eps_rad / deps_rad`` and output file.  The two
charge histograms over the radial coordinate (dielectricsphere.py:113-133) run on the device.
"""

from __future__ import annotations

import logging

import numpy as np

from ..core import SphereBase
from ..lib import _dielectric, _hist
from ..lib.util import charge_neutral, get_compound
from .dielectricplanar import _prefactor


@charge_neutral(filter="error")
class DielectricSphere(SphereBase):
    """Inverse radial component of the spherical dielectric tensor.

    Parameters follow the reference: ``atomgroup, bin_width=0.1, temperature=300, output_prefix="eps_sph",
    refgroup=None, concfreq=0, jitter=0.0, rmin=0, rmax=None, unwrap=True, pack=True``.
    """

    def __init__(self, atomgroup, bin_width: float = 0.1, temperature: float = 300, output_prefix: str = "eps_sph",
                 refgroup=None, concfreq: int = 0, jitter: float = 0.0, rmin: float = 0, rmax: float | None = None,
                 unwrap: bool = True, pack: bool = True) -> None:
        self._locals = locals()
        self.comp = get_compound(atomgroup)
        if rmin != 0 or rmax is not None:
            logging.warning(
                "Setting `rmin` and `rmax` might cut off molecules. This will lead to severe artifacts in the "
                "dielectric profiles."
            )
        super().__init__(atomgroup, concfreq=concfreq, jitter=jitter, refgroup=refgroup, rmin=rmin, rmax=rmax,
                         bin_width=bin_width, unwrap=unwrap, pack=pack, wrap_compound=self.comp)
        self.output_prefix = output_prefix
        self.bin_width = bin_width
        self.temperature = temperature

    def _prepare(self) -> None:
        logging.info("Analysis of the inverse radial component of the spherical dielectric tensor.")
        super()._prepare()
        self._charges = np.ascontiguousarray(self.atomgroup.charges, dtype=np.float64)
        self._all_charges = np.ascontiguousarray(self._universe.atoms.charges, dtype=np.float64)

    def _single_frame(self) -> float:
        super()._single_frame()
        o = self._obs
        center = self.box_center
        # charge enclosed per radial shell for the group and for the whole system (dielectricsphere.py:113-133)
        curQ = _dielectric.charge_per_bin(_hist.SPHERE, self.atomgroup.positions, self._charges, o.bin_edges, 0, center)
        o.m_r = -np.cumsum(curQ) / 4 / np.pi / o.bin_pos**2
        curQ_tot = _dielectric.charge_per_bin(_hist.SPHERE, self._universe.atoms.positions, self._all_charges,
                                              o.bin_edges, 0, center)
        o.m_r_tot = -np.cumsum(curQ_tot) / 4 / np.pi / o.bin_pos**2
        o.M_r = np.sum(o.m_r_tot * o.bin_width)
        o.mM_r = o.m_r * o.M_r
        return o.M_r

    def _conclude(self) -> None:
        super()._conclude()
        m, s, r = self.means, self.sems, self.results
        self._pref = _prefactor(self.temperature)
        cov_rad = m.mM_r - m.m_r * m.M_r
        dcov_rad = np.sqrt(s.mM_r**2 + s.m_r**2 * m.M_r**2 + m.m_r**2 * s.M_r**2)
        r.eps_rad = 1 - (4 * np.pi * r.bin_pos**2 * self._pref * cov_rad)
        r.deps_rad = 4 * np.pi * r.bin_pos**2 * self._pref * dcov_rad

    def save(self) -> None:
        r = self.results
        outdata_rad = np.array([r.bin_pos, r.eps_rad, r.deps_rad]).T
        self.savetxt("{}{}".format(self.output_prefix, "_rad.dat"), outdata_rad,
                     columns=["positions [Å]", "eps_rad - 1", "eps_rad error"])
