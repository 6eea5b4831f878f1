"""Cartesian dipolar order parameter profiles.

Drop-in for ``maicos.DiporderPlanar`` (``/root/reference/src/maicos/modules/diporderplanar.py``): same
constructor signature and defaults, ``results.bin_pos / profile / dprofile`` and output file.  The
binning runs in ``mb200_profile_hist`` through :class:`maicos_b200.core.ProfilePlanarBase`.
"""

from __future__ import annotations

import logging

from ..core import ProfilePlanarBase
from ..lib.weights import diporder_weights
from ..lib.util import unit_vectors_planar


class DiporderPlanar(ProfilePlanarBase):
    """Cartesian dipolar order parameter profiles.

    Parameters follow the reference:
    ``atomgroup, dim=2, zmin=None, zmax=None, bin_width=1, refgroup=None, sym=False,
    grouping="residues", unwrap=True, pack=True, bin_method="com", output="diporder_planar.dat",
    concfreq=0, pdim=2, order_parameter="P0", jitter=0.0``.
    """

    def __init__(
        self, atomgroup, dim: int = 2, zmin: float | None = None, zmax: float | None = None, bin_width:
        float = 1, refgroup=None, sym: bool = False, grouping: str = "residues", unwrap: bool = True, pack:
        bool = True, bin_method: str = "com", output: str = "diporder_planar.dat", concfreq: int = 0, pdim:
        int = 2, order_parameter: str = "P0", jitter: float = 0.0
    ) -> None:
        self._locals = locals()
        #: (mode, pdim) of mb200_compound_prologue: the projection needs no host data
        self._device_weight_mode = ({"P0": 1, "cos_theta": 2, "cos_2_theta": 3}.get(order_parameter), pdim)
        if self._device_weight_mode[0] is None:
            self._device_weight_mode = None

        def get_unit_vectors(atomgroup, grouping: str):
            return unit_vectors_planar(atomgroup=atomgroup, grouping=grouping, pdim=pdim)

        super().__init__(
            atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
            dim=dim, zmin=zmin, zmax=zmax, bin_width=bin_width, sym=sym, sym_odd=True,
            grouping=grouping, bin_method=bin_method, output=output,
            weighting_function=diporder_weights,
            weighting_function_kwargs={"order_parameter": order_parameter, "get_unit_vectors": get_unit_vectors},
            normalization="volume" if order_parameter == "P0" else "number",
        )

    def _prepare(self):
        logging.info("Analysis of the cartesian dipolar order parameters.")
        super()._prepare()
