"""Cylindrical dipolar order parameter profiles.

Drop-in for ``maicos.DiporderCylinder`` (``/root/reference/src/maicos/modules/dipordercylinder.py``): same
constructor signature and defaults, ``results.bin_pos / profile / dprofile`` and output file.  The
binning runs in ``mb200_profile_hist`` through :class:`maicos_b200.core.ProfileCylinderBase`.
"""

from __future__ import annotations

import logging

from ..core import ProfileCylinderBase
from ..lib.weights import diporder_weights
from ..lib.util import unit_vectors_cylinder


class DiporderCylinder(ProfileCylinderBase):
    """Cylindrical dipolar order parameter profiles.

    Parameters follow the reference:
    ``atomgroup, dim=2, zmin=None, zmax=None, bin_width=1, rmin=0, rmax=None, refgroup=None,
    grouping="residues", unwrap=True, pack=True, bin_method="com", output="diporder_cylinder.dat",
    concfreq=0, pdim="r", order_parameter="P0", jitter=0.0``.
    """

    def __init__(
        self, atomgroup, dim: int = 2, zmin: float | None = None, zmax: float | None = None, bin_width:
        float = 1, rmin: float = 0, rmax: float | None = None, refgroup=None, grouping: str = "residues",
        unwrap: bool = True, pack: bool = True, bin_method: str = "com", output: str =
        "diporder_cylinder.dat", concfreq: int = 0, pdim: str = "r", order_parameter: str = "P0", jitter:
        float = 0.0
    ) -> None:
        self._locals = locals()
        wm = {"P0": 1, "cos_theta": 2, "cos_2_theta": 3}.get(order_parameter)
        if wm is not None and pdim in ("r", "z") and dim in (0, 1, 2):
            # (weight mode, Cartesian pdim, unit-vector mode, cylinder axis): "z" projects on e_dim, "r" on the radial direction
            self._device_weight_mode = (wm, dim, 0, 0) if pdim == "z" else (wm, 0, 1, dim)

        def get_unit_vectors(atomgroup, grouping: str):
            return unit_vectors_cylinder(atomgroup=atomgroup, grouping=grouping, bin_method=bin_method, dim=dim, pdim=pdim)

        super().__init__(
            atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
            dim=dim, zmin=zmin, zmax=zmax, bin_width=bin_width, rmin=rmin, rmax=rmax,
            grouping=grouping, bin_method=bin_method, output=output,
            weighting_function=diporder_weights,
            weighting_function_kwargs={"order_parameter": order_parameter, "get_unit_vectors": get_unit_vectors},
            normalization="volume" if order_parameter == "P0" else "number",
        )

    def _prepare(self):
        logging.info("Analysis of the cylindrical dipolar order parameters.")
        super()._prepare()
