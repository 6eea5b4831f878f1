"""Spherical dipolar order parameter profiles.

Drop-in for ``maicos.DiporderSphere`` (``/root/reference/src/maicos/modules/dipordersphere.py``): same
constructor signature and defaults, ``results.bin_pos / profile / dprofile`` and output file.  The
binning runs in ``mb200_profile_hist`` through :class:`maicos_b200.core.ProfileSphereBase`.
"""

from __future__ import annotations

import logging

from ..core import ProfileSphereBase
from ..lib.weights import diporder_weights
from ..lib.util import unit_vectors_sphere


class DiporderSphere(ProfileSphereBase):
    """Spherical dipolar order parameter profiles.

    Parameters follow the reference:
    ``atomgroup, bin_width=1, rmin=0, rmax=None, refgroup=None, grouping="residues", unwrap=True,
    pack=True, bin_method="com", output="diporder_sphere.dat", concfreq=0, order_parameter="P0",
    jitter=0.0``.
    """

    def __init__(
        self, atomgroup, bin_width: float = 1, rmin: float = 0, rmax: float | None = None, refgroup=None,
        grouping: str = "residues", unwrap: bool = True, pack: bool = True, bin_method: str = "com", output:
        str = "diporder_sphere.dat", concfreq: int = 0, order_parameter: str = "P0", jitter: float = 0.0
    ) -> None:
        self._locals = locals()
        wm = {"P0": 1, "cos_theta": 2, "cos_2_theta": 3}.get(order_parameter)
        if wm is not None:
            self._device_weight_mode = (wm, 0, 2, 0)  # radial unit vectors about the box centre, formed on the device

        def get_unit_vectors(atomgroup, grouping: str):
            return unit_vectors_sphere(atomgroup=atomgroup, grouping=grouping, bin_method=bin_method)

        super().__init__(
            atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
            rmin=rmin, rmax=rmax, bin_width=bin_width,
            grouping=grouping, bin_method=bin_method, output=output,
            weighting_function=diporder_weights,
            weighting_function_kwargs={"order_parameter": order_parameter, "get_unit_vectors": get_unit_vectors},
            normalization="volume" if order_parameter == "P0" else "number",
        )

    def _prepare(self):
        logging.info("Analysis of the spherical dipolar order parameters.")
        super()._prepare()
