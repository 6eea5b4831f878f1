"""Cylindrical velocity (or flux) profiles.

Drop-in for ``maicos.VelocityCylinder`` (``/root/reference/src/maicos/modules/velocitycylinder.py``): same
constructor signature and defaults, ``results.bin_pos / profile / dprofile`` and output file.  The
binning runs in ``mb200_profile_hist`` through :class:`maicos_b200.core.ProfileCylinderBase`.
"""

from __future__ import annotations

import logging

from ..core import ProfileCylinderBase
from ..lib.weights import velocity_weights


class VelocityCylinder(ProfileCylinderBase):
    """Cylindrical velocity (or flux) profiles.

    Parameters follow the reference:
    ``atomgroup, dim=2, zmin=None, zmax=None, bin_width=1, rmin=0, rmax=None, refgroup=None,
    grouping="atoms", unwrap=True, pack=True, bin_method="com", output="velocity.dat", concfreq=0,
    vdim=0, flux=False, jitter=0.0``.
    """

    def __init__(
        self, atomgroup, dim: int = 2, zmin: float | None = None, zmax: float | None = None, bin_width:
        float = 1, rmin: float = 0, rmax: float | None = None, refgroup=None, grouping: str = "atoms",
        unwrap: bool = True, pack: bool = True, bin_method: str = "com", output: str = "velocity.dat",
        concfreq: int = 0, vdim: int = 0, flux: bool = False, jitter: float = 0.0
    ) -> None:
        self._locals = locals()
        if vdim not in [0, 1, 2]:
            raise ValueError("Velocity dimension can only be x=0, y=1 or z=2.")
        self._device_velocity_mode = (0, vdim)  # velocity_weights on the device (mb200_velocity_weights)
        super().__init__(
            atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
            dim=dim, zmin=zmin, zmax=zmax, bin_width=bin_width, rmin=rmin, rmax=rmax,
            grouping=grouping, bin_method=bin_method, output=output,
            weighting_function=velocity_weights,
            weighting_function_kwargs={"vdim": vdim},
            normalization="volume" if flux else "number",
        )

    def _prepare(self):
        logging.info("Analysis of the velocity profile.")
        super()._prepare()
