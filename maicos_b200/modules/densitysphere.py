"""Spherical partial density profiles.

Drop-in for ``maicos.DensitySphere`` (``/root/reference/src/maicos/modules/densitysphere.py``): same
constructor signature and defaults, ``results.bin_pos / profile / dprofile`` and output file.  The
binning runs in ``mb200_profile_hist`` through :class:`maicos_b200.core.ProfileSphereBase`.
"""

from __future__ import annotations

import logging

from ..core import ProfileSphereBase
from ..lib.weights import density_weights


class DensitySphere(ProfileSphereBase):
    """Spherical partial density profiles.

    Parameters follow the reference:
    ``atomgroup, dens="mass", bin_width=1, rmin=0, rmax=None, refgroup=None, grouping="atoms",
    unwrap=True, pack=True, bin_method="com", output="density.dat", concfreq=0, jitter=0.0``.
    """

    #: density weights depend on the topology only -> uploaded once, not per frame
    _static_weights = True

    def __init__(
        self, atomgroup, dens: str = "mass", bin_width: float = 1, rmin: float = 0, rmax: float | None =
        None, refgroup=None, grouping: str = "atoms", unwrap: bool = True, pack: bool = True, bin_method:
        str = "com", output: str = "density.dat", concfreq: int = 0, jitter: float = 0.0
    ) -> None:
        self._locals = locals()
        super().__init__(
            atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
            rmin=rmin, rmax=rmax, bin_width=bin_width,
            grouping=grouping, bin_method=bin_method, output=output,
            weighting_function=density_weights,
            weighting_function_kwargs={"dens": dens},
            normalization="volume",
        )

    def _prepare(self):
        logging.info(f"Analysis of the {self._locals['dens']} density profile.")
        super()._prepare()
