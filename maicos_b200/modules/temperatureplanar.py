"""Cartesian temperature profiles.

Drop-in for ``maicos.TemperaturePlanar`` (``/root/reference/src/maicos/modules/temperatureplanar.py``): same
constructor signature and defaults, ``results.bin_pos / profile / dprofile`` and output file.  The
binning runs in ``mb200_profile_hist`` through :class:`maicos_b200.core.ProfilePlanarBase`.
"""

from __future__ import annotations

import logging

from ..core import ProfilePlanarBase
from ..lib.weights import temperature_weights


class TemperaturePlanar(ProfilePlanarBase):
    """Cartesian temperature profiles.

    Parameters follow the reference:
    ``atomgroup, dim=2, zmin=None, zmax=None, bin_width=1, refgroup=None, sym=False, grouping="atoms",
    unwrap=True, pack=True, bin_method="com", output="temperature.dat", concfreq=0, jitter=0.0``.
    """

    def __init__(
        self, atomgroup, dim: int = 2, zmin: float | None = None, zmax: float | None = None, bin_width:
        float = 1, refgroup=None, sym: bool = False, grouping: str = "atoms", unwrap: bool = True, pack:
        bool = True, bin_method: str = "com", output: str = "temperature.dat", concfreq: int = 0, jitter:
        float = 0.0
    ) -> None:
        self._locals = locals()
        if grouping == "atoms":  # other groupings raise in temperature_weights, as in the reference (weights.py:118-123)
            self._device_velocity_mode = (2, 0)  # kinetic temperature per atom on the device (mb200_velocity_weights)
        super().__init__(
            atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter, concfreq=concfreq,
            dim=dim, zmin=zmin, zmax=zmax, bin_width=bin_width, sym=sym, sym_odd=False,
            grouping=grouping, bin_method=bin_method, output=output,
            weighting_function=temperature_weights,
            weighting_function_kwargs=None,
            normalization="number",
        )

    def _prepare(self):
        logging.info("Analysis of the temperature profile.")
        super()._prepare()
