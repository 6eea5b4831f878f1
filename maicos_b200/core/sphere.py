"""Spherical geometry: radial shells about the box centre.

Protocol of ``maicos.core.sphere`` (``/root/reference/src/maicos/core/sphere.py``): ``rmax``
defaults to half the smallest box length (85-104), shell volumes ``4 pi / 3 * diff(r^3)`` (125-137),
histogram of ``|r - centre|`` (215-223).
"""

from __future__ import annotations

import logging
from collections.abc import Callable

import numpy as np

from ..lib.math import transform_sphere
from .base import AnalysisBase, ProfileBase, Results


class SphereBase(AnalysisBase):
    """Analysis in a spherical geometry with radial limits ``rmin / rmax``."""

    def __init__(self, atomgroup, unwrap: bool, pack: bool, refgroup, jitter: float, concfreq: int, rmin: float,
                 rmax: None | float, bin_width: float, wrap_compound: str):
        super().__init__(atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter,
                         concfreq=concfreq, wrap_compound=wrap_compound)
        self.rmin = rmin
        self._rmax = rmax
        self._bin_width = bin_width

    @property
    def pos_sph(self) -> np.ndarray:
        """Spherical coordinates ``(r, phi, theta)`` of all atoms of the universe (current frame)."""
        return transform_sphere(self._universe.atoms.positions, origin=self.box_center)

    def _compute_lab_frame_sphere(self):
        box_half = self.box_lengths.min() / 2
        if self._rmax is None:
            self.rmax = box_half
        else:
            if self._rmax > box_half:
                logging.warning(
                    f"`rmax` is bigger than half the smallest box vector ({box_half:.2f}) "
                    "in the radial direction. This will lead to artifacts at the edges."
                )
            self.rmax = self._rmax
        self.rmin = np.float64(self.rmin)
        self.rmax = np.float64(self.rmax)

    def _prepare(self):
        self._compute_lab_frame_sphere()
        if self.rmin < 0:
            raise ValueError("Only values for `rmin` larger or equal 0 are allowed.")
        if self._rmax is not None and self._rmax <= self.rmin:
            raise ValueError("`rmax` can not be smaller than or equal to `rmin`!")
        try:
            if self._bin_width > 0:
                self.n_bins = int(np.ceil((self.rmax - self.rmin) / self._bin_width))
            else:
                raise ValueError("Binwidth must be a positive number.")
        except TypeError as err:
            raise ValueError("Binwidth must be a number.") from err

    def _single_frame(self):
        self._compute_lab_frame_sphere()
        o = self._obs
        o.R = self.rmax - self.rmin
        o.bin_edges = np.linspace(self.rmin, self.rmax, self.n_bins + 1, endpoint=True)
        o.bin_width = o.R / self.n_bins
        o.bin_pos = o.bin_edges[1:] - o.bin_width / 2
        o.bin_area = 4 * np.pi * o.bin_pos**2
        o.bin_volume = 4 * np.pi * np.diff(o.bin_edges**3) / 3

    def _conclude(self):
        self.results.bin_pos = self.means.bin_pos


class ProfileSphereBase(SphereBase, ProfileBase):
    """Radial profiles in a sphere."""

    _batch_size = 64

    def __init__(self, atomgroup, unwrap: bool, pack: bool, refgroup, jitter: float, concfreq: int, rmin: float,
                 rmax: None | float, bin_width: float, grouping: str, bin_method: str, output: str,
                 weighting_function: Callable, weighting_function_kwargs: dict | None, normalization: str):
        SphereBase.__init__(self, atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter,
                            concfreq=concfreq, rmin=rmin, rmax=rmax, bin_width=bin_width, wrap_compound=grouping)
        ProfileBase.__init__(self, atomgroup=self.atomgroup, grouping=grouping, bin_method=bin_method, output=output,
                             weighting_function=weighting_function,
                             weighting_function_kwargs=weighting_function_kwargs, normalization=normalization)

    def _prepare(self):
        SphereBase._prepare(self)
        ProfileBase._prepare(self)
        logging.info("Profile along the radial coordinate in a spherical coordinate system.")

    def _hist_geometry(self) -> dict:
        return {"geometry": 2, "dim": 0,
                "edges": np.linspace(self.rmin, self.rmax, self.n_bins + 1, endpoint=True),
                "center": self.box_center, "zrange": None}

    def _stage_frame(self):
        self._obs = Results()
        SphereBase._single_frame(self)
        return self._stage_profile()

    def _announce_staged(self, staged: list) -> None:
        self._announce_profile_rows(staged)

    _single_frame = AnalysisBase._single_frame_from_batch

    def _compute_staged(self, staged: list) -> list:
        return [(obs, obs["profile"][0]) for obs in self._compute_profiles(staged)]

    def _conclude(self):
        SphereBase._conclude(self)
        ProfileBase._conclude(self)
