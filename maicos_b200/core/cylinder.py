"""Cylindrical geometry: radial bins about the box centre, axis ``dim``.

Protocol of ``maicos.core.cylinder`` (``/root/reference/src/maicos/core/cylinder.py``): ``rmax``
defaults to half the smallest perpendicular box length (90-109), bin areas ``pi * diff(r^2)`` and
volumes ``area * L`` (132-145), the histogram is a ``(n_bins, 1)`` 2-D histogram over ``(r, z)`` so
atoms outside the axial window are dropped (229-243).
"""

from __future__ import annotations

import logging
from collections.abc import Callable

import numpy as np

from ..lib.math import transform_cylinder
from .base import AnalysisBase, ProfileBase, Results
from .planar import PlanarBase


class CylinderBase(PlanarBase):
    """Analysis in a cylindrical geometry (radial limits ``rmin / rmax``, axial ``zmin / zmax``)."""

    def __init__(self, atomgroup, unwrap: bool, pack: bool, refgroup, jitter: float, concfreq: int, dim: int,
                 zmin: None | float, zmax: None | float, bin_width: float, rmin: float, rmax: None | float,
                 wrap_compound: str):
        super().__init__(atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter,
                         concfreq=concfreq, wrap_compound=wrap_compound, dim=dim, zmin=zmin, zmax=zmax,
                         bin_width=bin_width)
        self.rmin = rmin
        self._rmax = rmax

    @property
    def pos_cyl(self) -> np.ndarray:
        """Cylinder coordinates ``(r, phi, z)`` of all atoms of the universe (current frame)."""
        return transform_cylinder(self._universe.atoms.positions, origin=self.box_center, dim=self.dim)

    def _compute_lab_frame_cylinder(self):
        box_half = self.box_lengths[self.odims].min() / 2
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
        super()._prepare()
        self._compute_lab_frame_cylinder()
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
        super()._single_frame()
        self._compute_lab_frame_cylinder()
        o = self._obs
        o.R = self.rmax - self.rmin
        o.bin_edges = np.linspace(self.rmin, self.rmax, self.n_bins + 1, endpoint=True)
        o.bin_width = o.R / self.n_bins
        o.bin_pos = o.bin_edges[1:] - o.bin_width / 2
        o.bin_area = np.pi * np.diff(o.bin_edges**2)
        o.bin_volume = o.bin_area * o.L

    def _conclude(self):
        super()._conclude()
        self.results.bin_pos = self.means.bin_pos


class ProfileCylinderBase(CylinderBase, ProfileBase):
    """Radial profiles in a cylinder."""

    _batch_size = 64

    def __init__(self, atomgroup, unwrap: bool, pack: bool, refgroup, jitter: float, concfreq: int, dim: int,
                 zmin: None | float, zmax: None | float, bin_width: float, rmin: float, rmax: None | float,
                 grouping: str, bin_method: str, output: str, weighting_function: Callable,
                 weighting_function_kwargs: None | dict, normalization: str):
        CylinderBase.__init__(self, atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter,
                              concfreq=concfreq, dim=dim, zmin=zmin, zmax=zmax, bin_width=bin_width, rmin=rmin,
                              rmax=rmax, wrap_compound=grouping)
        ProfileBase.__init__(self, atomgroup=self.atomgroup, grouping=grouping, bin_method=bin_method, output=output,
                             weighting_function=weighting_function,
                             weighting_function_kwargs=weighting_function_kwargs, normalization=normalization)

    def _prepare(self):
        CylinderBase._prepare(self)
        ProfileBase._prepare(self)
        logging.info(
            "Profile along the radial axis in a cylindrical coordinate system, "
            f"""with the {"xyz"[self.dim]}-axis as cylindrical axis."""
        )

    def _hist_geometry(self) -> dict:
        return {"geometry": 1, "dim": self.dim,
                "edges": np.linspace(self.rmin, self.rmax, self.n_bins + 1, endpoint=True),
                "center": self.box_center, "zrange": np.array([self.zmin, self.zmax], dtype=np.float64)}

    def _stage_frame(self):
        self._obs = Results()
        CylinderBase._single_frame(self)
        return self._stage_profile()

    def _announce_staged(self, staged: list) -> None:
        self._announce_profile_rows(staged)

    _single_frame = AnalysisBase._single_frame_from_batch

    def _compute_staged(self, staged: list) -> list:
        return [(obs, obs["profile"][0]) for obs in self._compute_profiles(staged)]

    def _conclude(self):
        CylinderBase._conclude(self)
        ProfileBase._conclude(self)
