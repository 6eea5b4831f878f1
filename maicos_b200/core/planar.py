"""Planar (slab) geometry: bins along one Cartesian axis.

Protocol of ``maicos.core.planar`` (``/root/reference/src/maicos/core/planar.py``):
``PlanarBase`` computes ``zmin / zmax`` in the lab frame from the *current* box every frame
(97-110), fixes ``n_bins`` in ``_prepare`` (112-131) and publishes the bin geometry as observables
(133-149); ``ProfilePlanarBase`` adds the weighted histogram (235-243), the centre-bin correlation
scalar (250) and optional symmetrisation (255-261).
"""

from __future__ import annotations

import logging
from collections.abc import Callable

import numpy as np

from ..lib.math import symmetrize
from .base import AnalysisBase, ProfileBase, Results


class PlanarBase(AnalysisBase):
    """Analysis in a slab geometry along ``dim`` with limits ``zmin / zmax`` relative to the box centre."""

    def __init__(self, atomgroup, unwrap: bool, pack: bool, refgroup, jitter: float, concfreq: int, dim: int,
                 zmin: None | float, zmax: None | float, bin_width: float, wrap_compound: str):
        super().__init__(atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter,
                         concfreq=concfreq, wrap_compound=wrap_compound)
        if dim not in [0, 1, 2]:
            raise ValueError("Dimension can only be x=0, y=1 or z=2.")
        self.dim = dim
        self._zmax = zmax
        self._zmin = zmin
        self._bin_width = bin_width

    @property
    def odims(self) -> np.ndarray:
        """The two dimensions perpendicular to ``dim``, e.g. ``(2, 0)`` for ``dim = 1``."""
        return np.roll(np.arange(3), -self.dim)[1:]

    def _compute_lab_frame_planar(self):
        """Lab-frame limits from the current box (double precision)."""
        zmin = 0 if self._zmin is None else self.box_center[self.dim] + self._zmin
        zmax = self.box_lengths[self.dim] if self._zmax is None else self.box_center[self.dim] + self._zmax
        self.zmin = np.float64(zmin)
        self.zmax = np.float64(zmax)

    def _prepare(self):
        self._compute_lab_frame_planar()
        if self._zmax is not None and self._zmin is not None and self._zmax <= self._zmin:
            raise ValueError("`zmax` can not be smaller or equal than `zmin`!")
        try:
            if self._bin_width > 0:
                self.n_bins = int(np.ceil((self.zmax - self.zmin) / self._bin_width))
            else:
                raise ValueError("Binwidth must be a positive number.")
        except TypeError as err:
            raise ValueError("Binwidth must be a number.") from err

    def _single_frame(self):
        self._compute_lab_frame_planar()
        o = self._obs
        o.L = self.zmax - self.zmin
        o.box_center = self.box_center
        o.bin_edges = np.linspace(self.zmin, self.zmax, self.n_bins + 1)
        o.bin_width = o.L / self.n_bins
        o.bin_pos = o.bin_edges[1:] - o.bin_width / 2
        o.bin_area = np.ones(self.n_bins) * np.prod(self.box_lengths[self.odims])
        o.bin_volume = o.bin_area * o.bin_width

    def _conclude(self):
        # lab frame -> refgroup frame
        self.results.bin_pos = self.means.bin_pos - self.means.box_center[self.dim]


class ProfilePlanarBase(PlanarBase, ProfileBase):
    """Profiles along a Cartesian axis."""

    _batch_size = 64

    def __init__(self, atomgroup, unwrap: bool, pack: bool, refgroup, jitter: float, concfreq: int, dim: int,
                 zmin: None | float, zmax: None | float, bin_width: float, sym: bool, sym_odd: bool, grouping: str,
                 bin_method: str, output: str, weighting_function: Callable,
                 weighting_function_kwargs: None | dict, normalization: str):
        PlanarBase.__init__(self, atomgroup=atomgroup, unwrap=unwrap, pack=pack, refgroup=refgroup, jitter=jitter,
                            concfreq=concfreq, dim=dim, zmin=zmin, zmax=zmax, bin_width=bin_width,
                            wrap_compound=grouping)
        ProfileBase.__init__(self, atomgroup=self.atomgroup, grouping=grouping, bin_method=bin_method, output=output,
                             weighting_function=weighting_function,
                             weighting_function_kwargs=weighting_function_kwargs, normalization=normalization)
        self.sym = sym
        if self.sym and self.refgroup is None:
            raise ValueError("For symmetrization the `refgroup` argument is required.")
        self.sym_odd = sym_odd

    def _prepare(self):
        PlanarBase._prepare(self)
        ProfileBase._prepare(self)
        logging.info(f"""Profile along {"xyz"[self.dim]}-axis normal to the plane.""")

    def _hist_geometry(self) -> dict:
        return {"geometry": 0, "dim": self.dim, "edges": np.linspace(self.zmin, self.zmax, self.n_bins + 1),
                "center": None, "zrange": None}

    def _stage_frame(self):
        self._obs = Results()
        PlanarBase._single_frame(self)
        return self._stage_profile()

    def _announce_staged(self, staged: list) -> None:
        self._announce_profile_rows(staged)

    _single_frame = AnalysisBase._single_frame_from_batch

    def _compute_staged(self, staged: list) -> list:
        return [(obs, obs["profile"][self.n_bins // 2]) for obs in self._compute_profiles(staged)]

    def _conclude(self):
        PlanarBase._conclude(self)
        if self.sym:
            symmetrize(self.sums.profile, inplace=True, is_odd=self.sym_odd)
            symmetrize(self.means.profile, inplace=True, is_odd=self.sym_odd)
            symmetrize(self.sems.profile, inplace=True, is_odd=False)
            if self.normalization == "number":
                symmetrize(self.sums.bincount, inplace=True, is_odd=self.sym_odd)
        ProfileBase._conclude(self)
