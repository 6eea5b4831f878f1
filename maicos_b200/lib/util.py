"""Small host utilities next to the hot path.

Counterparts in the reference (``/root/reference/src/maicos/lib/util.py``): ``get_center``
(642-680), ``unit_vectors_planar`` (684-706), ``correlation_analysis`` (309-355), the output
header helpers ``get_cli_input`` / ``atomgroup_header`` / ``get_module_input_str`` (391-430,
828-869).  The docstring templating and banner plumbing of the reference are out of scope.
"""

from __future__ import annotations

import functools
import inspect
import logging
import sys
import warnings
from pathlib import Path

import numpy as np

from .math import correlation_time


def triclinic_vectors(dimensions, dtype=np.float32) -> np.ndarray:
    """Lower-triangular cell matrix of ``[lx, ly, lz, alpha, beta, gamma]`` (Å, degrees).

    Restates ``MDAnalysis.lib.mdamath.triclinic_vectors`` (MDAnalysis >= 2.8.0, the reference's pinned dependency
    ``pyproject.toml:63``; third party, not vendored): rows are the cell vectors a = (lx, 0, 0), b in the xy plane, c;
    exact zeros for right angles, the zero matrix for an invalid cell, ``float32`` output like ``Timestep.dimensions``.
    The reference takes its reciprocal lattice from the DIAGONAL of this matrix (``modules/saxs.py:169-176``,
    ``modules/diporderstructurefactor.py:94``).
    """
    dim = np.asarray(dimensions, dtype=np.float64)
    lx, ly, lz, alpha, beta, gamma = dim[:6]
    if not (np.all(dim[:6] > 0.0) and alpha < 180.0 and beta < 180.0 and gamma < 180.0):
        return np.zeros((3, 3), dtype=dtype)
    if alpha == beta == gamma == 90.0:
        return np.diag(dim[:3].astype(dtype, copy=False))
    m = np.zeros((3, 3), dtype=np.float64)
    m[0, 0] = lx
    cos_alpha = 0.0 if alpha == 90.0 else np.cos(np.deg2rad(alpha))
    cos_beta = 0.0 if beta == 90.0 else np.cos(np.deg2rad(beta))
    if gamma == 90.0:
        cos_gamma, sin_gamma = 0.0, 1.0
    else:
        g = np.deg2rad(gamma)
        cos_gamma, sin_gamma = np.cos(g), np.sin(g)
    m[1, 0] = ly * cos_gamma
    m[1, 1] = ly * sin_gamma
    m[2, 0] = lz * cos_beta
    m[2, 1] = lz * (cos_alpha - cos_beta * cos_gamma) / sin_gamma
    with np.errstate(invalid="ignore"):
        m[2, 2] = np.sqrt(lz * lz - m[2, 0] ** 2 - m[2, 1] ** 2)
    if m[2, 2] > 0.0:  # NaN or <= 0: two angles sum to no more than the third
        return m.astype(dtype, copy=False)
    return np.zeros((3, 3), dtype=dtype)


def cell_diagonal(dimensions) -> np.ndarray:
    """``np.diag(triclinic_vectors(dimensions))`` -- the box the S(q) modules work with (``saxs.py:176``)."""
    dimensions = np.asarray(dimensions)
    if dimensions.shape[-1] == 3:  # lengths only: an orthorhombic cell
        return dimensions.astype(np.float32)
    if dimensions.ndim == 1 and dimensions[3] == 90.0 and dimensions[4] == 90.0 and dimensions[5] == 90.0 and \
            dimensions[0] > 0.0 and dimensions[1] > 0.0 and dimensions[2] > 0.0:
        # right angles: the matrix is diag(lx, ly, lz) in float32 (the per-frame call of the S(q) frame loops: 2 instead of
        # 25 microseconds)
        return dimensions[:3].astype(np.float32)
    return np.diag(triclinic_vectors(dimensions))


def is_orthorhombic(dimensions) -> bool:
    dimensions = np.asarray(dimensions)
    if dimensions.ndim == 1 and dimensions.shape[0] >= 6:  # scalar comparisons: called once per frame
        return bool(dimensions[3] == 90.0 and dimensions[4] == 90.0 and dimensions[5] == 90.0)
    return dimensions.shape[-1] == 3 or bool(np.all(dimensions[..., 3:6] == 90.0))


def get_center(atomgroup, bin_method: str, compound: str) -> np.ndarray:
    """Per-compound centre: ``cog`` (geometry), ``com`` (mass) or ``coc`` (absolute charge)."""
    if bin_method == "cog":
        weights = None
    elif bin_method == "com":
        weights = atomgroup.masses
    elif bin_method == "coc":
        weights = np.abs(atomgroup.charges)
    else:
        raise ValueError(f"'{bin_method}' is an unknown binning method. Use 'cog', 'com' or 'coc'.")
    return atomgroup.center(weights=weights, compound=compound)


def unit_vectors_planar(atomgroup, grouping: str, pdim: int) -> np.ndarray:  # noqa: ARG001
    """Cartesian unit vector along ``pdim``."""
    e = np.zeros(3)
    e[pdim] = 1.0
    return e


def unit_vectors_cylinder(atomgroup, grouping: str, bin_method: str, dim: int, pdim: str) -> np.ndarray:
    """Cartesian unit vectors of a cylinder frame with axis ``dim``: ``pdim`` = ``"z"`` or ``"r"``."""
    if pdim == "z":
        return unit_vectors_planar(atomgroup, grouping, dim)
    if pdim != "r":
        raise ValueError(f"'{pdim}' is an unknown direction for the projection. Use 'r' or 'z'.")
    rel = get_center(atomgroup=atomgroup, bin_method=bin_method, compound=grouping)
    rel -= atomgroup.universe.dimensions[:3] / 2
    rel[:, dim] = 0  # the radial direction has no axial component
    rel /= np.linalg.norm(rel, axis=1)[:, np.newaxis]
    return rel


def unit_vectors_sphere(atomgroup, grouping: str, bin_method: str) -> np.ndarray:
    """Radial Cartesian unit vectors about the box centre, one per compound."""
    rel = get_center(atomgroup=atomgroup, bin_method=bin_method, compound=grouping)
    rel -= atomgroup.universe.dimensions[:3] / 2
    rel /= np.linalg.norm(rel, axis=1)[:, np.newaxis]
    return rel


def get_compound(atomgroup) -> str:
    """Highest-order topology attribute: "molecules", "fragments", "residues" (lib/util.py:359-388)."""
    if hasattr(atomgroup, "molnums"):
        return "molecules"
    if hasattr(atomgroup, "fragments"):
        logging.info("Cannot use 'molecules'. Falling back to 'fragments'")
        return "fragments"
    if hasattr(atomgroup, "residues"):
        logging.info("Cannot use 'fragments'. Falling back to 'residues'")
        return "residues"
    raise AttributeError("Missing any connection information in `atomgroup`.")


def charge_neutral(filter: str):
    """Class decorator: warn (``filter``) about free charges in the AtomGroup, refuse a non-neutral universe
    (lib/util.py:470-517; tolerance 1e-4 e)."""

    def inner(original_class):
        original_prepare = original_class._prepare

        @functools.wraps(original_prepare)
        def wrapped(self):
            if not np.allclose(self.atomgroup.total_charge(compound=get_compound(self.atomgroup)), 0, atol=1e-4):
                with warnings.catch_warnings():
                    warnings.simplefilter(filter)
                    warnings.warn(
                        "At least one AtomGroup has free charges. Analysis for systems with free charges could lead "
                        "to severe artifacts!",
                        stacklevel=1,
                    )
            if not np.allclose(self.atomgroup.universe.atoms.total_charge(), 0, atol=1e-4):
                raise ValueError("Analysis for non-neutral systems is not supported.")
            return original_prepare(self)

        original_class._prepare = wrapped
        return original_class

    return inner


def correlation_analysis(timeseries: np.ndarray) -> float:
    """Correlation time of the per-frame scalar; warns when frames are visibly correlated."""
    if np.any(np.isnan(timeseries)):
        return -1
    if len(timeseries) <= 4:
        warnings.warn(
            "Your trajectory is too short to estimate a correlation time. Use the "
            "calculated error estimates with caution.",
            stacklevel=2,
        )
        return -1
    corrtime = correlation_time(timeseries)
    if corrtime == -1:
        warnings.warn(
            "Your trajectory does not provide sufficient statistics to estimate a "
            "correlation time. Use the calculated error estimates with caution.",
            stacklevel=2,
        )
    if corrtime > 0.5:
        warnings.warn(
            "Your data seems to be correlated with a correlation time which is "
            f"{corrtime + 1:.2f} times larger than your step size. Consider increasing "
            f"your step size by a factor of {int(np.ceil(2 * corrtime + 1)):d} to get "
            "a reasonable error estimate.",
            stacklevel=2,
        )
    return corrtime


def get_cli_input() -> str:
    """The command line, with arguments containing blanks re-quoted."""
    args = [f'"{a}"' if " " in a else a for a in sys.argv[1:]]
    return "{} {}".format(Path(sys.argv[0]).name, " ".join(args))


def atomgroup_header(ag) -> str:
    """``"<type> <count> & ..."`` summary of an atom group for output headers."""
    if not hasattr(ag, "types"):
        logging.warning("AtomGroup does not contain atom types. Not writing AtomGroup information to output.")
        return f"{len(ag.atoms)} unkown particles"
    names, counts = np.unique(ag.types, return_counts=True)
    return " & ".join(f"{n} {c}" for n, c in zip(names, counts))


def get_module_input_str(obj) -> str:
    """``Module(arg=..., ...).run(...)`` string of the call that produced an output file."""
    name = obj.__class__.__name__
    if not (hasattr(obj, "_locals") and hasattr(obj, "_run_locals")):
        return f"{name}(*args).run(*args)"

    def fmt(param, value):
        if type(value) is str:
            return f"{param}='{value}'"
        if param == "atomgroup" or (param == "refgroup" and value is not None):
            return f"{param}=<AtomGroup>"
        return f"{param}={value}"

    init_args = [p for p in inspect.getfullargspec(obj.__class__).args if p != "self"]
    run_args = [p for p in inspect.getfullargspec(obj.run).args if p != "self"]
    init_sig = ", ".join(fmt(p, obj._locals[p]) for p in init_args)
    run_sig = ", ".join(
        f"{p}='{obj._run_locals[p]}'" if type(obj._run_locals[p]) is str else f"{p}={obj._run_locals[p]}"
        for p in run_args
    )
    return f"{name}({init_sig}).run({run_sig})"
