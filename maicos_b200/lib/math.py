"""Numerical helpers on or next to the hot path (host side).

Public names follow ``maicos.lib.math`` (``/root/reference/src/maicos/lib/math.py``):

* :func:`structure_factor` / :func:`compute_structure_factor` -- the CUDA kernel behind the
  reference's Cython signature (``lib/math.py:15`` re-exports ``_cmath.structure_factor``),
* :func:`atomic_form_factor` (``math.py:623-702``),
* :func:`new_mean`, :func:`new_variance` -- the running frame statistics (``math.py:325-459``),
* :func:`transform_cylinder`, :func:`transform_sphere` (``math.py:705-778``),
* :func:`symmetrize` (``math.py:514-620``), :func:`center_cluster` (``math.py:462-511``),
* :func:`correlation`, :func:`correlation_time` (``math.py:113-322``), O(frames) host helpers
  used only for the ``corrtime`` attribute.
"""

from __future__ import annotations

import numpy as np

from . import tables
from ._sfactor import compute_structure_factor, sfactor_abs_error_bound, sfactor_maxn, structure_factor  # noqa: F401

__all__ = [
    "structure_factor", "compute_structure_factor", "atomic_form_factor", "new_mean", "new_variance",
    "transform_cylinder", "transform_sphere", "symmetrize", "center_cluster", "correlation",
    "correlation_time",
]


def atomic_form_factor(q, element: str):
    r"""Atomic form factor :math:`f(q) = \sum_{i=1}^4 a_i e^{-b_i q^2/(4\pi)^2} + c` in electrons.

    ``element`` is looked up in title case; the united-atom names ``CH1..CH4`` and ``NH1..NH3``
    are the heavy atom plus n hydrogens.  Unknown elements raise ``ValueError``.
    """
    q2 = np.asarray((q / (4 * np.pi)) ** 2)
    flat = q2.reshape(-1)
    total = None
    for mult, cm in tables.cm_parameter_sets(element):
        f = np.sum(cm.a * np.exp(-cm.b * flat[:, None]), axis=1) + cm.c
        f = f if mult == 1.0 else mult * f
        total = f if total is None else total + f
    return total.reshape(q2.shape)


def new_mean(old_mean, data, length: int):
    r"""Mean of ``length`` samples from the mean of the first ``length - 1`` and the newest one.

    :math:`\bar x_N = ((N-1)\,\bar x_{N-1} + x_N)/N`
    """
    return ((length - 1) * old_mean + data) / length


def new_variance(old_variance, old_mean, new_mean, data, length: int):
    r"""Population variance of ``length`` samples, updated from the previous variance and means.

    :math:`N\sigma_N^2 = (N-1)\sigma_{N-1}^2 + (x_N-\bar x_{N-1})(x_N-\bar x_N)`; a slightly
    negative result of rounding is clamped to zero.
    """
    s = old_variance * (length - 1) + (data - old_mean) * (data - new_mean)
    if type(s) is np.ndarray:
        s[s < 0] = 0
    elif s < 0:
        s = 0
    return s / length


def transform_cylinder(positions: np.ndarray, origin: np.ndarray, dim: int) -> np.ndarray:
    """Cartesian -> cylinder coordinates ``(r, phi, z)`` about ``origin`` with axis ``dim``."""
    out = np.zeros(positions.shape)
    odims = np.roll(np.arange(3), -dim)[1:]
    rel = positions - origin
    out[:, 0] = np.linalg.norm(rel[:, odims], axis=1)
    np.arctan2(*rel[:, odims].T, out=out[:, 1])
    out[:, 2] = positions[:, dim]
    return out


def transform_sphere(positions: np.ndarray, origin: np.ndarray) -> np.ndarray:
    """Cartesian -> spherical coordinates ``(r, phi, theta)`` about ``origin``."""
    out = np.zeros(positions.shape)
    rel = positions - origin
    out[:, 0] = np.linalg.norm(rel, axis=1)
    np.arctan2(rel[:, 1], rel[:, 0], out=out[:, 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        np.arccos(rel[:, 2] / out[:, 0], out=out[:, 2])
    return out


def symmetrize(m: np.ndarray, axis=None, inplace: bool = False, is_odd: bool = False) -> np.ndarray:
    """(Anti-)symmetrise ``m`` about the centre of ``axis``: ``(m +/- flip(m)) / 2``."""
    out = m.copy().astype("float")
    out += (-1 if is_odd else 1) * np.flip(m, axis=axis)
    out /= 2
    if inplace:
        m.dtype = np.dtype("float")
        m[...] = out
        return m
    return out


def center_cluster(ag, weights: np.ndarray) -> np.ndarray:
    """Weighted centre of an atom group under periodic boundaries (Bai & Breen 2008).

    Each coordinate is mapped onto a circle, the weighted mean direction is taken and mapped
    back, separately per dimension.
    """
    box = ag.universe.dimensions[:3]
    theta = (ag.positions / box) * 2 * np.pi
    wsum = weights.sum()
    xi = (np.cos(theta) * weights[:, None]).sum(axis=0) / wsum
    zeta = (np.sin(theta) * weights[:, None]).sum(axis=0) / wsum
    return (np.arctan2(-zeta, -xi) + np.pi) / (2 * np.pi) * box


def correlation(a: np.ndarray, b: np.ndarray | None = None, subtract_mean: bool = False) -> np.ndarray:
    """(Cross-)correlation function of equally spaced series via zero-padded FFT."""
    a = np.asarray(a, dtype=float)
    n = len(a)
    size = 2 * (1 << int(np.ceil(np.log2(n)))) if n > 1 else 2

    def spectrum(x):
        x = x - (np.mean(x) if subtract_mean else 0.0)
        buf = np.zeros(size)
        buf[: len(x)] = x
        return np.fft.fft(buf)

    fa = spectrum(a)
    fb = fa if b is None else spectrum(np.asarray(b, dtype=float))
    return np.real(np.fft.ifft(np.conj(fa) * fb)[:n]) / np.arange(n, 0, -1)


def correlation_time(timeseries: np.ndarray, method: str = "sokal", mintime: int = 3, sokal_factor: float = 8) -> float:
    """Integrated correlation time in units of the sampling interval (-1: not enough statistics)."""
    n = len(timeseries)
    if mintime > n:
        raise ValueError(f"mintime ({mintime}) has to be smaller then the length of `timeseries` ({n}).")
    corr = correlation(timeseries, subtract_mean=True)

    def tau_at(cutoff):
        return np.sum((1 - np.arange(1, cutoff) / n) * corr[1:cutoff] / corr[0])

    if method == "sokal":
        tau = -1
        for cutoff in range(mintime, n):
            tau = tau_at(cutoff)
            if cutoff >= sokal_factor * tau:
                break
            if cutoff > n / 3:
                return -1
        return tau
    if method == "chodera":
        cutoff = np.max([mintime, np.min(np.argwhere(corr < 0))])
        return tau_at(cutoff)
    raise ValueError(f"Unknown method: {method}. Chose either 'sokal' or 'chodera'.")
