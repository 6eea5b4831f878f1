"""Host wrapper of the CUDA structure-factor kernel behind the reference's call signature.

Mirrors ``maicos.lib._cmath.structure_factor`` (``/root/reference/src/maicos/lib/_cmath.pyx:21-126``):
same positional arguments, same return layout ``(scattering_vectors, structure_factors)`` as two
fresh C-contiguous float64 arrays of shape ``(maxn0, maxn1, maxn2)`` that are exactly zero in
rejected cells, and the same error behaviour (``ValueError`` for a dtype mismatch,
``AssertionError`` for shape violations).  The arithmetic runs in ``libmaicos_b200.so``.
"""

from __future__ import annotations

import ctypes

import numpy as np

from .. import _cabi


def _as_double_buffer(a, name: str, ndim: int) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim != ndim:
        raise ValueError(f"Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})")
    if a.dtype != np.float64:
        # same text Cython raises for a `double[:]` memoryview fed with another dtype
        kind = {"f": "float", "i": "long", "u": "unsigned long"}.get(a.dtype.kind, str(a.dtype))
        if a.dtype == np.float32:
            kind = "float"
        raise ValueError(f"Buffer dtype mismatch, expected 'double' but got '{kind}'")
    return a


def sfactor_maxn(dimensions, qmax: float) -> np.ndarray:
    """``maxn`` of the reference kernel (``_cmath.pyx:93-96``), incl. its float32 cast."""
    dims = np.ascontiguousarray(dimensions, dtype=np.float64)
    out = np.zeros(3, dtype=np.int32)
    _cabi.check(
        _cabi.load().mb200_sfactor_maxn(
            dims.ctypes.data_as(_cabi.c_double_p), float(qmax), out.ctypes.data_as(_cabi.c_i32_p)
        )
    )
    return out


def structure_factor(positions, dimensions, qmin, qmax, thetamin, thetamax, weights, *, block=None, device=None):
    r"""Scattering vectors and structure factors on the reciprocal lattice of the cell.

    .. math::
        S(\boldsymbol{q}) = \left[\sum_k w_k \cos(\boldsymbol{q}\boldsymbol{r}_k)\right]^2
                          + \left[\sum_k w_k \sin(\boldsymbol{q}\boldsymbol{r}_k)\right]^2

    evaluated for every Miller triple of the non-negative octant with
    ``qmin <= |q| <= qmax`` and ``thetamin <= theta <= thetamax`` (radians), see
    ``_cmath.pyx:101-124``.

    Parameters are those of the reference; ``block=(i, n)`` (keyword only, not in the reference)
    restricts the evaluation to q-block ``i`` of ``n`` for multi-GPU q sharding.
    """
    positions = _as_double_buffer(positions, "positions", 2)
    dimensions = _as_double_buffer(dimensions, "dimensions", 1)
    weights = _as_double_buffer(weights, "weights", 1)
    assert dimensions.shape[0] == 3
    assert positions.shape[1] == 3
    assert len(weights) == len(positions)
    weights = np.ascontiguousarray(weights)
    dims = np.ascontiguousarray(dimensions)

    maxn = sfactor_maxn(dims, qmax)
    shape = tuple(int(max(m, 0)) for m in maxn)
    q = np.zeros(shape, dtype=np.float64)
    s = np.zeros(shape, dtype=np.float64)
    if q.size == 0:
        return q, s
    ctx = _cabi.context(device)
    es = positions.itemsize
    if positions.strides[0] % es or positions.strides[1] % es:
        positions = np.ascontiguousarray(positions)
    b, nb = (0, 1) if block is None else (int(block[0]), int(block[1]))
    _cabi.check(
        _cabi.load().mb200_structure_factor_block(
            ctx.handle, _cabi.ptr(positions), positions.strides[0] // es, positions.strides[1] // es,
            positions.shape[0], dims.ctypes.data_as(_cabi.c_double_p), float(qmin), float(qmax),
            float(thetamin), float(thetamax), _cabi.ptr(weights), b, nb, _cabi.ptr(q), _cabi.ptr(s),
        )
    )
    return q, s


def sfactor_abs_error_bound(n_atoms: int, max_abs_weight: float = 1.0, precision: str = "i8x5") -> float:
    r"""Worst-case absolute error of one complex amplitude :math:`A(\boldsymbol q) = \sum_k w_k e^{i q r_k}`.

    ``S = |A|^2`` is what the kernel returns; its error is ``|dS| <= 2 |A| |dA| + |dA|^2``.  In the default
    split-precision arithmetic every table entry is quantised to :math:`2^{-38}` *absolutely* (after ``max|w|`` has
    been brought into ``(0.5, 1]`` by an exact power of two) and digit products below :math:`2^{-36}` are dropped:
    per atom and amplitude component at most :math:`2^{-37}` (quantisation) + :math:`2^{-35}` (truncation), hence

    .. math:: |dA| \le 16\, N\, 2^{-38}\, 2^{\lceil \log_2 \max|w| \rceil}.

    On disordered input the errors add incoherently (measured :math:`\approx \sqrt N\,2^{-38}`, i.e. a relative error of
    5e-9 on S where :math:`|A| \sim \sqrt N`); on a lattice the same few phase values repeat and the bound is what holds:
    where amplitudes cancel (``|A|`` below the bound) the *relative* error of S is unbounded and the ``S != 0`` mask can
    differ from the reference's, whose own rounding noise there is about ``N 2^-50`` (``tests/test_gpu_parity_r2.py``).
    ``precision="fp64"``: FP64 tables and accumulation, about ``N 2^-50``.
    """
    n = float(max(int(n_atoms), 1))
    w = float(max_abs_weight)
    scale = 1.0 if not np.isfinite(w) or w <= 0 else 2.0 ** np.ceil(np.log2(w))
    return (16.0 * n * 2.0 ** -38 if precision != "fp64" else n * 2.0 ** -50) * scale


#: pre-v0.11 name of the same function (``CHANGELOG.rst:38-40``)
compute_structure_factor = structure_factor
