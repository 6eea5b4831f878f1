"""ORACLE (test infrastructure): loaders for the two CPU structure-factor kernels.

* ``reference_structure_factor`` -- the reference's own Cython kernel
  (``/root/reference/src/maicos/lib/_cmath.pyx``) as built into ``oracle/_ref`` by
  ``oracle/Makefile``.  Travels to the GPU box as a prebuilt ``.so``.
* ``port_structure_factor`` -- the plain-C restatement ``oracle/sfactor_port.c``.
"""

from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sysconfig
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_PORT = _HERE / "liboracle_port.so"
_REF = _HERE / "_ref" / f"_cmath{sysconfig.get_config_var('EXT_SUFFIX')}"


def build(verbose: bool = False) -> None:
    """Build the C port and, when ``/root/reference`` is present, ``oracle/_ref``."""
    targets = ["port"]
    if Path("/root/reference/src/maicos/lib/_cmath.pyx").exists():
        targets.append("ref")
    out = subprocess.run(
        ["make", "-C", str(_HERE), *targets], capture_output=True, text=True, check=False
    )
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise RuntimeError("oracle build failed")


def have_reference() -> bool:
    return _REF.exists()


_ref_mod = None


def _load_ref():
    global _ref_mod
    if _ref_mod is None:
        spec = importlib.util.spec_from_file_location("_cmath", str(_REF))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_mod = mod
    return _ref_mod


def reference_structure_factor(positions, dimensions, qmin, qmax, thetamin, thetamax, weights):
    """Call the compiled reference kernel (same signature as ``_cmath.pyx:21-29``)."""
    return _load_ref().structure_factor(
        positions, dimensions, qmin, qmax, thetamin, thetamax, weights
    )


_port_lib = None


def _load_port():
    global _port_lib
    if _port_lib is None:
        if not _PORT.exists():
            build()
        lib = ctypes.CDLL(str(_PORT))
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_sfactor_maxn.argtypes = [dp, ctypes.c_double, ctypes.POINTER(ctypes.c_int32)]
        lib.oracle_sfactor_maxn.restype = None
        lib.oracle_structure_factor.argtypes = [
            dp, ctypes.c_ssize_t, ctypes.c_ssize_t, ctypes.c_ssize_t, dp,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp, dp,
        ]
        lib.oracle_structure_factor.restype = None
        lib.oracle_structure_factor_slab.argtypes = [
            dp, ctypes.c_ssize_t, ctypes.c_ssize_t, ctypes.c_ssize_t, dp,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, ctypes.c_int32, ctypes.c_int32, dp, dp,
        ]
        lib.oracle_structure_factor_slab.restype = None
        lib.oracle_set_threads.argtypes = [ctypes.c_int]
        lib.oracle_set_threads.restype = None
        lib.oracle_get_threads.restype = ctypes.c_int
        _port_lib = lib
    return _port_lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def port_maxn(dimensions, qmax):
    lib = _load_port()
    dims = np.ascontiguousarray(dimensions, dtype=np.float64)
    out = np.zeros(3, dtype=np.int32)
    lib.oracle_sfactor_maxn(_dp(dims), float(qmax), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return out


def port_structure_factor(positions, dimensions, qmin, qmax, thetamin, thetamax, weights):
    """Plain-C port with the reference's call signature and return layout."""
    lib = _load_port()
    pos = np.asarray(positions)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    dims = np.ascontiguousarray(dimensions, dtype=np.float64)
    if pos.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double'")
    assert dims.shape[0] == 3 and pos.shape[1] == 3 and len(w) == len(pos)
    maxn = port_maxn(dims, qmax)
    q = np.empty(tuple(int(m) for m in maxn), dtype=np.float64)
    s = np.empty_like(q)
    es = pos.itemsize
    lib.oracle_structure_factor(
        _dp(pos), pos.strides[0] // es, pos.strides[1] // es, pos.shape[0], _dp(dims),
        float(qmin), float(qmax), float(thetamin), float(thetamax), _dp(w), _dp(q), _dp(s),
    )
    return q, s


def port_structure_factor_slab(positions, dimensions, qmin, qmax, thetamin, thetamax, weights, h0, h1):
    """The port restricted to Miller indices ``h0 <= h < h1`` (full-size parity checks on a slab of the cube)."""
    lib = _load_port()
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    dims = np.ascontiguousarray(dimensions, dtype=np.float64)
    maxn = port_maxn(dims, qmax)
    q = np.empty(tuple(int(m) for m in maxn), dtype=np.float64)
    s = np.empty_like(q)
    lib.oracle_structure_factor_slab(_dp(pos), 3, 1, pos.shape[0], _dp(dims), float(qmin), float(qmax), float(thetamin),
                                     float(thetamax), _dp(w), int(h0), int(h1), _dp(q), _dp(s))
    return q, s


def structure_factor(positions, dimensions, qmin, qmax, thetamin, thetamax, weights):
    """Best available CPU kernel: the compiled reference if present, else the port."""
    if have_reference():
        return reference_structure_factor(
            np.asarray(positions), np.asarray(dimensions), qmin, qmax, thetamin, thetamax,
            np.asarray(weights),
        )
    return port_structure_factor(positions, dimensions, qmin, qmax, thetamin, thetamax, weights)


def kind() -> str:
    return "reference" if have_reference() else "port"


def n_threads() -> int:
    """Size of the OpenMP pool both kernels run on (they share the process' libgomp)."""
    return int(_load_port().oracle_get_threads())


def set_threads(n: int | None = None) -> int:
    """Give the OpenMP pool ``n`` threads (default: every core).  ``torch.distributed.run`` exports
    ``OMP_NUM_THREADS=1`` to its workers, which would time the reference on one core."""
    _load_port().oracle_set_threads(int(n or os.cpu_count() or 1))
    return n_threads()
