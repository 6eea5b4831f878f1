"""Library layer: the CUDA-backed kernels behind the reference's function names + host helpers."""

from . import math, tables, util, weights  # noqa: F401

__all__ = ["math", "tables", "util", "weights"]
