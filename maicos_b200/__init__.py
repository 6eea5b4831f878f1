"""B200-native implementation of the MAICoS structure-factor / profile hot path.

The analysis classes keep the API of ``maicos`` (``Saxs``, ``DiporderStructureFactor``,
``DensityPlanar / Cylinder / Sphere``, ``DiporderPlanar``, ...); the arithmetic runs in
``libmaicos_b200.so`` (hand-written CUDA for sm_100a behind the C ABI of ``include/maicos_b200.h``).
There is no CPU fallback.
"""

__version__ = "0.1.0+b200"

from . import core, lib, mdshim, parallel, synthetic  # noqa: E402
from .core import AnalysisBase, AnalysisCollection  # noqa: E402
from .modules import *  # noqa: E402,F401,F403
from .modules import __all__ as _modules_all  # noqa: E402



def set_precision(mode: str = "fp64", device: int | None = None) -> None:
    """Arithmetic of the fused S(q) paths (``Saxs``, ``DiporderStructureFactor``) on this process' GPU.

    ``"fp64"`` -- FP64 phase tables and FP64 DMMA accumulation: the parity mode and the default.
    ``"i8x5"`` -- split precision on the tcgen05 tensor cores (five int8 digits per component, exact
    int32 accumulation in TMEM): several times faster, ~1e-9 relative on binned S(q); isolated weak
    cells of the raw cube can exceed 1e-6, so ``lib.math.structure_factor`` always stays FP64.
    """
    from . import _cabi

    _cabi.context(device).set_precision(mode)


__all__ = ["AnalysisBase", "set_precision", "AnalysisCollection", "core", "lib", "mdshim", "parallel", "synthetic", *_modules_all]
