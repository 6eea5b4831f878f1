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



def set_precision(mode: str = "i8x5", device: int | None = None) -> None:
    """Arithmetic of the S(q) contraction (``Saxs``, ``DiporderStructureFactor``, ``lib.math.structure_factor``).

    ``"i8x5"`` -- the default: split precision on the tcgen05 tensor cores (five int8 digits per component =
    40-bit fixed point, exact int32 accumulation in TMEM, FP64 recombination).  Measured max relative error
    on any nonzero cell of the raw S cube 4.4e-9 (1.5k .. 1M atoms), 3e-13 on binned S(q); stated tolerance 1e-6.
    ``"fp64"`` -- FP64 phase tables and FP64 DMMA accumulation (~1e-12 against the reference kernel), ~2.6x slower.
    The environment variable ``MAICOS_B200_PRECISION`` selects the mode for the CLI.  Without ``device`` the mode applies
    to every device context of the process.
    """
    from . import _cabi

    if device is not None:
        _cabi.context(device).set_precision(mode)
    else:  # every context of this process, existing and future (single-process multi-GPU runs use one per device)
        _cabi.set_default_precision(mode)


__all__ = ["AnalysisBase", "set_precision", "AnalysisCollection", "core", "lib", "mdshim", "parallel", "synthetic", *_modules_all]
