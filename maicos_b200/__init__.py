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

__all__ = ["AnalysisBase", "AnalysisCollection", "core", "lib", "mdshim", "parallel", "synthetic", *_modules_all]
