"""Frame loop, statistics and geometry bases (same public names as ``maicos.core``)."""

from .base import AnalysisBase, AnalysisCollection, ProfileBase, Results
from .cylinder import CylinderBase, ProfileCylinderBase
from .planar import PlanarBase, ProfilePlanarBase
from .sphere import ProfileSphereBase, SphereBase

__all__ = [
    "AnalysisBase", "AnalysisCollection", "ProfileBase", "Results", "PlanarBase", "ProfilePlanarBase",
    "CylinderBase", "ProfileCylinderBase", "SphereBase", "ProfileSphereBase",
]
