"""Analysis modules on the CUDA hot path (same class names as ``maicos.modules``)."""

from .saxs import Saxs
from .diporderstructurefactor import DiporderStructureFactor
from .densityplanar import DensityPlanar
from .densitycylinder import DensityCylinder
from .densitysphere import DensitySphere
from .diporderplanar import DiporderPlanar
from .dipordercylinder import DiporderCylinder
from .dipordersphere import DiporderSphere
from .velocityplanar import VelocityPlanar
from .velocitycylinder import VelocityCylinder
from .temperatureplanar import TemperaturePlanar
from .dielectricplanar import DielectricPlanar
from .dielectriccylinder import DielectricCylinder
from .dielectricsphere import DielectricSphere

__all__ = ['Saxs', 'DiporderStructureFactor', 'DensityPlanar', 'DensityCylinder', 'DensitySphere', 'DiporderPlanar', 'DiporderCylinder', 'DiporderSphere', 'VelocityPlanar', 'VelocityCylinder', 'TemperaturePlanar', 'DielectricPlanar', 'DielectricCylinder', 'DielectricSphere']
