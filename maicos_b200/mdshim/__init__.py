"""In-memory stand-in for the MDAnalysis objects the analysis classes touch (see ``universe.py``)."""

from .readers import read_gro, read_trr, universe_from_gro
from .universe import AtomGroup, MemoryTrajectory, Timestep, Universe

__all__ = ["AtomGroup", "MemoryTrajectory", "Timestep", "Universe", "read_gro", "read_trr", "universe_from_gro"]
