"""In-memory stand-in for the MDAnalysis objects the analysis classes touch (see ``universe.py``)."""

from .readers import TRRTrajectory, read_gro, read_trr, universe_from_gro, write_trr
from .universe import AtomGroup, MemoryTrajectory, Timestep, Universe

__all__ = ["AtomGroup", "MemoryTrajectory", "Timestep", "TRRTrajectory", "Universe", "read_gro", "read_trr", "universe_from_gro", "write_trr"]
