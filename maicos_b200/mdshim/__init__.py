"""In-memory stand-in for the MDAnalysis objects the analysis classes touch (see ``universe.py``)."""

from .readers import TRRTrajectory, XTCTrajectory, read_gro, read_trr, read_xtc, universe_from_gro, write_trr
from .universe import AtomGroup, MemoryTrajectory, Timestep, Universe

__all__ = ["AtomGroup", "MemoryTrajectory", "Timestep", "TRRTrajectory", "Universe", "XTCTrajectory",
           "read_gro", "read_trr", "read_xtc", "universe_from_gro", "write_trr"]
