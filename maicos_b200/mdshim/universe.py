"""Minimal in-memory Universe / AtomGroup / trajectory with the MDAnalysis surface the hot path uses.

MDAnalysis is a hard dependency of the reference (``pyproject.toml:61-64``) but it is not
installable in the build image, so the analysis classes of this package are written against the
*duck-typed* subset listed in ``SURVEY.md`` section 8(b) and this module provides that subset for
orthorhombic cells:

``Universe.atoms / .dimensions / .trajectory``, ``AtomGroup.positions / velocities / n_atoms /
elements / names / types / masses / charges / resindices / molnums / select_atoms / center /
dipole_vector / total_mass / total_charge / accumulate / wrap / unwrap / translate`` and a
trajectory that yields ``Timestep`` objects with ``positions, dimensions, frame, time, volume,
copy()``.

Semantics follow the MDAnalysis 2.8 documentation (positions ``float32``, ``dimensions``
``float32[6]``, compound results ordered by ascending compound index, centres in ``float64``).
``unwrap`` makes a compound whole by the minimum image with respect to its first atom, which equals
the bond walk for compounds smaller than half the cell (all the water-like inputs used here).
When a real MDAnalysis AtomGroup is passed to the analysis classes this module is not involved.
"""

from __future__ import annotations

import numpy as np

_COMPOUND_ATTR = {
    "residues": "resindices",
    "molecules": "molnums",
    "segments": "segindices",
    "fragments": "fragindices",
}


class Timestep:
    """One trajectory frame: ``positions`` (float32 [n, 3], Å), ``dimensions`` (float32 [6])."""

    __slots__ = ("positions", "velocities", "dimensions", "frame", "time")

    def __init__(self, positions, dimensions, frame=0, time=0.0, velocities=None):
        self.positions = positions
        self.velocities = velocities
        self.dimensions = dimensions
        self.frame = frame
        self.time = time

    @property
    def volume(self) -> float:
        return float(np.prod(self.dimensions[:3].astype(np.float64)))

    def writable_positions(self) -> np.ndarray:
        """``positions`` for in-place modification.  Frames of a :class:`MemoryTrajectory` alias the trajectory's
        own storage (read-only view, no per-frame copy); the first transformation that writes takes a private copy."""
        if not self.positions.flags.writeable:
            self.positions = self.positions.copy()
        return self.positions

    @property
    def n_atoms(self) -> int:
        return len(self.positions)

    def copy(self) -> "Timestep":
        return Timestep(
            self.positions.copy(),
            None if self.dimensions is None else self.dimensions.copy(),
            self.frame,
            self.time,
            None if self.velocities is None else self.velocities.copy(),
        )


class _FrameIterator:
    """Iterable over selected frames of a :class:`MemoryTrajectory` (``trajectory[a:b:c]``)."""

    def __init__(self, trajectory, frames):
        self.trajectory = trajectory
        self.frames = np.asarray(frames, dtype=np.int64)

    def __len__(self):
        return len(self.frames)

    def __iter__(self):
        for f in self.frames:
            yield self.trajectory._goto(int(f))


class MemoryTrajectory:
    """Trajectory held in host memory: ``coordinates float32 [F, n, 3]``, per-frame or fixed box.

    ``frame_source`` (optional) is a callable ``f(frame_index) -> float32 [n, 3]`` used instead of
    a stored coordinate block, for synthetic trajectories that are generated on the fly.
    """

    def __init__(self, coordinates=None, dimensions=None, velocities=None, dt=1.0, times=None,
                 n_frames=None, frame_source=None, n_atoms=None, pinned=False, pinned_rows=None):
        """``pinned=True`` keeps the coordinate block in page-locked memory (one copy at construction): the analysis
        classes then hand frames to the device straight from the trajectory, without a staging copy.  A
        ``frame_source`` that serves rows of a page-locked block says so with ``pinned_rows(frame) -> (block, row)``."""
        self._pinned_block = None
        self._pinned_rows = pinned_rows
        if coordinates is not None:
            coordinates = np.asarray(coordinates, dtype=np.float32)
            if coordinates.ndim == 2:
                coordinates = coordinates[None]
            n_frames = coordinates.shape[0]
            n_atoms = coordinates.shape[1]
            if pinned:
                from .. import _cabi

                if _cabi.device_count() > 0:
                    self._pinned_block = _cabi._PinnedBlock(max(coordinates.nbytes, 1))
                    block = np.frombuffer(self._pinned_block.buf, dtype=np.float32, count=coordinates.size)
                    block = block.reshape(coordinates.shape)
                    np.copyto(block, coordinates)
                    coordinates = block
        elif frame_source is None or n_frames is None or n_atoms is None:
            raise ValueError("give either `coordinates` or `frame_source` with `n_frames` and `n_atoms`")
        self.coordinates = coordinates
        self.frame_source = frame_source
        self.n_frames = int(n_frames)
        self.n_atoms = int(n_atoms)
        if velocities is not None:
            velocities = np.asarray(velocities, dtype=np.float32)
            if velocities.ndim == 2:
                velocities = velocities[None]
        self.velocities = velocities
        if dimensions is not None:
            dimensions = np.asarray(dimensions, dtype=np.float32)
            if dimensions.shape[-1] == 3:
                dimensions = np.concatenate([dimensions, np.full(dimensions.shape[:-1] + (3,), 90, np.float32)], -1)
            if dimensions.ndim == 1:
                dimensions = np.broadcast_to(dimensions, (self.n_frames, 6))
        self.dimensions = dimensions
        self.dt = float(dt)
        self.times = None if times is None else np.asarray(times, dtype=np.float64)
        self.ts = None
        self._goto(0)

    def _goto(self, f: int) -> Timestep:
        if f < 0:
            f += self.n_frames
        if not 0 <= f < self.n_frames:
            raise IndexError(f"frame {f} out of range (0..{self.n_frames - 1})")
        if self.coordinates is None:
            pos = self.frame_source(f)
        else:  # copy-on-write: a read-only view until something modifies the frame (Timestep.writable_positions)
            pos = self.coordinates[f]
            if pos.flags.c_contiguous:
                pos = pos.view()
                pos.flags.writeable = False
            else:
                pos = pos.copy()
        vel = None if self.velocities is None else self.velocities[f].copy()
        dims = None if self.dimensions is None else np.array(self.dimensions[f], dtype=np.float32)
        t = self.times[f] if self.times is not None else f * self.dt
        self.ts = Timestep(np.ascontiguousarray(pos, dtype=np.float32), dims, f, float(t), vel)
        return self.ts

    @property
    def has_pinned_frames(self) -> bool:
        """True when the frames live in page-locked memory (``pinned_row`` can answer)."""
        return self._pinned_rows is not None or self._pinned_block is not None

    def pinned_row(self, f: int):
        """``(page-locked float32 block [*, n_atoms, 3], row)`` holding frame ``f``, or ``None``."""
        if self._pinned_rows is not None:
            return self._pinned_rows(int(f))
        if self._pinned_block is not None:
            return self.coordinates, int(f)
        return None

    def __del__(self):
        block, self._pinned_block = getattr(self, "_pinned_block", None), None
        if block is not None:
            self.coordinates = None
            try:
                block.free()
            except Exception:
                pass

    def __len__(self):
        return self.n_frames

    def __iter__(self):
        for f in range(self.n_frames):
            yield self._goto(f)

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            return self._goto(int(item))
        if isinstance(item, slice):
            return _FrameIterator(self, np.arange(self.n_frames)[item])
        item = np.asarray(item)
        if item.dtype == bool:
            item = np.nonzero(item)[0]
        return _FrameIterator(self, item)


class Universe:
    """Topology arrays + a trajectory.  All per-atom arrays have length ``n_atoms``."""

    def __init__(self, n_atoms, trajectory=None, names=None, types=None, elements=None, masses=None,
                 charges=None, resindices=None, resnames=None, molnums=None, segindices=None, bonds=None):
        self.n_atoms = int(n_atoms)
        n = self.n_atoms

        def arr(x, dtype, default):
            if x is None:
                x = default
            a = np.asarray(x, dtype=dtype)
            if a.shape != (n,):
                raise ValueError(f"per-atom attribute has shape {a.shape}, expected ({n},)")
            return a

        self._attrs = {}
        if names is not None:
            self._attrs["names"] = arr(names, object, None)
        if types is not None or names is not None:
            self._attrs["types"] = arr(types if types is not None else names, object, None)
        if elements is not None:
            self._attrs["elements"] = arr(elements, object, None)
        if masses is not None:
            self._attrs["masses"] = arr(masses, np.float64, None)
        if charges is not None:
            self._attrs["charges"] = arr(charges, np.float64, None)
        self._attrs["resindices"] = arr(resindices, np.int64, np.arange(n))
        if resnames is not None:
            self._attrs["resnames"] = arr(resnames, object, None)
        self._attrs["molnums"] = arr(molnums, np.int64, self._attrs["resindices"])
        self._attrs["segindices"] = arr(segindices, np.int64, np.zeros(n, np.int64))
        self.bonds = None if bonds is None else np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        self._attrs["fragindices"] = self._fragments()
        self.trajectory = trajectory if trajectory is not None else MemoryTrajectory(np.zeros((1, n, 3), np.float32))
        if self.trajectory.n_atoms != n:
            raise ValueError("trajectory and topology disagree on the number of atoms")
        self.atoms = AtomGroup(self, np.arange(n, dtype=np.int64))

    def _fragments(self) -> np.ndarray:
        n = self.n_atoms
        if self.bonds is None or len(self.bonds) == 0:
            return self._attrs["molnums"].copy() if self.bonds is None else np.arange(n)
        parent = np.arange(n)

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i

        for a, b in self.bonds:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        roots = np.array([find(i) for i in range(n)])
        return np.unique(roots, return_inverse=True)[1]

    @property
    def dimensions(self):
        return self.trajectory.ts.dimensions

    @dimensions.setter
    def dimensions(self, box):
        box = np.asarray(box, dtype=np.float32)
        if box.shape == (3,):
            box = np.concatenate([box, np.full(3, 90, np.float32)])
        self.trajectory.ts.dimensions = box
        if self.trajectory.dimensions is None:
            self.trajectory.dimensions = np.broadcast_to(box, (self.trajectory.n_frames, 6))

    def select_atoms(self, selection: str) -> "AtomGroup":
        return self.atoms.select_atoms(selection)


class AtomGroup:
    """Ordered set of atom indices into a :class:`Universe`."""

    def __init__(self, universe: Universe, indices: np.ndarray):
        self.universe = universe
        self.indices = np.asarray(indices, dtype=np.int64)
        self._all = len(self.indices) == universe.n_atoms and np.array_equal(self.indices, np.arange(universe.n_atoms))

    # ---- container protocol ---------------------------------------------------------------
    @property
    def atoms(self) -> "AtomGroup":
        return self

    @property
    def n_atoms(self) -> int:
        return len(self.indices)

    @property
    def ix(self) -> np.ndarray:
        """Topology indices of the atoms (MDAnalysis ``AtomGroup.ix``)."""
        return self.indices

    def __len__(self) -> int:
        return len(self.indices)

    def __getitem__(self, item) -> "AtomGroup":
        return AtomGroup(self.universe, np.atleast_1d(self.indices[item]))

    def __add__(self, other: "AtomGroup") -> "AtomGroup":
        return AtomGroup(self.universe, np.concatenate([self.indices, other.indices]))

    def __getattr__(self, name):
        # per-atom topology attributes; absent ones raise AttributeError like MDAnalysis
        attrs = self.__dict__.get("universe")._attrs if "universe" in self.__dict__ else {}
        if name in attrs:
            if self.__dict__.get("_all"):  # the whole universe: a read-only view instead of a gathered copy
                view = attrs[name].view()
                view.flags.writeable = False
                return view
            return attrs[name][self.indices]
        raise AttributeError(f"AtomGroup has no attribute {name!r}")

    @property
    def n_residues(self) -> int:
        return len(np.unique(self.resindices))

    @property
    def n_segments(self) -> int:
        return len(np.unique(self.segindices))

    @property
    def n_fragments(self) -> int:
        return len(np.unique(self.fragindices))

    # ---- coordinates ----------------------------------------------------------------------
    @property
    def positions(self) -> np.ndarray:
        ts = self.universe.trajectory.ts
        return ts.positions.copy() if self._all else ts.positions[self.indices]

    @positions.setter
    def positions(self, value):
        self.universe.trajectory.ts.writable_positions()[self.indices] = value

    @property
    def velocities(self) -> np.ndarray:
        ts = self.universe.trajectory.ts
        if ts.velocities is None:
            raise AttributeError("This Timestep has no velocities")
        return ts.velocities[self.indices]

    @property
    def dimensions(self):
        return self.universe.dimensions

    def translate(self, t) -> "AtomGroup":
        ts = self.universe.trajectory.ts
        t = np.asarray(t, dtype=np.float32)
        if self._all:
            ts.writable_positions()[...] += t
        else:
            ts.writable_positions()[self.indices] += t
        return self

    # ---- selections -----------------------------------------------------------------------
    def select_atoms(self, selection: str) -> "AtomGroup":
        """Tiny subset of the MDAnalysis selection language.

        ``all``; ``element|name|type|resname V [V ...]``; ``index|resid|resindex|molnum a[:b]`` (inclusive
        ranges, ``resid`` = resindex + 1); ``not``; ``and`` / ``or`` (``and`` binds tighter).
        """
        mask = np.zeros(len(self), dtype=bool)
        for clause in selection.split(" or "):
            m = np.ones(len(self), dtype=bool)
            for term in clause.split(" and "):
                m &= self._term(term.split())
            mask |= m
        return AtomGroup(self.universe, self.indices[mask])

    def _term(self, tok) -> np.ndarray:
        if not tok:
            raise ValueError("empty selection")
        if tok[0] == "not":
            return ~self._term(tok[1:])
        key, vals = tok[0], tok[1:]
        if key == "all":
            return np.ones(len(self), dtype=bool)
        text = {"element": "elements", "name": "names", "type": "types", "resname": "resnames"}
        if key in text:
            return np.isin(getattr(self, text[key]).astype(str), vals)
        nums = {"index": (self.indices, 0), "resindex": (None, 0), "resid": (None, 1), "molnum": (None, 0)}
        if key in nums:
            data = self.indices if key == "index" else (self.molnums if key == "molnum" else self.resindices + nums[key][1])
            m = np.zeros(len(self), dtype=bool)
            for v in vals:
                lo, _, hi = v.partition(":")
                m |= (data >= int(lo)) & (data <= int(hi or lo))
            return m
        raise ValueError(f"selection keyword {key!r} is not supported by the in-memory shim")

    # ---- compound reductions --------------------------------------------------------------
    def _compound_ids(self, compound: str) -> np.ndarray:
        """Dense 0..n_c-1 compound id per atom, numbered by ascending topology index."""
        c = compound.lower()
        if c == "group":
            return np.zeros(len(self), dtype=np.int64)
        if c not in _COMPOUND_ATTR:
            raise ValueError(f"Unrecognized compound definition: {compound}")
        return np.unique(getattr(self, _COMPOUND_ATTR[c]), return_inverse=True)[1]

    def accumulate(self, attribute, compound: str = "group"):
        values = getattr(self, attribute) if isinstance(attribute, str) else np.asarray(attribute)
        ids = self._compound_ids(compound)
        n_c = int(ids.max()) + 1 if len(ids) else 0
        values = np.asarray(values, dtype=np.float64)
        if values.ndim == 1:
            out = np.bincount(ids, weights=values, minlength=n_c)
        else:
            out = np.stack([np.bincount(ids, weights=values[:, d], minlength=n_c) for d in range(values.shape[1])], 1)
        return out[0] if compound.lower() == "group" else out

    def total_mass(self, compound: str = "group"):
        return self.accumulate("masses", compound)

    def total_charge(self, compound: str = "group"):
        return self.accumulate("charges", compound)

    def center(self, weights, compound: str = "group") -> np.ndarray:
        """Per-compound centre ``sum(w r) / sum(w)`` in float64 (plain mean for ``weights=None``)."""
        pos = self.positions.astype(np.float64)
        w = np.ones(len(self)) if weights is None else np.asarray(weights, dtype=np.float64)
        ids = self._compound_ids(compound)
        n_c = int(ids.max()) + 1 if len(ids) else 0
        wsum = np.bincount(ids, weights=w, minlength=n_c)
        num = np.stack([np.bincount(ids, weights=w * pos[:, d], minlength=n_c) for d in range(3)], 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            out = num / wsum[:, None]
        return out[0] if compound.lower() == "group" else out

    def center_of_mass(self, compound: str = "group") -> np.ndarray:
        return self.center(self.masses, compound)

    def center_of_geometry(self, compound: str = "group") -> np.ndarray:
        return self.center(None, compound)

    def dipole_vector(self, compound: str = "group") -> np.ndarray:
        """``sum_i q_i (r_i - R_com)`` per compound (e Å)."""
        pos = self.positions.astype(np.float64)
        ids = self._compound_ids(compound)
        com = np.atleast_2d(self.center(self.masses, compound))
        rel = pos - com[ids]
        out = self.accumulate(self.charges[:, None] * rel, compound)
        return out

    # ---- periodic boundary handling (orthorhombic) ----------------------------------------
    def _box(self) -> np.ndarray:
        dims = self.universe.dimensions
        if dims is None:
            raise ValueError("Universe has no dimensions")
        if len(dims) >= 6 and not np.all(np.asarray(dims[3:6]) == 90.0):
            raise NotImplementedError(
                "the in-memory shim wraps / unwraps orthorhombic cells only; use pack=False, unwrap=False or an "
                "MDAnalysis universe for triclinic cells")
        return dims[:3].astype(np.float32)

    def wrap(self, compound: str = "atoms") -> np.ndarray:
        """Move atoms (or whole compounds, by their centre of mass) into the primary cell, in place."""
        box = self._box()
        ts = self.universe.trajectory.ts
        pos = ts.writable_positions() if self._all else ts.positions[self.indices]
        if compound.lower() == "atoms":
            pos -= np.floor(pos / box) * box
        else:
            ids = self._compound_ids(compound)
            masses = self.masses if "masses" in self.universe._attrs else None
            com = np.atleast_2d(self.center(masses, compound))
            shift = (-np.floor(com / box.astype(np.float64)) * box).astype(np.float32)
            pos += shift[ids]
        if not self._all:
            ts.writable_positions()[self.indices] = pos
        return pos

    def unwrap(self, compound: str = "fragments") -> np.ndarray:
        """Make compounds whole (minimum image about their first atom) and put their COM in the cell."""
        box = self._box()
        ts = self.universe.trajectory.ts
        pos = ts.writable_positions() if self._all else ts.positions[self.indices]
        ids = self._compound_ids(compound)
        first = np.full(int(ids.max()) + 1 if len(ids) else 0, -1, dtype=np.int64)
        order = np.arange(len(ids))[::-1]
        first[ids[order]] = order  # lowest atom position of every compound
        anchor = pos[first[ids]]
        delta = pos - anchor
        delta -= box * np.round(delta / box)
        pos[...] = anchor + delta
        if not self._all:
            ts.writable_positions()[self.indices] = pos
        return self.wrap(compound=compound)
