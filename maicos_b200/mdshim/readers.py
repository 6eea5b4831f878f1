"""Hand-written GROMACS ``.gro``, ``.trr`` and ``.xtc`` readers (lengths converted nm -> Å).

These stand in for the MDAnalysis readers the reference relies on (third-party, not installable
here).  They cover what configuration 1 needs: the topology order of ``water.gro`` and the
101-frame ``water.trr`` (XDR big-endian, float32 or float64, x / v / f blocks).
"""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

from .universe import MemoryTrajectory, Timestep, Universe

#: atomic masses (u) for elements met in the water / slab inputs
MASSES = {"H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999, "NA": 22.98977, "CL": 35.45, "S": 32.06}


def guess_element(name: str) -> str:
    """Element symbol from a GROMACS atom name (``OW`` -> ``O``, ``HW1`` -> ``H``, ``CL`` -> ``Cl``)."""
    letters = "".join(ch for ch in name if ch.isalpha()).upper()
    if letters[:2] in ("CL", "NA", "MG", "ZN", "FE", "CA") and len(letters) <= 3 and not letters.startswith("CA1"):
        if letters[:2] in ("CL", "NA", "MG", "ZN", "FE"):
            return letters[:2].title()
    return letters[:1]


def read_gro(path) -> dict:
    """Parse a ``.gro`` file -> dict(names, resnames, resids, positions [n,3] f32 Å, velocities, dimensions)."""
    lines = Path(path).read_text().splitlines()
    n = int(lines[1])
    resid = np.empty(n, dtype=np.int64)
    resname, name = [], []
    pos = np.empty((n, 3), dtype=np.float64)
    vel = None
    for i in range(n):
        ln = lines[2 + i]
        resid[i] = int(ln[0:5])
        resname.append(ln[5:10].strip())
        name.append(ln[10:15].strip())
        rest = ln[20:].split()
        pos[i] = [float(rest[0]), float(rest[1]), float(rest[2])]
        if len(rest) >= 6:
            if vel is None:
                vel = np.zeros((n, 3), dtype=np.float64)
            vel[i] = [float(rest[3]), float(rest[4]), float(rest[5])]
    box = [float(x) for x in lines[2 + n].split()][:3]
    dims = np.array([box[0] * 10, box[1] * 10, box[2] * 10, 90, 90, 90], dtype=np.float32)
    return {
        "names": np.array(name, dtype=object),
        "resnames": np.array(resname, dtype=object),
        "resids": resid,
        "positions": (pos * 10).astype(np.float32),
        "velocities": None if vel is None else (vel * 10).astype(np.float32),
        "dimensions": dims,
    }


def read_trr(path) -> dict:
    """Parse every frame of a ``.trr`` -> dict(positions [F,n,3] f32 Å, velocities, dimensions [F,6], times)."""
    data = Path(path).read_bytes()
    off = 0
    frames_x, frames_v, dims, times = [], [], [], []
    while off < len(data):
        magic, slen = struct.unpack_from(">ii", data, off)
        if magic != 1993:
            raise ValueError(f"not a TRR frame at byte {off} (magic {magic})")
        off += 8
        (strlen,) = struct.unpack_from(">i", data, off)
        off += 4 + (strlen + 3) // 4 * 4
        ir, e, box, vir, pres, top, sym, xs, vs, fs, natoms, step, nre = struct.unpack_from(">13i", data, off)
        off += 52
        if box:
            real = box // 9
        elif xs:
            real = xs // (3 * natoms)
        else:
            real = (vs or fs) // (3 * natoms)
        fmt = ">f4" if real == 4 else ">f8"
        t, _lam = np.frombuffer(data, dtype=fmt, count=2, offset=off)
        off += 2 * real
        if box:
            b = np.frombuffer(data, dtype=fmt, count=9, offset=off).reshape(3, 3)
            dims.append([b[0, 0] * 10, b[1, 1] * 10, b[2, 2] * 10, 90, 90, 90])
        off += box + vir + pres
        if xs:
            frames_x.append(np.frombuffer(data, dtype=fmt, count=3 * natoms, offset=off).reshape(natoms, 3) * 10)
        off += xs
        if vs:
            frames_v.append(np.frombuffer(data, dtype=fmt, count=3 * natoms, offset=off).reshape(natoms, 3) * 10)
        off += vs + fs
        times.append(float(t))
    return {
        "positions": np.asarray(frames_x, dtype=np.float32),
        "velocities": np.asarray(frames_v, dtype=np.float32) if len(frames_v) == len(frames_x) else None,
        "dimensions": np.asarray(dims, dtype=np.float32),
        "times": np.asarray(times),
    }


def _scan_trr(mm) -> list[dict]:
    """Frame index of a TRR file from its headers alone: byte offsets of the x / v blocks, box, time, precision."""
    frames, off, size = [], 0, len(mm)
    while off < size:
        magic, _slen = struct.unpack_from(">ii", mm, off)
        if magic != 1993:
            raise ValueError(f"not a TRR frame at byte {off} (magic {magic})")
        off += 8
        (strlen,) = struct.unpack_from(">i", mm, off)
        off += 4 + (strlen + 3) // 4 * 4
        _ir, _e, box, vir, pres, _top, _sym, xs, vs, fs, natoms, _step, _nre = struct.unpack_from(">13i", mm, off)
        off += 52
        real = box // 9 if box else (xs // (3 * natoms) if xs else (vs or fs) // (3 * natoms))
        fmt = ">f4" if real == 4 else ">f8"
        t = float(np.frombuffer(mm, dtype=fmt, count=1, offset=off)[0])
        off += 2 * real
        dims = None
        if box:
            b = np.frombuffer(mm, dtype=fmt, count=9, offset=off).reshape(3, 3)
            dims = [b[0, 0] * 10, b[1, 1] * 10, b[2, 2] * 10, 90, 90, 90]
        off += box + vir + pres
        frames.append({"x": off if xs else None, "v": off + xs if vs else None, "n": natoms, "fmt": fmt, "time": t, "dims": dims})
        off += xs + vs + fs
    return frames


class _StreamingTrajectory:
    """A trajectory file that is memory-mapped and indexed once (headers only); a frame is decoded when it is visited.

    ``pinned_ring=k`` decodes into a ring of ``k`` page-locked frame buffers and serves those rows through
    ``pinned_row`` (the protocol of :class:`MemoryTrajectory`): the batched analyses then hand the decoded frames to the
    device where the reader put them, without a staging copy.  An analysis keeps at most ``(devices + 2) x batch``
    frames alive; ``AnalysisBase`` shrinks its batch to fit the ring.

    Subclasses provide ``_scan(mm) -> [dict(n, time, dims, v)]`` and ``_decode_x(frame index entry, out)`` (and
    ``_decode_v`` when the format carries velocities).
    """

    def __init__(self, path, pinned_ring: int = 0):
        import mmap

        self._file = open(path, "rb")
        self._mm = mmap.mmap(self._file.fileno(), 0, access=mmap.ACCESS_READ)
        self._index = self._scan(self._mm)
        if not self._index:
            raise ValueError("no frames in the trajectory")
        self.n_frames = len(self._index)
        self.n_atoms = int(self._index[0]["n"])
        if any(fr["n"] != self.n_atoms for fr in self._index):
            raise ValueError("the number of atoms changes inside the trajectory")
        have_box = all(fr["dims"] is not None for fr in self._index)
        self.dimensions = np.asarray([fr["dims"] for fr in self._index], dtype=np.float32) if have_box else None
        self.times = np.asarray([fr["time"] for fr in self._index])
        self.velocities = True if all(fr.get("v") is not None for fr in self._index) else None  # decoded per frame
        self.dt = float(self.times[1] - self.times[0]) if self.n_frames > 1 else 1.0
        self.coordinates = None
        self.ring_len = 0
        self._ring = self._ring_block = None
        self._ring_frame: list[int] = []
        if pinned_ring > 0:
            from .. import _cabi

            self.ring_len = int(min(pinned_ring, self.n_frames))
            nbytes = self.ring_len * self.n_atoms * 12
            if _cabi.device_count() > 0:
                self._ring_block = _cabi._PinnedBlock(max(nbytes, 1))
                self._ring = np.frombuffer(self._ring_block.buf, dtype=np.float32, count=self.ring_len * self.n_atoms * 3)
                self._ring = self._ring.reshape(self.ring_len, self.n_atoms, 3)
            else:
                self._ring = np.empty((self.ring_len, self.n_atoms, 3), dtype=np.float32)
            self._ring_frame = [-1] * self.ring_len
        self.ts = None
        self._goto(0)

    def _scan(self, mm) -> list[dict]:
        raise NotImplementedError

    def _decode_x(self, fr: dict, out: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def _decode_v(self, fr: dict, out: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def _goto(self, f: int) -> Timestep:
        if f < 0:
            f += self.n_frames
        if not 0 <= f < self.n_frames:
            raise IndexError(f"frame {f} out of range (0..{self.n_frames - 1})")
        fr = self._index[f]
        if self._ring is not None:
            slot = f % self.ring_len
            if self._ring_frame[slot] != f:
                self._ring_frame[slot] = -1
                self._decode_x(fr, self._ring[slot])
                self._ring_frame[slot] = f
            pos = self._ring[slot].view()
        else:
            pos = self._decode_x(fr, np.empty((self.n_atoms, 3), dtype=np.float32))
        pos.flags.writeable = False  # copy-on-write, like MemoryTrajectory (Timestep.writable_positions)
        vel = None
        if fr.get("v") is not None and self.velocities is not None:
            vel = self._decode_v(fr, np.empty((self.n_atoms, 3), dtype=np.float32))
        dims = None if self.dimensions is None else np.array(self.dimensions[f], dtype=np.float32)
        self.ts = Timestep(pos, dims, f, float(fr["time"]), vel)
        return self.ts

    @property
    def has_pinned_frames(self) -> bool:
        """True when frames are decoded into a ring of page-locked buffers (``pinned_ring=k``)."""
        return self._ring_block is not None

    def pinned_row(self, f: int):
        """``(ring block, row)`` when frame ``f`` is the one currently held by its ring slot, else ``None``."""
        if self._ring_block is None:
            return None
        slot = int(f) % self.ring_len
        return (self._ring, slot) if self._ring_frame[slot] == int(f) else None

    def __len__(self):
        return self.n_frames

    def __iter__(self):
        for f in range(self.n_frames):
            yield self._goto(f)

    def __getitem__(self, item):
        from .universe import _FrameIterator

        if isinstance(item, (int, np.integer)):
            return self._goto(int(item))
        if isinstance(item, slice):
            return _FrameIterator(self, np.arange(self.n_frames)[item])
        item = np.asarray(item)
        if item.dtype == bool:
            item = np.nonzero(item)[0]
        return _FrameIterator(self, item)

    def close(self) -> None:
        block, self._ring_block = self._ring_block, None
        self._ring = None
        if block is not None:
            block.free()
        if getattr(self, "_mm", None) is not None:
            self.ts = None
            try:
                self._mm.close()
            except BufferError:  # a decoded view is still alive somewhere: the map goes with the last reference
                pass
            self._mm = None
            self._file.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TRRTrajectory(_StreamingTrajectory):
    """Streaming ``.trr`` reader: big-endian XDR reals -> native float32, nm -> Å, ONE pass over the frame's bytes."""

    def _scan(self, mm) -> list[dict]:
        index = _scan_trr(mm)
        if any(fr["x"] is None for fr in index):
            raise ValueError("every TRR frame needs coordinates")
        return index

    def _decode(self, offset: int, fmt: str, out: np.ndarray) -> np.ndarray:
        src = np.frombuffer(self._mm, dtype=fmt, count=3 * self.n_atoms, offset=offset).reshape(self.n_atoms, 3)
        np.multiply(src, 10.0, out=out, casting="unsafe")  # byte order, precision and unit in one pass
        return out

    def _decode_x(self, fr: dict, out: np.ndarray) -> np.ndarray:
        return self._decode(fr["x"], fr["fmt"], out)

    def _decode_v(self, fr: dict, out: np.ndarray) -> np.ndarray:
        return self._decode(fr["v"], fr["fmt"], out)


def box_to_dimensions(box: np.ndarray) -> np.ndarray:
    """``[lx, ly, lz, alpha, beta, gamma]`` (Å, degrees; float32) of a 3 x 3 matrix of cell vectors in nm."""
    box = np.asarray(box, dtype=np.float64).reshape(3, 3) * 10.0
    lengths = np.linalg.norm(box, axis=1)
    if not np.all(lengths > 0):
        return np.zeros(6, dtype=np.float32)
    a, b, c = box

    def angle(u, v):
        return float(np.degrees(np.arccos(np.clip(np.dot(u, v) / (np.linalg.norm(u) * np.linalg.norm(v)), -1.0, 1.0))))

    off = box - np.diag(np.diag(box))
    angles = [90.0, 90.0, 90.0] if not off.any() else [angle(b, c), angle(a, c), angle(a, b)]
    return np.array([*lengths, *angles], dtype=np.float32)


class XTCTrajectory(_StreamingTrajectory):
    """Streaming ``.xtc`` reader (compressed coordinates, no velocities): frames are indexed from their headers and decoded
    by ``mb200_xtc_frame`` (C++ in ``libmaicos_b200``, host only) when they are visited, nm -> Å."""

    def _scan(self, mm) -> list[dict]:
        import ctypes

        from .. import _cabi

        lib = _cabi.load()
        self._bytes = np.frombuffer(mm, dtype=np.uint8)
        size, off, index = len(self._bytes), 0, []
        nat, step, t, nxt = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_float(), ctypes.c_size_t()
        box = np.zeros(9, dtype=np.float32)
        while off < size:
            _cabi.check(lib.mb200_xtc_frame(_cabi.ptr(self._bytes), size, off, ctypes.byref(nat), ctypes.byref(step),
                                            ctypes.byref(t), _cabi.ptr(box), None, 10.0, ctypes.byref(nxt)))
            index.append({"offset": off, "n": int(nat.value), "time": float(t.value), "step": int(step.value),
                          "dims": box_to_dimensions(box), "v": None})
            off = int(nxt.value)
        return index

    def _decode_x(self, fr: dict, out: np.ndarray) -> np.ndarray:
        from .. import _cabi

        assert out.flags.c_contiguous and out.dtype == np.float32
        _cabi.check(_cabi.load().mb200_xtc_frame(_cabi.ptr(self._bytes), len(self._bytes), fr["offset"], None, None, None, None,
                                                 _cabi.ptr(out), 10.0, None))
        return out

    def close(self) -> None:
        self._bytes = None
        super().close()


def write_trr(path, positions, dimensions, times=None, velocities=None) -> None:
    """Write float32 frames (Å) as a single-precision GROMACS ``.trr`` (orthorhombic box, x and optionally v blocks): the
    inverse of the readers above, used to put synthetic trajectories on disk."""
    positions = np.asarray(positions, dtype=np.float32)
    n_frames, n_atoms = positions.shape[:2]
    dimensions = np.broadcast_to(np.asarray(dimensions, dtype=np.float32), (n_frames, np.shape(dimensions)[-1]))
    version = b"GMX_trn_file"
    with open(path, "wb") as fh:
        for f in range(n_frames):
            xs = 12 * n_atoms
            vs = xs if velocities is not None else 0
            fh.write(struct.pack(">ii", 1993, len(version) + 1))
            fh.write(struct.pack(">i", len(version)) + version + b"\0" * (-len(version) % 4))
            fh.write(struct.pack(">13i", 0, 0, 36, 0, 0, 0, 0, xs, vs, 0, n_atoms, f, 0))
            fh.write(struct.pack(">ff", float(times[f]) if times is not None else float(f), 0.0))
            box = np.zeros((3, 3), dtype=">f4")
            box[0, 0], box[1, 1], box[2, 2] = dimensions[f, 0] / 10, dimensions[f, 1] / 10, dimensions[f, 2] / 10
            fh.write(box.tobytes())
            fh.write((positions[f] / np.float32(10)).astype(">f4").tobytes())
            if velocities is not None:
                fh.write((np.asarray(velocities[f], dtype=np.float32) / np.float32(10)).astype(">f4").tobytes())


def read_xtc(path) -> dict:
    """Decode every frame of an ``.xtc`` -> dict(positions [F,n,3] f32 Å, velocities None, dimensions [F,6], times)."""
    traj = XTCTrajectory(path)
    try:
        pos = np.empty((traj.n_frames, traj.n_atoms, 3), dtype=np.float32)
        for f, fr in enumerate(traj._index):
            traj._decode_x(fr, pos[f])
        return {"positions": pos, "velocities": None, "dimensions": None if traj.dimensions is None else traj.dimensions.copy(),
                "times": traj.times.copy()}
    finally:
        traj.close()


def universe_from_gro(gro_path, trr_path=None, charges=None, stream: bool = False, pinned_ring: int = 0) -> Universe:
    """Universe with the topology order of a ``.gro`` file and, optionally, a ``.trr`` or ``.xtc`` trajectory."""
    g = read_gro(gro_path)
    n = len(g["names"])
    elements = np.array([guess_element(nm) for nm in g["names"]], dtype=object)
    masses = np.array([MASSES.get(el.upper(), 0.0) for el in elements])
    resindices = np.concatenate([[0], np.cumsum(np.diff(g["resids"]) != 0)])
    if trr_path is None:
        traj = MemoryTrajectory(g["positions"][None], g["dimensions"],
                                None if g["velocities"] is None else g["velocities"][None])
    elif stream or pinned_ring:  # frames are decoded from the memory-mapped file when they are visited
        reader = XTCTrajectory if str(trr_path).lower().endswith(".xtc") else TRRTrajectory
        traj = reader(trr_path, pinned_ring=pinned_ring)
        if traj.n_atoms != n:
            raise ValueError("topology and trajectory disagree on the number of atoms")
    else:
        t = read_xtc(trr_path) if str(trr_path).lower().endswith(".xtc") else read_trr(trr_path)
        if t["positions"].shape[1] != n:
            raise ValueError("topology and trajectory disagree on the number of atoms")
        traj = MemoryTrajectory(t["positions"], t["dimensions"], t["velocities"], times=t["times"])
    return Universe(n, traj, names=g["names"], types=g["names"], elements=elements, masses=masses,
                    charges=charges, resindices=resindices, resnames=g["resnames"], molnums=resindices)
