"""Hand-written GROMACS ``.gro`` and ``.trr`` readers (orthorhombic cells, lengths converted nm -> Å).

These stand in for the MDAnalysis readers the reference relies on (third-party, not installable
here).  They cover what configuration 1 needs: the topology order of ``water.gro`` and the
101-frame ``water.trr`` (XDR big-endian, float32 or float64, x / v / f blocks).
"""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

from .universe import MemoryTrajectory, Universe

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


def universe_from_gro(gro_path, trr_path=None, charges=None) -> Universe:
    """Universe with the topology order of a ``.gro`` file and, optionally, a ``.trr`` trajectory."""
    g = read_gro(gro_path)
    n = len(g["names"])
    elements = np.array([guess_element(nm) for nm in g["names"]], dtype=object)
    masses = np.array([MASSES.get(el.upper(), 0.0) for el in elements])
    resindices = np.concatenate([[0], np.cumsum(np.diff(g["resids"]) != 0)])
    if trr_path is None:
        traj = MemoryTrajectory(g["positions"][None], g["dimensions"],
                                None if g["velocities"] is None else g["velocities"][None])
    else:
        t = read_trr(trr_path)
        if t["positions"].shape[1] != n:
            raise ValueError("topology and trajectory disagree on the number of atoms")
        traj = MemoryTrajectory(t["positions"], t["dimensions"], t["velocities"], times=t["times"])
    return Universe(n, traj, names=g["names"], types=g["names"], elements=elements, masses=masses,
                    charges=charges, resindices=resindices, resnames=g["resnames"], molnums=resindices)
