"""Seeded synthetic inputs of the benchmark configurations (``BASELINE.json`` / ``SURVEY.md`` 8(d)).

Positions are float32 Å like MDAnalysis stores them.  Molecules are whole and their atoms lie in the
primary cell for the planar slab; for bulk water a few hydrogens may stick out of ``[0, L)`` (the
``Saxs`` pack step handles that, as in the reference).
"""

from __future__ import annotations

import numpy as np

#: SPC/E geometry and parameters
OH, HOH = 1.0, np.deg2rad(109.47)
Q_O, Q_H = -0.8476, 0.4238
M_O, M_H = 15.9994, 1.008
#: 33.37 molecules / nm^3
RHO_MOL = 33.37e-3  # per Å^3


def cubic_box_for(n_mol: int) -> float:
    """Edge (Å) of the cubic cell holding ``n_mol`` SPC/E molecules at liquid density."""
    return float((n_mol / RHO_MOL) ** (1.0 / 3.0))


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    a, b, c, d = q.T
    return np.stack([
        np.stack([a*a + b*b - c*c - d*d, 2*(b*c - a*d), 2*(b*d + a*c)], -1),
        np.stack([2*(b*c + a*d), a*a - b*b + c*c - d*d, 2*(c*d - a*b)], -1),
        np.stack([2*(b*d - a*c), 2*(c*d + a*b), a*a - b*b - c*c + d*d], -1),
    ], 1)


def spce_molecules(n_mol: int, box, seed: int, zlo: float | None = None, zhi: float | None = None):
    """Molecule centres (oxygen sites) on a jittered lattice + rigid SPC/E bodies.

    Returns ``(oxygen [n_mol, 3], h1_rel, h2_rel)`` in float64; ``zlo / zhi`` confine the oxygens
    along z (slab)."""
    rng = np.random.default_rng(seed)
    box = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,)).copy()
    lo = np.zeros(3)
    ext = box.copy()
    if zlo is not None:
        lo[2], ext[2] = zlo, zhi - zlo
    vol = np.prod(ext)
    a = (vol / n_mol) ** (1.0 / 3.0)
    cells = np.maximum(1, np.ceil(ext / a - 1e-9).astype(int))
    while np.prod(cells) < n_mol:
        cells[np.argmin(cells * a / ext)] += 1
    idx = rng.permutation(int(np.prod(cells)))[:n_mol]
    ijk = np.stack(np.unravel_index(idx, cells), -1)
    spacing = ext / cells
    ox = lo + (ijk + 0.5) * spacing + rng.uniform(-0.5, 0.5, (n_mol, 3)) * np.minimum(spacing * 0.6, 1.0)
    rot = _random_rotations(rng, n_mol)
    h1 = np.array([OH * np.sin(HOH / 2), 0.0, OH * np.cos(HOH / 2)])
    h2 = np.array([-OH * np.sin(HOH / 2), 0.0, OH * np.cos(HOH / 2)])
    return ox, rot @ h1, rot @ h2


def spce_water_frames(n_mol: int, L: float, seed: int, n_frames: int = 1, sigma: float = 0.3) -> np.ndarray:
    """``float32 [n_frames, 3 n_mol, 3]``: frame 0 plus per-frame Gaussian displacements per molecule."""
    ox, h1, h2 = spce_molecules(n_mol, L, seed)
    out = np.empty((n_frames, 3 * n_mol, 3), dtype=np.float32)
    for f in range(n_frames):
        o = ox if f == 0 else ox + np.random.default_rng(seed + f).normal(scale=sigma, size=ox.shape)
        o = o - np.floor(o / L) * L
        out[f, 0::3] = o
        out[f, 1::3] = o + h1
        out[f, 2::3] = o + h2
    return out


def spce_topology(n_mol: int) -> dict:
    """Per-atom topology arrays of ``n_mol`` SPC/E molecules in (O, H, H) order."""
    elements = np.tile(np.array(["O", "H", "H"], dtype=object), n_mol)
    names = np.tile(np.array(["OW", "HW1", "HW2"], dtype=object), n_mol)
    return {
        "names": names, "types": names, "elements": elements,
        "masses": np.tile([M_O, M_H, M_H], n_mol), "charges": np.tile([Q_O, Q_H, Q_H], n_mol),
        "resindices": np.repeat(np.arange(n_mol), 3), "molnums": np.repeat(np.arange(n_mol), 3),
    }


def spce_universe(n_mol: int, seed: int, n_frames: int = 1, L: float | None = None, sigma: float = 0.3):
    """In-memory universe of bulk SPC/E water (configs 2, 3, 5)."""
    from .mdshim import MemoryTrajectory, Universe

    L = cubic_box_for(n_mol) if L is None else L
    frames = spce_water_frames(n_mol, L, seed, n_frames, sigma)
    traj = MemoryTrajectory(frames, np.array([L, L, L, 90, 90, 90], dtype=np.float32))
    return Universe(3 * n_mol, traj, **spce_topology(n_mol))


def confined_slab_frames(n_water_mol: int, n_carbon: int, lx: float, ly: float, lz: float, gap: float, seed: int,
                         n_frames: int = 1, sigma: float = 0.2):
    """Config 4 flavour: two carbon sheets at ``z = lz/2 -/+ gap/2`` and SPC/E water between them.

    Returns ``(frames float32 [F, n_carbon + 3 n_water_mol, 3], topology dict)``; carbons first.
    """
    rng = np.random.default_rng(seed)
    per_sheet = n_carbon // 2
    zc = lz / 2
    sheets = []
    for s, z in enumerate((zc - gap / 2, zc + gap / 2)):
        m = per_sheet if s == 0 else n_carbon - per_sheet
        nx = int(np.ceil(np.sqrt(m * lx / ly)))
        ny = int(np.ceil(m / nx))
        gx, gy = np.meshgrid((np.arange(nx) + 0.5) * lx / nx, (np.arange(ny) + 0.5) * ly / ny, indexing="ij")
        xy = np.stack([gx.ravel(), gy.ravel()], -1)[:m]
        sheets.append(np.concatenate([xy, np.full((m, 1), z)], 1))
    carbon = np.concatenate(sheets)
    ox, h1, h2 = spce_molecules(n_water_mol, [lx, ly, lz], seed + 1, zlo=zc - gap / 2 + 2.5, zhi=zc + gap / 2 - 2.5)
    n = n_carbon + 3 * n_water_mol
    frames = np.empty((n_frames, n, 3), dtype=np.float32)
    box = np.array([lx, ly, lz])
    for f in range(n_frames):
        o = ox if f == 0 else ox + np.random.default_rng(seed + 100 + f).normal(scale=sigma, size=ox.shape)
        o[:, :2] -= np.floor(o[:, :2] / box[:2]) * box[:2]
        frames[f, :n_carbon] = carbon
        frames[f, n_carbon + 0::3] = o
        frames[f, n_carbon + 1::3] = o + h1
        frames[f, n_carbon + 2::3] = o + h2
    w = spce_topology(n_water_mol)
    topo = {
        "names": np.concatenate([np.full(n_carbon, "C", dtype=object), w["names"]]),
        "elements": np.concatenate([np.full(n_carbon, "C", dtype=object), w["elements"]]),
        "masses": np.concatenate([np.full(n_carbon, 12.011), w["masses"]]),
        "charges": np.concatenate([np.zeros(n_carbon), w["charges"]]),
        "resindices": np.concatenate([np.arange(n_carbon), n_carbon + w["resindices"]]),
    }
    topo["types"] = topo["names"]
    topo["molnums"] = topo["resindices"]
    del rng
    return frames, topo
