"""Shared fixtures.  ``-m gpu`` tests need a B200; everything else runs on CPU."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """The C port of the oracle is compiled on demand (gcc); oracle/_ref travels prebuilt."""
    from oracle import kernel

    if not (ROOT / "oracle" / "liboracle_port.so").exists():
        kernel.build()


def load_golden(name: str):
    return np.load(GOLDEN / name, allow_pickle=False)


@pytest.fixture(scope="session")
def water_gro():
    """Shipped 1530-atom SPC/E box (reference tests/data/water/water.gro) as arrays."""
    g = load_golden("water_gro.npz")
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def water_frames():
    """Frames 0, 33, 67, 100 of the reference's examples/water.trr (NPT: the box changes)."""
    g = load_golden("water_trr_frames.npz")
    return {k: g[k] for k in g.files}


def water_universe(positions, dimensions, elements):
    """SPC/E universe (O, H, H per residue) on the in-memory shim."""
    from maicos_b200.mdshim import MemoryTrajectory, Universe

    positions = np.asarray(positions, dtype=np.float32)
    if positions.ndim == 2:
        positions = positions[None]
    n = positions.shape[1]
    elements = np.asarray(elements).astype(object)
    masses = np.where(elements == "O", 15.9994, 1.008)
    charges = np.where(elements == "O", -0.8476, 0.4238)
    res = np.arange(n) // 3
    traj = MemoryTrajectory(positions, np.asarray(dimensions, dtype=np.float32))
    names = np.where(elements == "O", "OW", "HW").astype(object)
    return Universe(n, traj, names=names, types=names, elements=elements, masses=masses, charges=charges,
                    resindices=res, molnums=res)


def spce_box(n_mol, L, seed, n_frames=1, sigma=0.3):
    """Seeded synthetic SPC/E water (lattice + jitter, random orientations) -> float32 [F, 3 n_mol, 3]."""
    from maicos_b200.synthetic import spce_water_frames

    return spce_water_frames(n_mol, L, seed, n_frames=n_frames, sigma=sigma)


def compact_to_cube(g, name):
    shape = tuple(int(x) for x in g[f"{name}__shape"])
    q = np.zeros(int(np.prod(shape)))
    s = np.zeros_like(q)
    q[g[f"{name}__cells"]] = g[f"{name}__q"]
    s[g[f"{name}__cells"]] = g[f"{name}__s"]
    return q.reshape(shape), s.reshape(shape)


def dipole_universe(positions, orientations, box=10.0):
    """The reference's ``dipoles()`` fixture (tests/modules/test_dielectricplanar.py:33-60) on the shim: one molecule
    per dipole, charges -1 / +1 one Angstrom apart (tests/data/dipole.gro / dipole.itp), moment 1 e A along
    ``orientation``, cubic box of 10 A."""
    from maicos_b200.mdshim import MemoryTrajectory, Universe

    pos = []
    for p, o in zip(positions, orientations):
        o = np.asarray(o, dtype=np.float64)
        o = o / np.linalg.norm(o)
        pos += [np.asarray(p, dtype=np.float64) - 0.5 * o, np.asarray(p, dtype=np.float64) + 0.5 * o]
    n = len(pos)
    res = np.arange(n) // 2
    traj = MemoryTrajectory(np.asarray(pos, dtype=np.float32)[None], np.float32([box, box, box, 90, 90, 90]))
    names = np.array(["A", "B"] * (n // 2), dtype=object)
    return Universe(n, traj, names=names, types=np.array(["X"] * n, dtype=object), masses=np.ones(n),
                    charges=np.array([-1.0, 1.0] * (n // 2)), resindices=res, molnums=res,
                    bonds=np.stack([np.arange(0, n, 2), np.arange(1, n, 2)], 1))
