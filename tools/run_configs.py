"""Run the BASELINE.json configurations (reduced frame counts) through the public classes and print one JSON summary.

C1 shipped water (golden frames), C3 DiporderStructureFactor on 300k-atom water, C4 DensityPlanar + DiporderPlanar on
a 1M-atom confined slab, C5 Saxs on one 1M-atom frame (qmax = 3/A).  C2 is bench.py.  Wall-clock through `.run()`.
"""
import json
import sys
import time
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
warnings.simplefilter("ignore")
import maicos_b200 as mb  # noqa: E402
from maicos_b200 import _cabi  # noqa: E402
from maicos_b200.lib.math import sfactor_maxn  # noqa: E402
from maicos_b200.mdshim import MemoryTrajectory, Universe  # noqa: E402

import os
out = {"precision": os.environ.get("MAICOS_B200_PRECISION", "i8x5")}
ctx = _cabi.context()


def n_valid_q(L, qmax):
    from oracle import kernel
    q, _ = kernel.port_structure_factor(np.zeros((1, 3)), np.double(np.float32([L, L, L])), 0.0, qmax, 0.0, np.pi, np.ones(1))
    return int(np.count_nonzero(q))


def timed(fn):
    """Wall time of fn(), best of two (a fresh box's first pass through a code path pays one-off costs: pinned
    allocations, workspace growth, worker thread start)."""
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        r = fn()
        ctx.synchronize()
        dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (r, dt)
    return best


# ---- C1 ------------------------------------------------------------------------------------------
from conftest import water_universe  # noqa: E402
g = np.load(ROOT / "tests/golden/water_trr_frames.npz")
el = np.load(ROOT / "tests/golden/water_gro.npz")["elements"]
u = water_universe(np.tile(g["positions"], (25, 1, 1)), np.tile(g["dimensions"], (25, 1)), el)
mb.Saxs(u.atoms).run(stop=8)
s, dt = timed(lambda: mb.Saxs(u.atoms).run())
out["C1_saxs_shipped_water"] = {"frames": s.n_frames, "frames_per_s": s.n_frames / dt, "n_atoms": 1530,
                                "note": "4 golden NPT frames of examples/water.trr tiled to 100 frames, default q range"}

# ---- C3 ------------------------------------------------------------------------------------------
n_mol = 100_000
u = mb.synthetic.spce_universe(n_mol, 20261020, n_frames=3)
L = float(u.dimensions[0])
for qmax, dq in ((2.0, 0.01), (6.0, 0.01)):
    nq = n_valid_q(L, qmax)
    a = mb.DiporderStructureFactor(u.atoms, qmax=qmax, dq=dq, unwrap=False, pack=False)
    a.run()
    a, dt = timed(lambda: mb.DiporderStructureFactor(u.atoms, qmax=qmax, dq=dq, unwrap=False, pack=False).run())
    out[f"C3_dipsf_300k_qmax{qmax:g}"] = {
        "frames": a.n_frames, "frames_per_s": a.n_frames / dt, "n_compounds": n_mol, "n_q_valid": nq,
        "atomq_per_s": 3 * n_mol * nq * a.n_frames / dt, "maxn": [int(x) for x in sfactor_maxn(np.double([L, L, L]), qmax)],
        "timings": a.timings}

# ---- C4 ------------------------------------------------------------------------------------------
n_c, n_w = 100_000, 300_000
fr, topo = mb.synthetic.confined_slab_frames(n_w, n_c, 215.0, 215.0, 215.4, 200.0, 20261021, n_frames=24)
u = Universe(fr.shape[1], MemoryTrajectory(fr, np.float32([215.0, 215.0, 215.4, 90, 90, 90])), **topo)
mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False).run()
a, dt = timed(lambda: mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False).run())
out["C4_densityplanar_1M"] = {"frames": a.n_frames, "frames_per_s": a.n_frames / dt, "n_atoms": int(fr.shape[1]),
                              "n_bins": a.n_bins, "timings": a.timings}
water = u.select_atoms("element O H")
mb.DiporderPlanar(water, bin_width=0.1, unwrap=False, pack=False).run()  # warm-up (pinned staging, device buffers)
a, dt = timed(lambda: mb.DiporderPlanar(water, bin_width=0.1, unwrap=False, pack=False).run())
out["C4_diporderplanar_300k_mol"] = {"frames": a.n_frames, "frames_per_s": a.n_frames / dt, "n_bins": a.n_bins,
                                     "timings": a.timings, "note": "unwrap-free water slab; COM + dipole per molecule formed on the device (mb200_compound_prologue)"}

# ---- C5 ------------------------------------------------------------------------------------------
n_mol = 333_333
u = mb.synthetic.spce_universe(n_mol, 20261022, n_frames=1)
L = float(u.dimensions[0])
nq = n_valid_q(L, 3.0)
mb.Saxs(u.atoms, qmax=3.0).run()
a, dt = timed(lambda: mb.Saxs(u.atoms, qmax=3.0).run())
out["C5_saxs_1M_single_frame_qmax3"] = {"frames": 1, "seconds": dt, "n_atoms": 3 * n_mol, "n_q_valid": nq,
                                         "atomq_per_s": 3 * n_mol * nq / dt,
                                         "maxn": [int(x) for x in sfactor_maxn(np.double([L, L, L]), 3.0)], "timings": a.timings}
print(json.dumps(out))
