"""Developer probe (GPU box): Saxs on an NPT trajectory -- every frame its own cell, so every frame its own q-plan."""
import os
import sys
import time
import warnings
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb  # noqa: E402
from maicos_b200.mdshim import MemoryTrajectory, Universe  # noqa: E402
from maicos_b200.synthetic import cubic_box_for, spce_topology, spce_water_frames  # noqa: E402

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 33333
F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
L = cubic_box_for(n_mol)
fr = spce_water_frames(n_mol, L, 5, n_frames=F)
topo = spce_topology(n_mol)
scale = 1.0 + 1e-4 * np.sin(np.arange(F))[:, None]           # a barostat's breathing: F distinct cells
dims_npt = np.float32(np.concatenate([np.full((F, 3), L) * scale, np.full((F, 3), 90.0)], axis=1))
dims_nvt = np.float32([L, L, L, 90, 90, 90])
for label, dims in (("NVT", dims_nvt), ("NPT", dims_npt), ("NVT", dims_nvt), ("NPT", dims_npt)):
    u = Universe(fr.shape[1], MemoryTrajectory(fr, dims, pinned=True), **topo)
    mb.Saxs(u.atoms, qmax=2.0).run(stop=8)
    t = time.perf_counter()
    a = mb.Saxs(u.atoms, qmax=2.0).run()
    dt = time.perf_counter() - t
    print(f"{label}: {F / dt:8.1f} frames/s  {dt * 1e3:8.1f} ms  " + " ".join(f"{k}={v * 1e3:.1f}" for k, v in a.timings.items()), flush=True)
if os.environ.get("MAICOS_B200_TRACE_HOST"):
    print("traced")
