"""cProfile of Saxs.run() on the bench workload (host-side staging cost)."""
import cProfile, pstats, sys, time, warnings
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb
from maicos_b200.mdshim import MemoryTrajectory, Universe
from maicos_b200.synthetic import spce_topology, spce_water_frames
N_MOL, BOX = 33333, 99.97
pool = spce_water_frames(N_MOL, BOX, 1, n_frames=40)
topo = spce_topology(N_MOL)
def universe(nf):
    traj = MemoryTrajectory(dimensions=np.float32([BOX, BOX, BOX, 90, 90, 90]), n_frames=nf,
                            frame_source=lambda f: pool[f % 40].copy(), n_atoms=3 * N_MOL)
    return Universe(3 * N_MOL, traj, **topo)
mb.Saxs._batch_size = 40
mb.Saxs(universe(80).atoms, qmax=2.0).run()
for nf in (400, 1000):
    u = universe(nf)
    t0 = time.perf_counter()
    t1 = time.perf_counter()
    sx = mb.Saxs(u.atoms, qmax=2.0)
    sx.run()
    dt = time.perf_counter() - t0
    print(f"{nf} frames: wall {dt:.3f} s = {dt / nf * 1e3:.3f} ms/frame; timings {sx.timings}")
