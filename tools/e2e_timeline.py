"""Where the fixed cost of Saxs.run() goes (bench workload, 400 frames): timestamps of constructor, _prepare, first / last
device call and the end of run(), relative to the start."""
import sys, time, warnings
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
def universe(nf, off):
    traj = MemoryTrajectory(dimensions=np.float32([BOX, BOX, BOX, 90, 90, 90]), n_frames=nf,
                            frame_source=lambda f: pool[(f + off) % 40].copy(), n_atoms=3 * N_MOL)
    return Universe(3 * N_MOL, traj, **topo)
mb.Saxs._batch_size = 40
mb.Saxs(universe(120, 0).atoms, qmax=2.0).run()
marks = []
orig_batch = mb.Saxs._device_batch
def traced(self, staged):
    t0 = time.perf_counter(); out = orig_batch(self, staged); marks.append((t0, time.perf_counter(), len(staged))); return out
mb.Saxs._device_batch = traced
orig_prepare = mb.Saxs._call_prepare
def prep(self):
    t0 = time.perf_counter(); orig_prepare(self); marks.append(("prepare", t0, time.perf_counter()))
mb.Saxs._call_prepare = prep
for rep in range(2):
    marks.clear()
    u = universe(400, 17)
    T0 = time.perf_counter()
    sx = mb.Saxs(u.atoms, qmax=2.0); t_ctor = time.perf_counter()
    sx.run(); T1 = time.perf_counter()
    prep_m = [m for m in marks if m[0] == "prepare"][0]
    calls = [m for m in marks if m[0] != "prepare"]
    print(f"rep {rep}: total {1e3 * (T1 - T0):.1f} ms | ctor {1e3 * (t_ctor - T0):.1f} | prepare {1e3 * (prep_m[2] - prep_m[1]):.1f} (ends at {1e3 * (prep_m[2] - T0):.1f}) | "
          f"first device call starts at {1e3 * (calls[0][0] - T0):.1f} | calls " + " ".join(f"{c[2]}f:{1e3 * (c[1] - c[0]):.1f}" for c in calls[:4]) +
          f" ... | last call ends at {1e3 * (calls[-1][1] - T0):.1f} | idle between calls {1e3 * sum(calls[i + 1][0] - calls[i][1] for i in range(len(calls) - 1)):.2f} ms")
