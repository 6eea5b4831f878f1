"""Single-process multi-GPU (parallel.use_devices): Saxs on the bench workload over every visible GPU, no launcher."""
import json, sys, time, warnings
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
warnings.simplefilter("ignore")
import maicos_b200 as mb
from maicos_b200 import _cabi, parallel
from maicos_b200.mdshim import MemoryTrajectory, Universe
from maicos_b200.synthetic import spce_topology, spce_water_frames

N_MOL, BOX, B = 33_333, 99.97, 40
n_dev = _cabi.device_count()
frames = spce_water_frames(N_MOL, BOX, 7, n_frames=8 * B)
topo = spce_topology(N_MOL)
out = {"devices": n_dev, "frames_per_run": None, "runs": {}}
mb.Saxs._batch_size = B
for devs in ([0], list(range(min(2, n_dev))), list(range(min(4, n_dev))), list(range(n_dev))):
    if len(devs) in out["runs"]:
        continue
    F = 10 * B * len(devs)
    traj = MemoryTrajectory(np.tile(frames, (F // len(frames) + 1, 1, 1))[:F], np.float32([BOX, BOX, BOX, 90, 90, 90]), pinned=True)
    u = Universe(3 * N_MOL, traj, **topo)
    parallel.use_devices(devs)
    try:
        mb.Saxs(u.atoms, qmax=2.0, dq=0.1).run(stop=2 * B * len(devs))  # warm-up on every device
        t0 = time.perf_counter()
        a = mb.Saxs(u.atoms, qmax=2.0, dq=0.1).run()
        for d in devs:
            _cabi.context(d).synchronize()
        dt = time.perf_counter() - t0
    finally:
        parallel.use_devices([])
    nq = 18054
    out["runs"][len(devs)] = {"frames": F, "seconds": dt, "frames_per_s": F / dt, "atomq_per_s": F * 3.0 * N_MOL * nq / dt,
                              "timings": a.timings, "I0": float(a.results.scattering_intensities[3])}
    del traj, u
print(json.dumps(out))
