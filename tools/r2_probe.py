"""Round-2 developer probe (GPU box): host-side phases of single huge-frame calls, q-block timings on one GPU."""
import os, sys, time, warnings
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
warnings.simplefilter("ignore")
os.environ["MAICOS_B200_TRACE_HOST"] = "1"
import maicos_b200 as mb
from maicos_b200 import _cabi
from maicos_b200.mdshim import MemoryTrajectory, Universe
from maicos_b200.synthetic import spce_topology, spce_water_frames, cubic_box_for

n_mol = 333_333
L = cubic_box_for(n_mol)
fr = spce_water_frames(n_mol, L, 1, n_frames=1)
u = Universe(3 * n_mol, MemoryTrajectory(fr, np.float32([L, L, L, 90, 90, 90]), pinned=True), **spce_topology(n_mol))
ctx = _cabi.context()
for i in range(3):
    t0 = time.perf_counter()
    a = mb.Saxs(u.atoms, qmax=3.0).run()
    ctx.synchronize()
    print(f"run {i}: wall {1e3 * (time.perf_counter() - t0):.2f} ms  timings {a.timings}", file=sys.stderr)
# emulate 8 q-blocks on this one GPU: block b of 8, device + wall time per block
from maicos_b200 import parallel
for nb in (2, 4, 8):
    parallel._state.update(world=nb, rank=0, mode="qblocks")
    import maicos_b200.parallel as P
    keep = P.allreduce_sum
    P.allreduce_sum = lambda buf: buf
    try:
        for i in range(2):
            t0 = time.perf_counter()
            a = mb.Saxs(u.atoms, qmax=3.0).run()
            ctx.synchronize()
            print(f"q-block 0/{nb} run {i}: wall {1e3 * (time.perf_counter() - t0):.2f} ms timings {a.timings}", file=sys.stderr)
    finally:
        P.allreduce_sum = keep
        parallel._state.update(world=1, rank=0, mode="frames")
