"""Developer probe (GPU box): DensityPlanar on a 1M-atom slab, page-locked vs pageable trajectory, run after run."""
import sys
import time
import warnings
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb  # noqa: E402
from maicos_b200 import _cabi  # noqa: E402
from maicos_b200.mdshim import MemoryTrajectory, Universe  # noqa: E402
from maicos_b200.synthetic import confined_slab_frames  # noqa: E402

F = 96
box = [215.0, 215.0, 215.4]
fr, topo = confined_slab_frames(300_000, 100_000, *box, 200.0, 20261021, n_frames=F)
u_pg = Universe(fr.shape[1], MemoryTrajectory(fr, np.float32([*box, 90, 90, 90])), **topo)
u = Universe(fr.shape[1], MemoryTrajectory(fr, np.float32([*box, 90, 90, 90]), pinned=True), **topo)
lib, ctx = _cabi.load(), _cabi.context()
WATER = {id(u): u.select_atoms("element O H"), id(u_pg): u_pg.select_atoms("element O H")}


def pstat():
    o = np.zeros(3, dtype=np.int64)
    if hasattr(lib, "mb200_ctx_prefetch_stats"):
        _cabi.check(lib.mb200_ctx_prefetch_stats(ctx.handle, _cabi.ptr(o)))
    return o


def run(uu, label, cls=None, **kw):
    p0 = pstat()
    t = time.perf_counter()
    if cls is None:
        a = mb.DensityPlanar(uu.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False).run(**kw)
    else:
        a = cls(WATER[id(uu)], bin_width=0.1, unwrap=False, pack=False).run(**kw)
    dt = time.perf_counter() - t
    print(f"{label:10s} {a.n_frames / dt:8.0f} fps  {dt * 1e3:7.1f} ms  prefetch {(pstat() - p0).tolist()}  "
          + " ".join(f"{k}={v * 1e3:.1f}" for k, v in a.timings.items()), flush=True)


for i in range(3):
    run(u, "pinned")
for i in range(3):
    run(u_pg, "pageable")
for i in range(4):
    run(u, "pinned")
run(u, "pinned/3", stop=F // 3)
for i in range(3):
    run(u, "pinned")
for i in range(4):
    run(u, "diporder", cls=mb.DiporderPlanar)
for i in range(3):
    run(u_pg, "diporder pg", cls=mb.DiporderPlanar)
import os  # noqa: E402
if os.environ.get("MAICOS_B200_TRACE_HOST"):
    print("traced run above", flush=True)
if os.environ.get("MB200_PROFILE"):
    import cProfile
    import pstats
    water = u.select_atoms("element O H")
    pr = cProfile.Profile()
    pr.enable()
    mb.DiporderPlanar(water, bin_width=0.1, unwrap=False, pack=False).run()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
