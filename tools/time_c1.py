"""C1 (shipped water, 1530 atoms, NPT boxes) through Saxs.run(): frames/s over repeated runs + host time split."""
import sys, time, warnings
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
warnings.simplefilter("ignore")
import maicos_b200 as mb
from conftest import water_universe
g = np.load(ROOT / "tests/golden/water_trr_frames.npz")
el = np.load(ROOT / "tests/golden/water_gro.npz")["elements"]
u = water_universe(np.tile(g["positions"], (25, 1, 1)), np.tile(g["dimensions"], (25, 1)), el)
mb.Saxs(u.atoms).run(stop=8)
for rep in range(4):
    t0 = time.perf_counter(); s = mb.Saxs(u.atoms).run(); dt = time.perf_counter() - t0
    print(rep, f"{s.n_frames / dt:.0f} frames/s", {k: round(v * 1e3, 2) for k, v in s.timings.items()}, "ms")
if len(sys.argv) > 1:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable(); mb.Saxs(u.atoms).run(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
