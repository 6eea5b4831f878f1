"""cProfile of Saxs.run() on config 5 (one 1M-atom frame, qmax = 3/A): host-side cost around the device call."""
import cProfile, pstats, sys, time, warnings
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb
u = mb.synthetic.spce_universe(333_333, 20261022, n_frames=1)
mb.Saxs(u.atoms, qmax=3.0).run()
t0 = time.perf_counter(); a = mb.Saxs(u.atoms, qmax=3.0).run(); print("wall", time.perf_counter() - t0, a.timings)
pr = cProfile.Profile(); pr.enable(); mb.Saxs(u.atoms, qmax=3.0).run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
