"""Config 3 (DiporderStructureFactor, 100k molecules, default qmax = 6 / qmax = 2) once, for a kernel launch list."""
import sys, time, warnings
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb
qmax = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
u = mb.synthetic.spce_universe(100_000, 20261020, n_frames=3)
for rep in range(2):
    t0 = time.perf_counter(); a = mb.DiporderStructureFactor(u.atoms, qmax=qmax, dq=0.01, unwrap=False, pack=False).run()
    print(f"qmax {qmax}: wall {(time.perf_counter() - t0) * 1e3:.1f} ms for 3 frames, timings {a.timings}")
