"""One rank's share of config 5 on ONE GPU (q-block 0 of N): where the time of a q-sharded call goes."""
import sys, time, warnings
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb
from maicos_b200 import parallel, _cabi
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
u = mb.synthetic.spce_universe(333_333, 20261022, n_frames=1)
parallel.q_block = lambda: (0, nb)
ctx = _cabi.context(None)
ctx.set_timing(True)
for rep in range(3):
    t0 = time.perf_counter(); a = mb.Saxs(u.atoms, qmax=3.0).run(); dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"block 0/{nb}: wall {dt*1e3:.1f} ms, device call {a.timings['device_call_s']*1e3:.1f} ms, contraction {st['contract_ns']*1e-6:.2f} ms, items {st['ctas']}, launches {st['launches']}")
