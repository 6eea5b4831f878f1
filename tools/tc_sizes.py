"""Developer tool: split-precision contraction timing over a few (atoms, box, qmax) shapes (fused Saxs call)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import maicos_b200 as mb
from maicos_b200 import _cabi
mb.set_precision(sys.argv[1] if len(sys.argv) > 1 else "i8x5")
ctx = _cabi.context(); ctx.set_timing(True)
for n_mol, qmax, nf in ((510, 6.0, 50), (3000, 3.0, 20), (33333, 2.0, 8), (100000, 2.0, 3), (100000, 6.0, 1), (333333, 3.0, 1)):
    u = mb.synthetic.spce_universe(n_mol, 1, n_frames=nf)
    mb.Saxs(u.atoms, qmax=qmax).run()
    t0 = time.perf_counter(); a = mb.Saxs(u.atoms, qmax=qmax).run(); ctx.synchronize(); dt = time.perf_counter() - t0
    st = ctx.stats()
    print(f"n_mol={n_mol:7d} qmax={qmax} frames={nf:3d} wall={dt*1e3:8.2f} ms kernel={st['contract_ns']/1e6:8.3f} ms items={st['ctas']} "
          f"atomq={st['atom_q']:.3e} -> {st['atom_q']/max(st['contract_ns'],1)*1e9:.3e} atom*q/s in the contraction", flush=True)
