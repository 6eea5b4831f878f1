"""Profiling driver: a few structure_factor calls on the C2 hydrogen group (66 666 atoms, qmax 2)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from maicos_b200 import _cabi
from maicos_b200.lib._sfactor import structure_factor
rng = np.random.default_rng(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 66666
L = float(sys.argv[2]) if len(sys.argv) > 2 else 99.97
qmax = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
pos = rng.uniform(0, L, (n, 3)); w = np.ones(n); dims = np.array([L] * 3)
ctx = _cabi.context(); ctx.set_timing(True)
for _ in range(3):
    structure_factor(pos, dims, 0, qmax, 0, np.pi, w)
    st = ctx.stats()
    print(st, f"{st['atom_q']/max(st['contract_ns'],1)*1e9:.3e} atom*q/s")
