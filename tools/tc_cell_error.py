"""Per-cell error of the split-precision (i8x5) contraction against the FP64 DMMA contraction on the raw S cube."""
import sys
sys.path.insert(0, ".")
import numpy as np
import maicos_b200 as mb
from maicos_b200.lib.math import structure_factor

rng = np.random.default_rng(7)
for n, L, qmax, wkind in ((1530, 24.87, 6.0, "ones"), (33333, 99.97, 2.0, "ones"), (99999, 99.97, 2.0, "ones"),
                          (100000, 144.18, 2.0, "unit dipole"), (300000, 144.18, 3.0, "ones"), (999999, 215.37, 2.0, "ones")):
    pos = rng.uniform(0, L, (n, 3))
    w = np.ones(n) if wkind == "ones" else rng.uniform(-1, 1, n)
    box = np.array([L, L, L])
    mb.set_precision("fp64"); q0, s0 = structure_factor(pos, box, 0, qmax, 0, np.pi, w)
    mb.set_precision("i8x5"); q1, s1 = structure_factor(pos, box, 0, qmax, 0, np.pi, w)
    nz = s0 != 0
    rel = np.abs(s1[nz] - s0[nz]) / s0[nz]
    dA = np.abs(np.sqrt(s1[nz]) - np.sqrt(s0[nz]))
    print(f"n={n:7d} L={L} qmax={qmax} w={wkind:11s} cells={nz.sum():7d} mask_equal={np.array_equal(s1 != 0, nz)} q_equal={np.array_equal(q0, q1)} "
          f"max rel dS={rel.max():.2e} median={np.median(rel):.2e} cells>1e-6: {(rel > 1e-6).sum()} >1e-8: {(rel > 1e-8).sum()} "
          f"max d|A|={dA.max():.2e} min S/mean S={s0[nz].min() / s0[nz].mean():.2e}", flush=True)
mb.set_precision("fp64")
