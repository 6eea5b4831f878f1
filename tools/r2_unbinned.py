"""Saxs(bin_spectrum=False) frames per second through the class (batched ``mb200_saxs_frames_cells``)."""
import json
import sys
import time
import warnings
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb  # noqa: E402

n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 33333
F = int(sys.argv[2]) if len(sys.argv) > 2 else 64
u = mb.synthetic.spce_universe(n_mol, 3, n_frames=F)
out = {}
for qmax in (2.0, 4.0):
    mb.Saxs(u.atoms, bin_spectrum=False, qmax=qmax).run(stop=8)
    t = time.perf_counter()
    a = mb.Saxs(u.atoms, bin_spectrum=False, qmax=qmax).run()
    dt = time.perf_counter() - t
    out[f"qmax{qmax}"] = {"atoms": u.atoms.n_atoms, "frames": F, "cells": int(np.prod(a.max_n)), "frames_per_s": F / dt}
print(json.dumps(out))
