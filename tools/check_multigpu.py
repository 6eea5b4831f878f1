"""Run under torchrun on N GPUs: frame-sharded and q-block-sharded Saxs / DiporderSF / DensityPlanar vs one-GPU results."""
import sys
import warnings
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
warnings.simplefilter("ignore")
import maicos_b200 as mb  # noqa: E402
from maicos_b200 import parallel  # noqa: E402


def run_all():
    u = mb.synthetic.spce_universe(2000, 7, n_frames=13)
    out = {}
    s = mb.Saxs(u.atoms, qmax=2.5).run()
    out["saxs_I"], out["saxs_dI"], out["saxs_cnt"] = s.results.scattering_intensities, s.results.dscattering_intensities, s.sums.bincount
    d = mb.DiporderStructureFactor(u.atoms, qmax=2.0, dq=0.05, grouping="residues").run(step=2)
    out["dip_S"] = d.results.structure_factors
    p = mb.DensityPlanar(u.atoms, bin_width=0.5, unwrap=False).run()
    out["dens"], out["ddens"], out["dens_cnt"] = p.results.profile, p.results.dprofile, p.sums.bincount
    return out


serial = run_all()                      # every rank alone on its own GPU
import os
for mode in ("frames", "qblocks"):
    parallel.init(mode=mode)
    sharded = run_all()
    if parallel.rank() == 0:
        for k in serial:
            a, b = np.asarray(serial[k], dtype=float), np.asarray(sharded[k], dtype=float)
            err = np.nanmax(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)) if a.size else 0.0
            exact = "cnt" in k
            ok = np.array_equal(a, b) if exact else err < (1e-9 if not k.startswith(("saxs_d", "ddens")) else 1e-6)
            print(f"[{mode} x{parallel.world()}] {k:10s} max rel diff {err:.2e} {'OK' if ok else 'MISMATCH'}")
            assert ok
    parallel.barrier()
parallel.shutdown()
if int(os.environ.get("RANK", "0")) == 0:
    print("MULTI-GPU OK")
