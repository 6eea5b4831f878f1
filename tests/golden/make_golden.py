"""Generate the committed golden fixtures of tests/golden/ (run in the build container only).

Inputs come from the reference checkout (``/root/reference``: shipped water box and trajectory) and
outputs from the reference's OWN compiled Cython kernel (``oracle/_ref``, built by
``oracle/Makefile`` from ``/root/reference/src/maicos/lib/_cmath.pyx``) plus the NumPy restatement
of the module code (``oracle/restatement.py``).  ``/root/reference`` does not exist on the GPU box,
so the vectors are committed as small ``.npz`` files.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from maicos_b200.mdshim import read_gro, read_trr  # noqa: E402  (file parsing only)
from maicos_b200.mdshim.readers import guess_element  # noqa: E402
from oracle import kernel, restatement  # noqa: E402

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference")
assert kernel.have_reference(), "build oracle/_ref first: make -C oracle ref"
sf = kernel.reference_structure_factor

# ---- inputs -------------------------------------------------------------------------------------
gro = read_gro(REF / "tests/data/water/water.gro")
elements = np.array([guess_element(n) for n in gro["names"]])
np.savez_compressed(OUT / "water_gro.npz", positions=gro["positions"], dimensions=gro["dimensions"],
                    names=gro["names"].astype(str), elements=elements.astype(str))
trr = read_trr(REF / "examples/water.trr")
sel = np.array([0, 33, 67, 100])
np.savez_compressed(OUT / "water_trr_frames.npz", frames=sel, positions=trr["positions"][sel],
                    dimensions=trr["dimensions"][sel], times=trr["times"][sel])

pos = np.double(gro["positions"])
box = np.double(gro["dimensions"][:3])
ones = np.ones(len(pos))


def compact(q, s):
    nz = np.nonzero(s.ravel())[0]
    return dict(shape=np.array(q.shape), cells=nz.astype(np.int32), q=q.ravel()[nz], s=s.ravel()[nz])


# ---- reference kernel outputs (path 1) -----------------------------------------------------------
cases = {
    # tests/lib/test_math.py:183-209
    "theta_window": dict(qmin=0.0, qmax=0.5, thetamin=np.pi / 4, thetamax=np.pi / 2),
    # tests/lib/test_math.py:34-147 (the 50-entry table documents S for qmax = 1)
    "qmax1": dict(qmin=0.0, qmax=1.0, thetamin=0.0, thetamax=np.pi),
    "qmin_window": dict(qmin=0.8, qmax=2.0, thetamin=0.3, thetamax=2.0),
    "default_saxs": dict(qmin=0.0, qmax=6.0, thetamin=0.0, thetamax=np.pi),
}
store = {}
for name, kw in cases.items():
    q, s = sf(pos, box, kw["qmin"], kw["qmax"], kw["thetamin"], kw["thetamax"], ones)
    for k, v in compact(q, s).items():
        store[f"{name}__{k}"] = v
    store[f"{name}__args"] = np.array([kw["qmin"], kw["qmax"], kw["thetamin"], kw["thetamax"]])
# signed weights on an orthorhombic cell (the DiporderStructureFactor flavour)
rng = np.random.default_rng(20261019)
p2 = rng.uniform(-3.0, 40.0, (777, 3))
w2 = rng.normal(size=777)
b2 = np.array([31.0, 27.5, 36.25])
q, s = sf(p2, b2, 0.0, 1.5, 0.0, np.pi, w2)
for k, v in compact(q, s).items():
    store[f"ortho_signed__{k}"] = v
store["ortho_signed__args"] = np.array([0.0, 1.5, 0.0, np.pi])
store["ortho_signed__positions"] = p2
store["ortho_signed__weights"] = w2
store["ortho_signed__box"] = b2
np.savez_compressed(OUT / "sfactor_reference.npz", **store)

# ---- module-level restatement with the reference kernel -----------------------------------------
# tests/modules/test_saxs.py:57-71: Saxs(water.gro, qmax=20).results.scattering_intensities[0] == 1.6047
res, stats, series = restatement.run_saxs(gro["positions"][None], gro["dimensions"][None, :3], elements, qmax=20,
                                          sfactor=sf)
print("Saxs(water.gro, qmax=20) I[0] =", res["scattering_intensities"][0], "(reference test: 1.6047 rtol 1e-3)")
np.savez_compressed(OUT / "saxs_water_gro_qmax20.npz", **res, series=series)
# default Saxs over four frames of the NPT trajectory (box changes frame to frame)
res, stats, series = restatement.run_saxs(trr["positions"][sel], trr["dimensions"][sel, :3], elements, sfactor=sf)
np.savez_compressed(OUT / "saxs_water_trr_default.npz", **res, series=series,
                    **{f"sum_{k}": v for k, v in stats.sums.items()},
                    **{f"sem_{k}": v for k, v in stats.sems.items()})
print("wrote", sorted(p.name for p in OUT.glob("*.npz")))
