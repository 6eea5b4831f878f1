"""Developer smoke check on a GPU box: CUDA S(q) + histograms vs the oracle (not a test)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from maicos_b200 import _cabi
from maicos_b200.lib._sfactor import structure_factor
from oracle import kernel as ok

rng = np.random.default_rng(1)
ctx = _cabi.context()
ctx.set_timing(True)

def check(n, L, qmin, qmax, thmin, thmax, weights=None, label=""):
    pos = rng.uniform(-0.3 * L[0], 1.4 * L[0], (n, 3))
    w = np.ones(n) if weights is None else weights
    dims = np.array(L, dtype=np.float64)
    t = time.time(); q1, s1 = ok.structure_factor(pos, dims, qmin, qmax, thmin, thmax, w); tc = time.time() - t
    t = time.time(); q2, s2 = structure_factor(pos, dims, qmin, qmax, thmin, thmax, w); tg = time.time() - t
    st = ctx.stats()
    mask_eq = np.array_equal(s1 != 0, s2 != 0)
    q_eq = np.array_equal(q1, q2)
    nz = s1 != 0
    rel = np.max(np.abs(s1[nz] - s2[nz]) / s1[nz]) if nz.any() else 0.0
    print(f"{label:28s} n={n:7d} shape={q1.shape} nq={nz.sum():7d} mask_eq={mask_eq} q_eq={q_eq} maxrel={rel:.2e} "
          f"cpu={tc*1e3:8.1f}ms gpu={tg*1e3:8.1f}ms kern={st['contract_ns']/1e6:7.3f}ms ctas={st['ctas']} splits={st['k_splits']}")
    return mask_eq and q_eq and rel < 1e-6

ok_all = True
ok_all &= check(1530, [24.8683] * 3, 0, 6, 0, np.pi, label="water-size default")
ok_all &= check(1530, [24.8683] * 3, 0, 0.5, np.pi / 4, np.pi / 2, label="theta window")
ok_all &= check(1000, [20.0, 25.0, 31.0], 0.3, 3.0, 0.2, 1.2, label="ortho box q/theta window")
ok_all &= check(7, [10.0, 10.0, 10.0], 0, 2, 0, np.pi, label="7 atoms")
ok_all &= check(1, [10.0, 10.0, 10.0], 0, 2, 0, np.pi, label="1 atom")
ok_all &= check(5000, [30.0, 30, 30], 0, 2, 0, np.pi, weights=rng.normal(size=5000), label="signed weights")
ok_all &= check(33333, [99.97] * 3, 0, 2, 0, np.pi, label="C2 oxygen group")
ok_all &= check(100, [50.0] * 3, 0, 0.05, 0, np.pi, label="no valid q")
print("S(q) ALL OK" if ok_all else "S(q) FAILURES")

# repeated call timing (plan cached)
pos = rng.uniform(0, 99.97, (66666, 3)); w = np.ones(66666); dims = np.array([99.97] * 3)
for _ in range(3):
    t = time.time(); structure_factor(pos, dims, 0, 2, 0, np.pi, w); tg = time.time() - t
    st = ctx.stats()
    print(f"C2 H group: wall={tg*1e3:.2f} ms kern={st['contract_ns']/1e6:.3f} ms atom_q={st['atom_q']:.3e} "
          f"-> {st['atom_q']/max(st['contract_ns'],1)*1e9:.3e} atom*q/s (kernel), ctas={st['ctas']} dmma={st['dmma_steps']}")

# histograms
lib = _cabi.load()
def hist(geom, pos, dim, w, edges, center=None, zrange=None):
    F, n = pos.shape[0], pos.shape[1]
    nb = edges.shape[1] - 1
    hw = np.zeros((F, nb)); hc = np.zeros((F, nb), dtype=np.int64)
    _cabi.check(lib.mb200_profile_hist(ctx.handle, geom, _cabi.ptr(pos), int(pos.dtype == np.float32), 0, F, n, dim,
        _cabi.ptr(w), int(w is not None and w.ndim == 2), 0, _cabi.ptr(edges), nb,
        _cabi.ptr(center), _cabi.ptr(zrange), _cabi.ptr(hw) if w is not None else None, _cabi.ptr(hc)))
    return hw, hc

F, n, nb = 3, 200000, 2153
L = np.array([215.37, 215.37, 215.3])
pos = rng.uniform(-5, 220, (F, n, 3)).astype(np.float32)
pos[0, :50, 2] = 0.0; pos[0, 50:100, 2] = np.float32(215.3)
w = rng.uniform(1, 16, n)
edges = np.stack([np.linspace(np.float64(0), np.float64(L[2]) * (1 + 0.001 * f), nb + 1) for f in range(F)])
hw, hc = hist(0, pos, 2, w, edges)
okh = True
for f in range(F):
    rw, _ = np.histogram(pos[f, :, 2], bins=nb, range=(edges[f, 0], edges[f, -1]), weights=w)
    rc, _ = np.histogram(pos[f, :, 2], bins=nb, range=(edges[f, 0], edges[f, -1]))
    okh &= np.array_equal(rc, hc[f]); 
    print("planar frame", f, "counts equal", np.array_equal(rc, hc[f]), "maxrel w", np.max(np.abs(rw - hw[f]) / np.maximum(np.abs(rw), 1e-300)))
center = np.tile(L / 2, (F, 1)); zr = np.tile(np.array([10.0, 200.0]), (F, 1))
edges_r = np.stack([np.linspace(np.float64(0.0), np.float64(100.0), 501) for f in range(F)])
hw, hc = hist(1, pos, 2, w, edges_r, center, zr)
for f in range(F):
    d = pos[f].astype(np.float64) - center[f]
    r = np.linalg.norm(d[:, [0, 1]], axis=1)
    rw = np.histogram2d(r, pos[f, :, 2], bins=(500, 1), range=((0.0, 100.0), (10.0, 200.0)), weights=w)[0][:, 0]
    rc = np.histogram2d(r, pos[f, :, 2], bins=(500, 1), range=((0.0, 100.0), (10.0, 200.0)))[0][:, 0]
    okh &= np.array_equal(rc.astype(np.int64), hc[f])
    print("cyl frame", f, "counts equal", np.array_equal(rc.astype(np.int64), hc[f]), "maxrel w", np.max(np.abs(rw - hw[f]) / np.maximum(np.abs(rw), 1e-300)))
hw, hc = hist(2, pos, 0, w, edges_r, center, None)
for f in range(F):
    d = pos[f].astype(np.float64) - center[f]
    r = np.linalg.norm(d, axis=1)
    rw, _ = np.histogram(r, bins=500, range=(0.0, 100.0), weights=w); rc, _ = np.histogram(r, bins=500, range=(0.0, 100.0))
    okh &= np.array_equal(rc, hc[f])
    print("sph frame", f, "counts equal", np.array_equal(rc, hc[f]), "maxrel w", np.max(np.abs(rw - hw[f]) / np.maximum(np.abs(rw), 1e-300)))
print("HIST ALL OK" if okh else "HIST FAILURES")
