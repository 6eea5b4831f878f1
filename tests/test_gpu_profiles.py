"""Parity of the CUDA slab / shell histograms (path 2) against numpy: counts bit-exact."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import maicos_b200 as mb
from maicos_b200.core import ProfileCylinderBase, ProfilePlanarBase, ProfileSphereBase
from maicos_b200.lib._hist import profile_hist
from maicos_b200.lib.weights import density_weights
from maicos_b200.mdshim import MemoryTrajectory, Universe
from oracle import restatement

pytestmark = pytest.mark.gpu
WTOL = 1e-12  # weighted sums: float64 atomics in arbitrary order


def frames(F, n, lo, hi, dtype, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (F, n, 3)).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_bins", [1, 7, 2153, 5000])
def test_planar_vs_numpy(dtype, n_bins):
    F, n = 3, 50_000
    pos = frames(F, n, -5, 220, dtype)
    pos[0, :50, 2] = 0.0
    pos[0, 50:100, 2] = dtype(215.3)  # exactly on the inclusive right edge of frame 0
    w = np.random.default_rng(1).uniform(-1, 16, (F, n))
    edges = np.stack([np.linspace(np.float64(0), np.float64(215.3) * (1 + 1e-3 * f), n_bins + 1) for f in range(F)])
    pos[1, :n_bins + 1, 2] = edges[1].astype(dtype)  # values sitting on bin edges
    for weights in (None, w[:1], w):
        hw, hc = profile_hist(0, pos, 2, weights, edges)
        for f in range(F):
            wf = None if weights is None else weights[min(f, len(weights) - 1)]
            assert_array_equal(hc[f], restatement.hist_planar(pos[f], None, 2, n_bins, edges[f, 0], edges[f, -1]))
            if weights is not None:
                ref = restatement.hist_planar(pos[f], wf, 2, n_bins, edges[f, 0], edges[f, -1])
                assert_allclose(hw[f], ref, rtol=WTOL, atol=WTOL * np.abs(wf).sum())


@pytest.mark.parametrize("nb", [333, 1500])
@pytest.mark.parametrize("dim", [0, 1, 2])
def test_cylinder_and_sphere_vs_numpy(dim, nb):
    F, n = 2, 40_000
    L = np.array([60.0, 70.0, 80.0])
    pos = frames(F, n, -3, 83, np.float32, seed=dim)
    w = np.random.default_rng(2).normal(size=(F, n))
    center = np.tile(L / 2, (F, 1))
    zr = np.array([[5.0, 55.0], [0.0, 60.0]])
    er = np.stack([np.linspace(np.float64(0.5), np.float64(29.0 + f), nb + 1) for f in range(F)])
    hw, hc = profile_hist(1, pos, dim, w, er, center, zr)
    for f in range(F):
        p = pos[f]
        rc = restatement.hist_cylinder(p, None, dim, nb, er[f, 0], er[f, -1], zr[f, 0], zr[f, 1], center[f])
        rw = restatement.hist_cylinder(p, w[f], dim, nb, er[f, 0], er[f, -1], zr[f, 0], zr[f, 1], center[f])
        assert_array_equal(hc[f], rc.astype(np.int64))
        assert_allclose(hw[f], rw, rtol=WTOL, atol=WTOL * np.abs(w[f]).sum())
    hw, hc = profile_hist(2, pos, 0, w, er, center, None)
    for f in range(F):
        assert_array_equal(hc[f], restatement.hist_sphere(pos[f], None, nb, er[f, 0], er[f, -1], center[f]))
        assert_allclose(hw[f], restatement.hist_sphere(pos[f], w[f], nb, er[f, 0], er[f, -1], center[f]),
                        rtol=WTOL, atol=WTOL * np.abs(w[f]).sum())


def test_heavy_collisions_unaligned_frames_and_nonfinite_weights():
    """Half of the atoms in one bin (carry chains of the fixed-point limbs), unaligned frames, inf / nan weights."""
    F, n, nb = 3, 33_333, 1200  # n * 12 B is not a multiple of 16: frames start unaligned
    rng = np.random.default_rng(8)
    pos = rng.uniform(0, 50, (F, n, 3)).astype(np.float32)
    pos[:, : n // 2, 2] = np.float32(17.31)  # half of the atoms in one bin
    w = rng.normal(size=(F, n)) * 1e3
    e = np.tile(np.linspace(np.float64(0), np.float64(50), nb + 1), (F, 1))
    hw, hc = profile_hist(0, pos, 2, w, e)
    for f in range(F):
        assert_array_equal(hc[f], restatement.hist_planar(pos[f], None, 2, nb, 0.0, 50.0))
        assert_allclose(hw[f], restatement.hist_planar(pos[f], w[f], 2, nb, 0.0, 50.0), rtol=WTOL,
                        atol=WTOL * np.abs(w[f]).sum())
    hw2, _ = profile_hist(0, pos, 2, w, e)
    assert_array_equal(hw2, hw)  # fixed-point sums: bitwise reproducible
    w[1, 5] = np.inf
    w[2, 7] = np.nan
    hw, hc = profile_hist(0, pos, 2, w, e)
    assert_allclose(hw[0], restatement.hist_planar(pos[0], w[0], 2, nb, 0.0, 50.0), rtol=WTOL, atol=WTOL * np.abs(w[0]).sum())
    assert np.isinf(hw[1]).sum() == 1 and np.isnan(hw[2]).sum() == 1
    assert_array_equal(hc[1], restatement.hist_planar(pos[1], None, 2, nb, 0.0, 50.0))


@pytest.mark.parametrize("n_bins", [1, 31, 512, 513, 1024, 1537, 2154, 3000, 4096])
def test_sorted_kernel_bitwise_equals_atomics_kernel(n_bins):
    """Counting-sort kernel (default) vs the three-atomics kernel: identical exact integer sums, every bins-per-thread
    instantiation, every layout (float32 aligned / unaligned, float64, scalar, digitize, cylinder, sphere, hand-made
    edges that are not numpy's linspace), tiles that are not full, weights of mixed sign and magnitude."""
    from maicos_b200 import _cabi
    ctx = _cabi.context(None)
    rng = np.random.default_rng(n_bins)
    F, n = 3, 29_999  # n * 12 B is not a multiple of 16
    pos32 = rng.uniform(-2, 52, (F, n, 3)).astype(np.float32)
    pos32[0, : n // 3, 2] = np.float32(11.7)  # heavy collisions: segment sums longer than one tile share
    w = rng.normal(size=(F, n)) * np.exp(rng.uniform(-20, 20, (F, n)))
    lin = np.stack([np.linspace(np.float64(0), np.float64(50 + f), n_bins + 1) for f in range(F)])
    hand = lin.copy()
    if n_bins > 2:
        hand[:, 1:-1] += rng.uniform(-0.3, 0.3, (F, n_bins - 1)) * (50.0 / n_bins)  # still increasing, not a linspace
    center = np.tile([25.0, 26.0, 27.0], (F, 1))
    zr = np.tile([3.0, 45.0], (F, 1))
    radial = np.stack([np.linspace(np.float64(0.25), np.float64(24 + f), n_bins + 1) for f in range(F)])
    cases = [
        (0, pos32, 2, w, lin, None, None, False),
        (0, pos32, 0, w[:1], hand, None, None, False),
        (0, pos32.astype(np.float64), 1, w, lin, None, None, False),
        (3, np.ascontiguousarray(pos32[:, :, 1]), 0, w, lin, None, None, True),
        (3, np.ascontiguousarray(pos32[:, :, 1]).astype(np.float64), 0, w, hand, None, None, False),
        (1, pos32, 1, w, radial, center, zr, False),
        (2, pos32, 0, w, radial, center, None, False),
        (0, pos32[:, :4097], 2, w[:, :4097], lin, None, None, True),
        (0, pos32[:, :5], 2, w[:, :5], lin, None, None, False),
    ]
    try:
        for geom, p, dim, ww, e, c, z, dig in cases:
            ctx.set_hist_mode("atomics")
            ref_w, ref_c = profile_hist(geom, np.ascontiguousarray(p), dim, np.ascontiguousarray(ww), e, c, z, digitize=dig)
            for other in ("sorted", "chunks"):  # counting-sort tiles; chunked non-returning atomics (the default)
                ctx.set_hist_mode(other)
                hw, hc = profile_hist(geom, np.ascontiguousarray(p), dim, np.ascontiguousarray(ww), e, c, z, digitize=dig)
                assert_array_equal(hc, ref_c)
                assert_array_equal(hw, ref_w)
            assert hc.sum() > 0
    finally:
        ctx.set_hist_mode("chunks")
    # and against numpy for the linspace case
    hw, hc = profile_hist(0, pos32, 2, w, lin)
    for f in range(F):
        assert_array_equal(hc[f], restatement.hist_planar(pos32[f], None, 2, n_bins, lin[f, 0], lin[f, -1]))
        assert_allclose(hw[f], restatement.hist_planar(pos32[f], w[f], 2, n_bins, lin[f, 0], lin[f, -1]), rtol=WTOL,
                        atol=WTOL * np.abs(w[f]).sum())


def test_empty_and_ragged_inputs():
    e = np.linspace(0.0, 1.0, 11)[None]
    hw, hc = profile_hist(0, np.zeros((1, 0, 3), np.float32), 2, np.zeros((1, 0)), e)
    assert not hc.any() and not hw.any()
    pos = np.full((1, 5, 3), np.nan, np.float32)  # NaNs are dropped like numpy's range test does
    pos[0, 0] = 0.55
    hw, hc = profile_hist(0, pos, 2, None, e)
    assert hc.sum() == 1 and hc[0, 5] == 1 and hw is None


def test_full_size_properties():
    """Config-4 size (1M atoms, 2154 bins): counts sum to the in-range atoms, linear in the weights."""
    n, nb = 1_000_000, 2154
    pos = frames(2, n, -1.0, 216.0, np.float32, seed=9)
    w = np.random.default_rng(3).uniform(1, 16, (1, n))
    e = np.tile(np.linspace(np.float64(0), np.float64(215.37), nb + 1), (2, 1))
    hw, hc = profile_hist(0, pos, 2, w, e)
    for f in range(2):
        z = pos[f, :, 2].astype(np.float64)
        inside = (z >= 0) & (z <= 215.37)
        assert hc[f].sum() == inside.sum()
        assert_allclose(hw[f].sum(), w[0][inside].sum(), rtol=1e-12)
        assert_array_equal(hc[f], np.histogram(z, nb, (0, 215.37))[0])
    hw2, _ = profile_hist(0, pos, 2, 3.0 * w, e)
    assert_allclose(hw2, 3.0 * hw, rtol=1e-12)


def uniform_universe(n_atoms=125, n_frames=400, seed=1634123):
    """tests/core/test_planar.py:357-393 (fewer frames)."""
    coords = np.random.default_rng(seed).random((n_frames, n_atoms, 3)).astype(np.float32)
    return Universe(n_atoms, MemoryTrajectory(coords, np.array([2, 2, 2, 90, 90, 90], np.float32)),
                    masses=np.ones(n_atoms), charges=np.ones(n_atoms), resindices=np.arange(n_atoms),
                    segindices=np.arange(n_atoms))


def number_weights(ag, grouping, scale=1):
    return scale * density_weights(ag, grouping, dens="number")


PLANAR = dict(weighting_function=number_weights, weighting_function_kwargs=None, normalization="number", dim=2,
              zmin=None, zmax=None, bin_width=0.1, refgroup=None, sym=False, sym_odd=False, grouping="atoms",
              unwrap=False, pack=True, bin_method="com", concfreq=0, output="profile.dat", jitter=False)


@pytest.mark.parametrize("normalization", ["volume", "number", "None"])
def test_planar_profile_known_answers(normalization):
    """tests/core/test_planar.py:451-473."""
    u = uniform_universe()
    p = ProfilePlanarBase(atomgroup=u.atoms, **{**PLANAR, "normalization": normalization}).run()
    vals = {"volume": 125 / (8 / 2), "number": 1, "None": 125 * p.n_bins / 2 / 100}[normalization]
    desired = np.zeros(p.n_bins)
    desired[:10] = vals
    if normalization == "number":
        desired[10:] = np.nan
    assert_allclose(p.results.profile.flatten(), desired, rtol=3e-2)
    assert p.results.profile.dtype == np.float64


def test_compute_histogram_hooks():
    """tests/core/test_planar.py:509-528, test_cylinder.py:529-547, test_sphere.py:531-549."""
    u = uniform_universe(n_frames=2)
    p = ProfilePlanarBase(atomgroup=u.atoms, **PLANAR)
    p._setup_frames(u.trajectory)
    p._prepare()
    pts = np.linspace(3 * [p.zmin], 3 * [p.zmax], p.n_bins)
    assert_array_equal(p._compute_histogram(pts, weights=None), 1)
    assert_array_equal(p._compute_histogram(pts, weights=5 * np.ones(p.n_bins)), 5)
    common = {k: v for k, v in PLANAR.items() if k not in ("sym", "sym_odd", "dim", "zmin", "zmax")}
    c = ProfileCylinderBase(atomgroup=u.atoms, dim=2, zmin=None, zmax=None, rmin=0, rmax=None, **common)
    c._setup_frames(u.trajectory)
    c._prepare()
    r = (np.arange(c.n_bins) + 0.5) * c.rmax / c.n_bins
    pts = np.stack([1.0 + r, np.ones_like(r), np.ones_like(r)], 1)
    assert_array_equal(c._compute_histogram(pts, weights=None), 1)
    s = ProfileSphereBase(atomgroup=u.atoms, rmin=0, rmax=None, **common)
    s._setup_frames(u.trajectory)
    s._prepare()
    assert_array_equal(s._compute_histogram(pts, weights=5 * np.ones(len(pts))), 5)


def test_correlation_bin_and_single_frame():
    """tests/core/test_planar.py:536-545."""
    u = uniform_universe(n_frames=3)
    for n_bins in (1, 2, 3):
        p = ProfilePlanarBase(atomgroup=u.atoms, **{**PLANAR, "bin_width": 2 / n_bins}).run(stop=1)
        assert p._single_frame() == p._obs.profile[n_bins // 2]


def slab_universe(n_frames=5):
    fr, topo = mb.synthetic.confined_slab_frames(3000, 800, 42.0, 40.0, 70.0, 36.0, 20261021, n_frames=n_frames)
    traj = MemoryTrajectory(fr, np.array([42.0, 40.0, 70.0, 90, 90, 90], np.float32))
    return Universe(fr.shape[1], traj, **topo)


def _restated_profile(u, ag, hist_fn, weights_fn, positions_fn, normalization, geometry_fn, n_frames):
    stats = restatement.RunningStats()
    for f in range(n_frames):
        u.trajectory[f]
        obs = geometry_fn()
        obs.update(restatement.profile_single_frame(hist_fn, positions_fn(), weights_fn(), normalization,
                                                    obs["bin_volume"]))
        stats.update(obs)
    return stats


def test_density_planar_class_vs_restatement():
    u = slab_universe()
    water = u.select_atoms("element O H")
    ana = mb.DensityPlanar(water, dens="mass", bin_width=0.1, unwrap=False, pack=False).run()
    nb = ana.n_bins
    assert nb == 700

    def geom():
        return restatement.planar_geometry(np.double(u.dimensions[:3]), 2, nb)[2]

    st = _restated_profile(u, water, lambda p, w: restatement.hist_planar(p, w, 2, nb, 0.0, np.float64(70.0)),
                           lambda: water.masses, lambda: water.positions, "volume", geom, 5)
    assert_array_equal(ana.sums.bincount, st.sums["bincount"])
    assert_allclose(ana.results.profile, st.means["profile"], rtol=1e-12, atol=1e-15)
    assert_allclose(ana.results.dprofile, st.sems["profile"], rtol=1e-9, atol=1e-15)
    assert_allclose(ana.results.bin_pos, st.means["bin_pos"] - 35.0)


def test_diporder_planar_and_collection_vs_restatement():
    u = slab_universe(n_frames=4)
    water = u.select_atoms("element O H")
    dens = mb.DensityPlanar(water, dens="charge", bin_width=0.5, unwrap=False, pack=False, grouping="residues")
    dip = mb.DiporderPlanar(water, bin_width=0.5, unwrap=False, pack=False, order_parameter="cos_theta")
    mb.AnalysisCollection(dens, dip).run()
    nb = dip.n_bins

    def geom():
        return restatement.planar_geometry(np.double(u.dimensions[:3]), 2, nb)[2]

    def w_cos():
        mu = water.dipole_vector(compound="residues")
        return (mu / np.linalg.norm(mu, axis=1)[:, None])[:, 2]

    st = _restated_profile(u, water, lambda p, w: restatement.hist_planar(p, w, 2, nb, 0.0, np.float64(70.0)),
                           w_cos, lambda: water.center(water.masses, compound="residues"), "number", geom, 4)
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = st.sums["profile"] / st.sums["bincount"]
    assert_array_equal(dip.sums.bincount, st.sums["bincount"])
    assert_allclose(dip.results.profile, ref, rtol=1e-10, atol=1e-14, equal_nan=True)
    solo = mb.DensityPlanar(water, dens="charge", bin_width=0.5, unwrap=False, pack=False, grouping="residues").run()
    assert_allclose(dens.results.profile, solo.results.profile, rtol=1e-12, atol=1e-15)


def test_density_cylinder_and_sphere_classes():
    u = slab_universe(n_frames=3)
    water = u.select_atoms("element O H")
    cyl = mb.DensityCylinder(water, dens="number", bin_width=0.5, unwrap=False, pack=False).run()
    sph = mb.DensitySphere(water, dens="mass", bin_width=0.5, unwrap=False, pack=False).run()
    center = np.array([21.0, 20.0, 35.0])
    st_c, st_s = restatement.RunningStats(), restatement.RunningStats()
    for f in range(3):
        u.trajectory[f]
        p = water.positions
        st_c.update({"c": restatement.hist_cylinder(p, None, 2, cyl.n_bins, 0.0, 20.0, 0.0, 70.0, center)})
        st_s.update({"c": restatement.hist_sphere(p, None, sph.n_bins, 0.0, 20.0, center)})
    assert cyl.rmax == 20.0 and sph.rmax == 20.0
    assert_array_equal(cyl.sums.bincount, st_c.sums["c"])
    assert_array_equal(sph.sums.bincount, st_s.sums["c"])
    edges = np.linspace(0.0, 20.0, cyl.n_bins + 1)
    vol = np.pi * np.diff(edges**2) * 70.0
    assert_allclose(cyl.results.profile, st_c.means["c"] / vol, rtol=1e-12)


def test_device_compound_prologue_equals_host_prologue(monkeypatch):
    """mb200_compound_prologue (unwrap, wrap, centres, dipole weights) vs the shim's host path: bit for bit."""
    import maicos_b200.core.base as base
    import maicos_b200.modules.diporderstructurefactor as dsf

    # molecules scattered over periodic images and broken across the cell faces
    u = mb.synthetic.spce_universe(1500, 77, n_frames=4)
    L = u.dimensions[:3]
    rng = np.random.default_rng(4)
    for f in range(4):
        ts = u.trajectory[f]
        u.trajectory.coordinates[f] = ts.positions + L * rng.integers(-1, 2, ts.positions.shape)

    def run_all():
        out = {}
        for op in ("P0", "cos_theta", "cos_2_theta"):
            a = mb.DiporderPlanar(u.atoms, bin_width=0.7, order_parameter=op, pdim=1).run()
            out[op] = (a.results.profile.copy(), a.sums.bincount.copy(), a._prologue is not None)
        a = mb.DensityPlanar(u.atoms, dens="mass", grouping="residues", bin_method="coc", bin_width=0.9).run()
        out["dens"] = (a.results.profile.copy(), a.sums.bincount.copy(), a._prologue is not None)
        a = mb.DensitySphere(u.atoms, dens="charge", grouping="molecules", bin_method="cog", bin_width=1.1, unwrap=False).run()
        out["sph"] = (a.results.profile.copy(), a.sums.bincount.copy(), a._prologue is not None)
        a = mb.DiporderStructureFactor(u.atoms, qmax=2.0, dq=0.1, grouping="residues").run()
        out["dsf"] = (a.results.structure_factors.copy(), np.zeros(1), a._prologue is not None)
        return out

    dev = run_all()
    assert all(v[2] for v in dev.values()), "device prologue should be active for whole contiguous molecules"
    monkeypatch.setattr(base, "compound_runs", lambda ag, c: None)
    monkeypatch.setattr(dsf, "compound_runs", lambda ag, c: None)
    host = run_all()
    assert not any(v[2] for v in host.values())
    for k in dev:
        assert_array_equal(dev[k][1], host[k][1])
        assert_allclose(dev[k][0], host[k][0], rtol=1e-12, atol=1e-15, equal_nan=True)


def test_device_prologue_not_used_for_partial_compounds():
    u = slab_universe(n_frames=2)
    oxy = u.select_atoms("element O")  # one atom of every residue: incomplete compounds
    a = mb.DensityPlanar(oxy, dens="mass", grouping="residues", bin_width=1.0, unwrap=False, pack=False).run()
    assert a._prologue is None and a.sums.bincount.sum() > 0
