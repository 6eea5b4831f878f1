"""Parity of the CUDA structure-factor path against the oracle (all calls go through the C ABI).

Tolerances: S(q) relative error <= 1e-6 on every nonzero cell (north_star), the nonzero mask, |q| values and
bin counts bit-exact.  Every test runs in both arithmetic modes of the contraction: "i8x5" (default: split
precision on the tcgen05 tensor cores, measured <= 5e-9 per cell) and "fp64" (FP64 DMMA, measured ~1e-12).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal
from conftest import compact_to_cube, load_golden, water_universe

import maicos_b200 as mb
from maicos_b200 import _cabi
from maicos_b200.lib.math import structure_factor
from oracle import kernel, restatement

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # stated tolerance of the path (BASELINE.json north_star)


@pytest.fixture(autouse=True, params=["i8x5", "fp64"])
def precision(request):
    mb.set_precision(request.param)
    yield request.param
    mb.set_precision("i8x5")


def check(pos, box, qmin, qmax, tmin, tmax, w):
    q1, s1 = kernel.structure_factor(pos, box, qmin, qmax, tmin, tmax, w)
    q2, s2 = structure_factor(pos, box, qmin, qmax, tmin, tmax, w)
    assert q2.shape == q1.shape and q2.dtype == np.float64 and s2.dtype == np.float64
    assert q2.flags.c_contiguous and s2.flags.c_contiguous
    assert_array_equal(q2, q1)
    assert_array_equal(s2 != 0, s1 != 0)
    nz = s1 != 0
    if nz.any():
        assert_allclose(s2[nz], s1[nz], rtol=RTOL)
    return q2, s2


@pytest.mark.parametrize("case", ["theta_window", "qmax1", "qmin_window", "default_saxs", "ortho_signed"])
def test_golden_reference_kernel_outputs(water_gro, case):
    """Against outputs of the reference's own compiled kernel (tests/golden/sfactor_reference.npz)."""
    g = load_golden("sfactor_reference.npz")
    q_ref, s_ref = compact_to_cube(g, case)
    qmin, qmax, tmin, tmax = g[f"{case}__args"]
    if case == "ortho_signed":
        pos, box, w = g[f"{case}__positions"], g[f"{case}__box"], g[f"{case}__weights"]
    else:
        pos, box = np.double(water_gro["positions"]), np.double(water_gro["dimensions"][:3])
        w = np.ones(len(pos))
    q, s = structure_factor(pos, box, qmin, qmax, tmin, tmax, w)
    assert_array_equal(q, q_ref)
    assert_array_equal(s != 0, s_ref != 0)
    nz = s_ref != 0
    assert_allclose(s[nz], s_ref[nz], rtol=RTOL)


def test_reference_theta_table(water_gro):
    """tests/lib/test_math.py:183-209 run against the CUDA kernel."""
    pos, box = np.double(water_gro["positions"]), np.double(water_gro["dimensions"][:3])
    q, s = structure_factor(pos, box, 0, 0.5, np.pi / 4, np.pi / 2, np.ones(len(pos)))
    q, s = q.flatten(), s.flatten()
    nz = np.where(s != 0)[0]
    o = np.argsort(q[nz])
    assert_allclose(q[nz][o], [0.25, 0.25, 0.36, 0.36, 0.36, 0.44], rtol=1e-1)
    assert_allclose(s[nz][o], [6.91, 835.3, 192.54, 157.79, 506.76, 99.65], rtol=1e-1)


@pytest.mark.parametrize("n,box,qmin,qmax,tmin,tmax", [
    (1, [10.0, 10, 10], 0, 2, 0, np.pi),
    (7, [10.0, 10, 10], 0, 2, 0, np.pi),
    (8, [10.0, 11, 12], 0, 2.5, 0, np.pi),
    (9, [10.0, 11, 12], 0.5, 2.5, 0.1, 1.0),
    (1000, [20.0, 25.0, 31.0], 0.3, 3.0, 0.2, 1.2),
    (4097, [40.0, 18.0, 23.0], 0, 1.9, 0, np.pi),
    (100, [50.0] * 3, 0, 0.05, 0, np.pi),       # no valid q at all
    (100, [50.0] * 3, 0, 0.13, 0, np.pi),       # a handful of q
    (33333, [99.97] * 3, 0, 2.0, 0, np.pi),     # oxygen group of config 2
])
def test_random_inputs_against_oracle(n, box, qmin, qmax, tmin, tmax):
    rng = np.random.default_rng(n)
    L = np.array(box)
    pos = rng.uniform(-0.3 * L, 1.4 * L, (n, 3))
    check(pos, L, qmin, qmax, tmin, tmax, np.ones(n))


def test_signed_weights_and_strided_views():
    rng = np.random.default_rng(11)
    n = 5000
    big = rng.uniform(0, 30, (n, 6))
    pos = big[:, ::2]  # non-contiguous view, like a column selection
    w = rng.normal(size=2 * n)[::2]
    assert not pos.flags.c_contiguous
    check(pos, np.array([30.0, 30, 30]), 0, 2, 0, np.pi, w)
    posT = np.asfortranarray(rng.uniform(0, 30, (n, 3)))
    check(posT, np.array([30.0, 30, 30]), 0, 2, 0, np.pi, np.ascontiguousarray(w))


def test_empty_group_and_zero_weights():
    q, s = structure_factor(np.zeros((0, 3)), np.array([10.0, 10, 10]), 0, 2, 0, np.pi, np.zeros(0))
    assert q.shape == (4, 4, 4) and np.count_nonzero(q) > 0 and not s.any()
    # zero weights: S == 0 everywhere -> the reference's "S != 0" mask is empty as well
    pos = np.random.default_rng(0).uniform(0, 10, (64, 3))
    check(pos, np.array([10.0, 10, 10]), 0, 2, 0, np.pi, np.zeros(64))


def test_q_blocks_sum_to_full_result():
    rng = np.random.default_rng(2)
    pos = rng.uniform(0, 40, (3000, 3))
    w = np.ones(3000)
    box = np.array([40.0, 40, 40])
    q, s = structure_factor(pos, box, 0, 2, 0, np.pi, w)
    for nb in (2, 3, 8):
        tot = np.zeros_like(s)
        for b in range(nb):
            qb, sb = structure_factor(pos, box, 0, 2, 0, np.pi, w, block=(b, nb))
            assert_array_equal(qb, q)
            assert np.all((tot == 0) | (sb == 0))  # blocks own disjoint cells
            tot += sb
        assert_array_equal(tot, s)


def test_size_independent_properties_full_size(precision):
    """Config-2 / config-5 sizes: properties instead of the (hours long) CPU oracle.

    S(q) is invariant under a global translation and under periodic images of single atoms, it is
    quadratic in the weights, a single atom gives S == 1 on every accepted q, and S <= (sum |w|)^2."""
    rng = np.random.default_rng(5)
    n, L = 200_000, 125.9
    box = np.array([L, L, L])
    pos = rng.uniform(0, L, (n, 3))
    w = np.ones(n)
    q0, s0 = structure_factor(pos, box, 0, 2.0, 0, np.pi, w)
    nz = s0 != 0
    assert nz.sum() > 30000
    _, s1 = structure_factor(pos + np.array([3.21, -7.5, 11.125]), box, 0, 2.0, 0, np.pi, w)
    assert_allclose(s1[nz], s0[nz], rtol=1e-6, atol=1e-6 * n * 1e-3)
    img = pos + box * rng.integers(-2, 3, (n, 3))
    _, s2 = structure_factor(img, box, 0, 2.0, 0, np.pi, w)
    assert_allclose(s2[nz], s0[nz], rtol=1e-6, atol=1e-6 * n * 1e-3)
    _, s3 = structure_factor(pos, box, 0, 2.0, 0, np.pi, 2.5 * w)
    assert_allclose(s3[nz], 6.25 * s0[nz], rtol=1e-12 if precision == "fp64" else 1e-7)  # i8x5: other digits of 0.625 Ex
    assert s0.max() <= n * n
    q4, s4 = structure_factor(pos[:1], box, 0, 2.0, 0, np.pi, w[:1])
    assert_allclose(s4[q4 != 0], 1.0, rtol=1e-12 if precision == "fp64" else 1e-10)
    # a spot check of 40 cells against a direct NumPy sum
    cells = rng.choice(np.flatnonzero(nz), 40, replace=False)
    hkl = np.stack(np.unravel_index(cells, s0.shape), -1)
    for (h, k, l), c in zip(hkl, cells):
        ph = pos @ (2 * np.pi / box * np.array([h, k, l]))
        ref = np.cos(ph).sum() ** 2 + np.sin(ph).sum() ** 2
        assert_allclose(s0.ravel()[c], ref, rtol=1e-6, atol=1e-6)


def _saxs_frames_cabi(pos, boxes, elements, qmin, qmax, tmin, tmax, n_bins, pack):
    groups = [np.nonzero(elements == e)[0].astype(np.int32) for e in np.unique(elements)]
    ptr = np.concatenate([[0], np.cumsum([len(g) for g in groups])]).astype(np.int64)
    flat = np.concatenate(groups).astype(np.int32)
    from maicos_b200.lib import tables

    cm = np.stack([tables.cm_packed(str(e)) for e in np.unique(elements)])
    F, n, G = pos.shape[0], pos.shape[1], len(groups)
    s = np.zeros((F, G, n_bins)); i = np.zeros((F, G, n_bins)); c = np.zeros((F, G, n_bins), dtype=np.int64)
    _cabi.check(_cabi.load().mb200_saxs_frames(
        _cabi.context().handle, _cabi.ptr(pos), 0, F, n, _cabi.ptr(boxes), _cabi.ptr(flat), _cabi.ptr(ptr), G,
        _cabi.ptr(cm), qmin, qmax, tmin, tmax, n_bins, int(pack), 0, 1, _cabi.ptr(s), _cabi.ptr(i), _cabi.ptr(c)))
    return s, i, c


def test_saxs_frames_against_restatement(water_frames, water_gro):
    """Fused frame body (wrap, S, f^2 S, three |q| histograms) on NPT frames: bin counts exact."""
    pos = np.ascontiguousarray(water_frames["positions"])
    boxes = np.ascontiguousarray(water_frames["dimensions"][:, :3])
    el = water_gro["elements"]
    for (qmin, qmax, dq, tmin, tmax) in [(0, 6, 0.1, 0, np.pi), (0.5, 3.0, 0.05, 0.3, 2.0)]:
        n_bins = int(np.ceil((qmax - qmin) / dq))
        s, i, c = _saxs_frames_cabi(pos, boxes, el, qmin, qmax, tmin, tmax, n_bins, pack=True)
        for f in range(len(pos)):
            p = restatement.wrap_atoms_f32(pos[f], boxes[f])
            obs, val = restatement.saxs_single_frame(p, boxes[f], el, qmin, qmax, tmin, tmax, n_bins)
            assert_array_equal(c[f, -1], obs["bincount"])
            assert_allclose(s[f].sum(0), obs["structure_factors"], rtol=RTOL)
            assert_allclose(i[f].sum(0), obs["scattering_intensities"], rtol=RTOL)
            assert_allclose(s[f, -1, -1], val, rtol=RTOL)


def test_saxs_class_golden(water_gro, water_frames, tmp_path, monkeypatch):
    """Saxs.run() end to end vs the restatement driven by the reference kernel (tests/golden)."""
    monkeypatch.chdir(tmp_path)
    u = water_universe(water_gro["positions"], water_gro["dimensions"], water_gro["elements"])
    saxs = mb.Saxs(u.atoms, qmax=20, output="foo").run()
    assert_allclose(saxs.results.scattering_intensities[0], 1.6047, rtol=1e-3)  # tests/modules/test_saxs.py:57-71
    g = load_golden("saxs_water_gro_qmax20.npz")
    for k in ("scattering_vectors", "structure_factors", "scattering_intensities"):
        assert_allclose(saxs.results[k], g[k], rtol=RTOL)
    saxs.save()
    loaded = np.loadtxt("foo.dat")
    assert_allclose(saxs.results.scattering_vectors, loaded[:, 0])
    assert_allclose(saxs.results.structure_factors, loaded[:, 1])
    assert_allclose(saxs.results.scattering_intensities, loaded[:, 2])

    u = water_universe(water_frames["positions"], water_frames["dimensions"], water_gro["elements"])
    saxs = mb.Saxs(u.atoms).run()
    g = load_golden("saxs_water_trr_default.npz")
    for k in ("scattering_vectors", "structure_factors", "scattering_intensities", "dstructure_factors",
              "dscattering_intensities"):
        assert_allclose(saxs.results[k], g[k], rtol=RTOL, atol=1e-12)
    assert_array_equal(saxs.sums.bincount, g["sum_bincount"])
    assert_allclose(saxs.timeseries, g["series"], rtol=RTOL)
    # a frame subset and the single-frame hook
    sub = mb.Saxs(u.atoms).run(start=1, step=2)
    assert sub.n_frames == 2 and np.all(np.isfinite(sub.results.structure_factors))
    assert_array_equal(sub.frames, [1, 3])
    v = sub._single_frame()  # the hook stays callable on its own: a batch of one frame
    assert np.isfinite(v) and sub._obs.structure_factors.shape == (sub.n_bins,)


def test_saxs_unbinned(water_gro, tmp_path, monkeypatch):
    """tests/modules/test_saxs.py:90-107."""
    monkeypatch.chdir(tmp_path)
    u = water_universe(water_gro["positions"], water_gro["dimensions"], water_gro["elements"])
    saxs = mb.Saxs(u.atoms, bin_spectrum=False, qmax=2, output="foo").run()
    assert type(saxs.scattering_vector_factors).__name__ == "ndarray"
    assert_array_equal(saxs.results.scattering_vectors, np.sort(saxs.results.scattering_vectors))
    saxs.save()
    loaded = np.loadtxt("foo.dat")
    assert_allclose(saxs.results.miller_indices, loaded[:, 1:4])
    assert_allclose(saxs.results.structure_factors, loaded[:, 4])


def test_dipsf_frames_against_restatement():
    """DiporderStructureFactor: fused three-component call and the class, vs the restatement."""
    u = mb.synthetic.spce_universe(600, 20261020, n_frames=3)
    ana = mb.DiporderStructureFactor(u.atoms, qmax=3.0, dq=0.05, grouping="residues").run()
    stats = restatement.RunningStats()
    for f in range(3):
        u.trajectory[f]
        u.atoms.unwrap(compound="residues")
        u.atoms.wrap(compound="residues")
        pos = u.atoms.center(u.atoms.masses, compound="residues")
        mu = u.atoms.dipole_vector(compound="residues")
        w3 = np.ascontiguousarray((mu / np.linalg.norm(mu, axis=1)[:, None]).T)
        obs, _ = restatement.dipsf_single_frame(pos, w3, np.double(u.dimensions[:3]), 0, 3.0, ana.n_bins)
        stats.update(obs)
    keep = np.where(stats.means["structure_factors"] != 0)[0]
    assert_allclose(ana.results.structure_factors, stats.means["structure_factors"][keep], rtol=RTOL)
    q = np.arange(0, 3.0, 0.05) + 0.025
    assert_allclose(ana.results.scattering_vectors, q[keep])


def test_split_precision_against_fp64_mode(water_gro, water_frames, precision):
    """The two arithmetic modes against each other: binned S, I within 1e-7, counts exact; per-cell S within 1e-7."""
    if precision == "fp64":
        pytest.skip("one comparison of the two modes is enough")
    try:
        u = water_universe(water_frames["positions"], water_frames["dimensions"], water_gro["elements"])
        saxs = mb.Saxs(u.atoms).run()
        g = load_golden("saxs_water_trr_default.npz")
        for k in ("scattering_vectors", "structure_factors", "scattering_intensities", "dstructure_factors"):
            assert_allclose(saxs.results[k], g[k], rtol=1e-7, atol=1e-12)
        assert_array_equal(saxs.sums.bincount, g["sum_bincount"])
        # theta window + qmin, bigger box, three-component DiporderStructureFactor
        u2 = mb.synthetic.spce_universe(3000, 20261020, n_frames=2)
        a_tc = mb.Saxs(u2.atoms, qmin=0.3, qmax=3.0, dq=0.05, thetamin=20, thetamax=100).run()
        d_tc = mb.DiporderStructureFactor(u2.atoms, qmax=3.0, dq=0.05, grouping="residues").run()
        pos = np.double(u2.atoms.positions)
        box = np.double(u2.dimensions[:3])
        w = np.random.default_rng(3).normal(size=len(pos)) * 7.0  # |w| > 1: exact power-of-two rescaling inside
        q_tc, s_tc = structure_factor(pos, box, 0, 3.0, 0, np.pi, w)
        mb.set_precision("fp64")
        a_64 = mb.Saxs(u2.atoms, qmin=0.3, qmax=3.0, dq=0.05, thetamin=20, thetamax=100).run()
        d_64 = mb.DiporderStructureFactor(u2.atoms, qmax=3.0, dq=0.05, grouping="residues").run()
        q_64, s_64 = structure_factor(pos, box, 0, 3.0, 0, np.pi, w)
        assert_array_equal(a_tc.sums.bincount, a_64.sums.bincount)
        assert_allclose(a_tc.results.scattering_intensities, a_64.results.scattering_intensities, rtol=1e-7)
        assert_allclose(d_tc.results.structure_factors, d_64.results.structure_factors, rtol=1e-7)
        assert_array_equal(q_tc, q_64)
        assert_array_equal(s_tc != 0, s_64 != 0)
        nz = s_64 != 0
        assert_allclose(s_tc[nz], s_64[nz], rtol=1e-7)
        # non-finite weights cannot be cut into digits: such a call takes the FP64 kernel in either mode
        mb.set_precision("i8x5")
        w_bad = np.ones(len(pos)); w_bad[5] = np.inf
        _, s_bad = structure_factor(pos, box, 0, 1.0, 0, np.pi, w_bad)
        assert not np.isfinite(s_bad[s_bad != 0]).any()
    finally:
        mb.set_precision("i8x5")


def test_batched_pipeline_is_independent_of_the_batch_size():
    """Frame staging, queued device calls and the announced host->device prefetch (mb200_prefetch_frames): seven frames in
    batches of two (four device calls, three of them prefetched under the previous call) give the numbers of one batch."""
    u = mb.synthetic.spce_universe(900, 20261026, n_frames=7)
    ref = mb.Saxs(u.atoms, qmax=2.5).run()
    old = mb.Saxs._batch_size
    try:
        mb.Saxs._batch_size = 2
        got = mb.Saxs(u.atoms, qmax=2.5).run()
        mb.Saxs._batch_size = 1
        one = mb.Saxs(u.atoms, qmax=2.5).run(step=2)
    finally:
        mb.Saxs._batch_size = old
    for k in ("structure_factors", "scattering_intensities", "dstructure_factors", "dscattering_intensities"):
        assert_array_equal(got.results[k], ref.results[k])
    assert_array_equal(got.sums.bincount, ref.sums.bincount)
    assert_array_equal(got.timeseries, ref.timeseries)
    assert_array_equal(one.frames, [0, 2, 4, 6])
    assert_allclose(one.timeseries, ref.timeseries[::2], rtol=0, atol=0)


@pytest.mark.parametrize("n,box,qmax", [
    (1, (21.0, 22.0, 23.0), 2.0),           # one atom: every K-step but the first is padding
    (17, (30.0, 30.0, 30.0), 3.0),          # ragged: 16 + 1 atoms
    (4097, (25.0, 31.0, 37.0), 4.0),        # maxn up to 24: l tile of 24 values
    (9000, (99.97, 99.97, 99.97), 2.0),     # the bench shape: maxn 32, l tile exactly 32 (the limit of the TMEM variant)
    (3000, (103.0, 99.0, 104.0), 2.0),      # l extent 34 > 32: both settings take the shared-memory kernel
])
def test_tmem_operand_variant_bitwise_equals_shared_memory_variant(n, box, qmax, precision):
    """Split-precision contraction with the generated A operand in tensor memory (default when the l tile has at most 32
    values) vs the same arithmetic through shared memory: exact int32 accumulation, so the raw S cube is bit-identical."""
    if precision != "i8x5":
        pytest.skip("one arithmetic mode only")
    from maicos_b200 import _cabi
    rng = np.random.default_rng(n)
    pos = rng.uniform(-3.0, 1.2 * max(box), (n, 3))
    w = rng.normal(size=n)
    ctx = _cabi.context(None)
    try:
        ctx.set_precision("i8x5-smem")
        q0, s0 = structure_factor(pos, np.double(box), 0.0, qmax, 0.0, np.pi, w)
        ctx.set_precision("i8x5")
        q1, s1 = structure_factor(pos, np.double(box), 0.0, qmax, 0.0, np.pi, w)
    finally:
        ctx.set_precision("i8x5")
    assert_array_equal(q1, q0)
    assert_array_equal(s1, s0)
    assert np.count_nonzero(s1) > 0
    check(pos, np.double(box), 0.0, qmax, 0.0, np.pi, w)
