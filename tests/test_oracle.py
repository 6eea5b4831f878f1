"""The oracle against the reference's own golden vectors and against the compiled reference kernel.

CPU only.  Citations: reference tests under ``/root/reference/tests`` (MAICoS v0.11.2).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal
from conftest import compact_to_cube, load_golden

from oracle import kernel, restatement


def _sorted_nonzero(q, s):
    q, s = q.flatten(), s.flatten()
    nz = np.where(s != 0)[0]
    o = np.argsort(q[nz], kind="stable")
    return q[nz][o], s[nz][o]


KERNELS = [("port", kernel.port_structure_factor)]
if kernel.have_reference():
    KERNELS.append(("reference", kernel.reference_structure_factor))


@pytest.mark.parametrize("name,sf", KERNELS)
def test_sfactor_angle_table(water_gro, name, sf):
    """tests/lib/test_math.py:183-209 -- six q vectors in the theta window, rtol 1e-1 there."""
    pos = np.double(water_gro["positions"])
    box = np.double(water_gro["dimensions"][:3])
    q, s = _sorted_nonzero(*sf(pos, box, 0, 0.5, np.pi / 4, np.pi / 2, np.ones(len(pos))))
    assert_allclose(q, [0.25, 0.25, 0.36, 0.36, 0.36, 0.44], rtol=1e-1)
    assert_allclose(s, [6.91, 835.3, 192.54, 157.79, 506.76, 99.65], rtol=1e-1)
    # the same call, to the printed digits of the reference's table (sorted within equal |q|)
    assert_allclose(np.sort(s), np.sort([6.91, 835.3, 192.54, 157.79, 506.76, 99.65]), rtol=5e-4)


@pytest.mark.parametrize("name,sf", KERNELS)
def test_sfactor_50_entry_table(water_gro, name, sf):
    """tests/lib/test_math.py:34-147 -- the 50 smallest q vectors for qmax = 1 (q rtol 1e-2)."""
    q_tab = np.array([0.25] * 3 + [0.36] * 3 + [0.44] + [0.51] * 3 + [0.56] * 6 + [0.62] * 3 + [0.71] * 3
                     + [0.76] * 6 + [0.8] * 6 + [0.84] * 3 + [0.88] + [0.91] * 6 + [0.95] * 6)
    s_tab = np.array([186.43, 6.91, 835.3, 506.76, 192.54, 157.79, 99.65, 587.47, 788.63, 518.17, 458.65, 1.69,
                      399.91, 610.34, 1213.59, 41.18, 931.98, 629.12, 98.85, 315.22, 100.84, 119.42, 213.18, 461.77,
                      399.64, 803.88, 174.83, 32.09, 199.19, 424.69, 1735.52, 1377.32, 125.05, 261.75, 429.61, 20.9,
                      271.45, 422.34, 107.59, 379.52, 6.69, 535.33, 109.21, 669.97, 1253.54, 394.2, 196.1, 139.89,
                      87.96, 417.02])
    pos = np.double(water_gro["positions"])
    box = np.double(water_gro["dimensions"][:3])
    q, s = _sorted_nonzero(*sf(pos, box, 0, 1.0, 0, np.pi, np.ones(len(pos))))
    assert_allclose(q[:50], q_tab, atol=6e-3)  # table printed to two decimals
    # the S column is documentation in the reference (its comparison never runs); as multisets per
    # |q| shell the values are reproduced to the printed precision
    for shell in np.unique(q_tab):
        m = q_tab == shell
        assert_allclose(np.sort(s[:50][m]), np.sort(s_tab[m]), rtol=2e-3)


@pytest.mark.parametrize("case", ["theta_window", "qmax1", "qmin_window", "default_saxs", "ortho_signed"])
def test_port_matches_reference_fixture(water_gro, case):
    """C port == outputs of the reference's compiled kernel stored in tests/golden (mask exact)."""
    g = load_golden("sfactor_reference.npz")
    q_ref, s_ref = compact_to_cube(g, case)
    qmin, qmax, tmin, tmax = g[f"{case}__args"]
    if case == "ortho_signed":
        pos, box, w = g[f"{case}__positions"], g[f"{case}__box"], g[f"{case}__weights"]
    else:
        pos, box = np.double(water_gro["positions"]), np.double(water_gro["dimensions"][:3])
        w = np.ones(len(pos))
    q, s = kernel.port_structure_factor(pos, box, qmin, qmax, tmin, tmax, w)
    assert q.shape == q_ref.shape
    assert_array_equal(s != 0, s_ref != 0)
    assert_array_equal(q, q_ref)
    nz = s_ref != 0
    assert_allclose(s[nz], s_ref[nz], rtol=1e-9)


@pytest.mark.skipif(not kernel.have_reference(), reason="oracle/_ref not built")
def test_port_matches_live_reference_random():
    rng = np.random.default_rng(7)
    for n, L, qmax in [(1, [9.0, 9, 9], 2.0), (50, [12.0, 15.0, 11.0], 2.5), (3000, [30.0, 31.0, 29.0], 1.2)]:
        pos = rng.uniform(-5, 35, (n, 3))
        w = rng.normal(size=n)
        a = kernel.reference_structure_factor(pos, np.array(L), 0.1, qmax, 0.2, 2.9, w)
        b = kernel.port_structure_factor(pos, np.array(L), 0.1, qmax, 0.2, 2.9, w)
        assert_array_equal(a[0], b[0])
        assert_array_equal(a[1] != 0, b[1] != 0)
        nz = a[1] != 0
        assert_allclose(b[1][nz], a[1][nz], rtol=1e-9)


def test_maxn_uses_float32_divisor():
    """_cmath.pyx:96 -- the float32 cast can change maxn by one against the float64 rule."""
    hits = 0
    rng = np.random.default_rng(3)
    for _ in range(20000):
        L = rng.uniform(10, 300)
        k = rng.integers(1, 200)
        qmax = k * 2 * np.pi / L  # exactly on a lattice point in exact arithmetic
        m32 = kernel.port_maxn(np.array([L, L, L]), qmax)[0]
        m64 = int(np.ceil(qmax / (2 * np.pi / L)))
        assert m32 == int(np.ceil(qmax / np.float64(np.float32(2 * np.pi / L))))
        hits += m32 != m64
    assert hits > 0


def test_saxs_one_frame_golden(water_gro):
    """tests/modules/test_saxs.py:57-71 -- I[0] == 1.6047 (rtol 1e-3) for water.gro, qmax = 20."""
    res, _, _ = restatement.run_saxs(water_gro["positions"][None], water_gro["dimensions"][None, :3],
                                     water_gro["elements"], qmax=20)
    assert_allclose(res["scattering_intensities"][0], 1.6047, rtol=1e-3)
    g = load_golden("saxs_water_gro_qmax20.npz")
    for k in ("scattering_vectors", "structure_factors", "scattering_intensities"):
        assert_allclose(res[k], g[k], rtol=1e-9)


def test_saxs_unbinned_restatement_bins_to_the_binned_one(water_gro):
    """``bin_spectrum=False`` (saxs.py:188-189, 238-239) keeps every Miller vector; histogramming its cubes by |q| per
    group is the binned frame body, so the summed cubes binned over their nonzero cells give the binned observables when
    every accepted S is nonzero (one frame of water.gro, qmax = 2; tests/modules/test_saxs.py:90-107 is structural)."""
    pos = restatement.wrap_atoms_f32(water_gro["positions"], water_gro["dimensions"][:3])
    box = np.asarray(water_gro["dimensions"][:3], dtype=np.float32)
    el = water_gro["elements"]
    max_n = tuple(np.ceil(2 / (2 * np.pi / np.double(box))).astype(int))
    cubes, last = restatement.saxs_single_frame_unbinned(pos, box, el, 0, 2, 0, np.pi, max_n)
    binned, _ = restatement.saxs_single_frame(pos, box, el, 0, 2, 0, np.pi, 20)
    q, s_h = kernel.structure_factor(np.double(pos[el == "O"]), np.double(box), 0, 2, 0, np.pi, np.ones((el == "O").sum()))
    assert cubes["structure_factors"].shape == q.shape and last == s_h.flatten()[-1]  # O sorts last: H, O
    nz = cubes["structure_factors"].flatten() != 0
    for k in ("structure_factors", "scattering_intensities"):
        h, _ = np.histogram(q.flatten()[nz], bins=20, range=(0, 2), weights=cubes[k].flatten()[nz])
        assert_allclose(h, binned[k], rtol=1e-12)
    res = restatement.saxs_conclude_unbinned(cubes, len(el), box, max_n)
    assert_array_equal(res["scattering_vectors"], np.sort(res["scattering_vectors"]))
    assert len(res["miller_indices"]) == int(np.prod(max_n)) and res["structure_factors"][0] == 0  # q = 0 is rejected


def test_saxs_npt_frames_golden(water_frames, water_gro):
    res, stats, series = restatement.run_saxs(water_frames["positions"], water_frames["dimensions"][:, :3],
                                              water_gro["elements"])
    g = load_golden("saxs_water_trr_default.npz")
    for k in ("scattering_vectors", "structure_factors", "scattering_intensities", "dstructure_factors"):
        assert_allclose(res[k], g[k], rtol=1e-9)
    assert_array_equal(stats.sums["bincount"], g["sum_bincount"])


def test_form_factor_tables():
    """tests/lib/test_math.py:605-672."""
    ref = np.array([[0.00, 1.000], [0.01, 0.998], [0.02, 0.991], [0.03, 0.980], [0.04, 0.966], [0.05, 0.947],
                    [0.06, 0.925], [0.07, 0.900], [0.08, 0.872], [0.09, 0.842], [0.10, 0.811], [0.22, 0.424],
                    [0.46, 0.090]])
    assert_allclose(restatement.atomic_form_factor(4 * np.pi * ref[:, 0], "H"), ref[:, 1], rtol=5e-3)
    for el, n in [("C", 6), ("Cval", 6), ("CVAL", 6), ("O", 8), ("CH1", 7), ("CH2", 8), ("CH3", 9), ("CH4", 10),
                  ("NH1", 8), ("NH2", 9), ("NH3", 10)]:
        assert_allclose(restatement.atomic_form_factor(0.0, el), n, rtol=1e-3)
    with pytest.raises(ValueError, match="Element 'foo' not found"):
        restatement.atomic_form_factor(0.0, "foo")


def test_histogram_known_answers():
    """tests/core/test_planar.py:509-528, test_cylinder.py:529-547, test_sphere.py:531-549."""
    n_bins, zmin, zmax = 20, np.float64(0), np.float64(2)
    pos = np.linspace(3 * [zmin], 3 * [zmax], n_bins)
    assert_array_equal(restatement.hist_planar(pos, None, 2, n_bins, zmin, zmax), 1)
    assert_array_equal(restatement.hist_planar(pos, 5 * np.ones(n_bins), 2, n_bins, zmin, zmax), 5)
    # sphere: points on the diagonal at radii inside each shell
    center = np.array([1.0, 1.0, 1.0])
    r = (np.arange(10) + 0.5) * 0.1
    p = center + np.outer(r, np.ones(3) / np.sqrt(3))
    assert_array_equal(restatement.hist_sphere(p, None, 10, 0.0, 1.0, center), 1)


def test_running_stats_contract():
    """tests/core/test_base.py:222-229: sems == std / sqrt(n), means, sums."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(57, 5))
    st = restatement.RunningStats()
    for row in x:
        st.update({"o": row, "c": float(row[0])})
    assert_allclose(st.sums["o"], x.sum(0))
    assert_allclose(st.means["o"], x.mean(0))
    assert_allclose(st.sems["o"], x.std(0) / np.sqrt(len(x)))
    assert_allclose(st.sems["c"], x[:, 0].std() / np.sqrt(len(x)))


def test_shim_semantics_reproduce_reference_diporder_planar_values(water_gro):
    """tests/modules/test_diporderplanar.py:41-82: 45 regression numbers produced by MAICoS + MDAnalysis.

    Pins the in-memory shim's unwrap / refgroup centring / compound wrap / centre-of-mass / dipole
    semantics (MDAnalysis itself is absent) together with the restated planar histogram and its
    normalisations (core/base.py:436-451, 804-835)."""
    from conftest import water_universe
    from maicos_b200.lib.math import center_cluster
    from reference_values import DIPORDER_PLANAR

    for dim in range(3):
        u = water_universe(water_gro["positions"], water_gro["dimensions"], water_gro["elements"])
        ag = u.atoms
        ag.unwrap(compound="residues")
        L = u.dimensions[:3].astype(np.float64)
        ag.translate(L / 2 - center_cluster(ag, ag.masses))
        ag.wrap(compound="residues")
        pos = ag.center(ag.masses, compound="residues")
        mu = ag.dipole_vector(compound="residues")
        n_bins = int(np.ceil(L[dim] / 5))
        zmin, zmax, obs = restatement.planar_geometry(L, dim, n_bins)
        cnt = restatement.hist_planar(pos, None, dim, n_bins, zmin, zmax)
        cos = mu[:, dim] / np.linalg.norm(mu, axis=1)
        for op, w in (("P0", mu[:, dim]), ("cos_theta", cos), ("cos_2_theta", cos * cos)):
            prof = restatement.hist_planar(pos, w, dim, n_bins, zmin, zmax)
            out = prof / obs["bin_volume"] if op == "P0" else prof / cnt
            assert_allclose(out, DIPORDER_PLANAR[dim][op], rtol=1e-2)


def test_density_cylinder_reference_means(water_gro):
    """tests/modules/test_densitycylinder.py:38-48 through the restated cylinder histogram."""
    from conftest import water_universe
    from maicos_b200.lib.weights import density_weights
    from reference_values import DENSITY_CYLINDER_MEAN

    u = water_universe(water_gro["positions"], water_gro["dimensions"], water_gro["elements"])
    u.atoms.wrap(compound="atoms")
    L = u.dimensions[:3].astype(np.float64)
    rmax = L[:2].min() / 2
    nb = int(np.ceil(rmax))
    vol = np.pi * np.diff(np.linspace(0, rmax, nb + 1) ** 2) * L[2]
    for dens, mean in DENSITY_CYLINDER_MEAN.items():
        w = density_weights(u.atoms, "atoms", dens)
        prof = restatement.hist_cylinder(u.atoms.positions, w, 2, nb, 0.0, rmax, 0.0, L[2], L / 2) / vol
        assert_allclose(prof.mean(), mean, atol=1e-4, rtol=1e-2)


@pytest.mark.parametrize("orientation,M_par,M_perp", [([0, 0, 1], [0, 0], 1), ([1, 0, 0], [1, 0], 0), ([0, 1, 0], [0, 1], 0),
                                                      ([1, 1, 1], [1 / np.sqrt(3), 1 / np.sqrt(3)], 1 / np.sqrt(3))])
def test_dielectric_restatement_dipole_known_answers(orientation, M_par, M_perp):
    """Pins oracle.restatement.dielectric_planar_frame: tests/modules/test_dielectricplanar.py:104-196 (single dipole and
    5 x 5 x 5 grid: total dipole moment, and the volume integral of the virtual-cut dipole density equals it)."""
    from conftest import dipole_universe

    xx, yy, zz = np.meshgrid(np.arange(0, 10, 2) + 1, np.arange(0, 10, 2) + 1, np.arange(0, 10, 2) + 1)
    grid = np.array([xx.flatten(), yy.flatten(), zz.flatten()]).T.tolist()
    for centres in ([[5, 5, 5]], grid):
        u = dipole_universe(centres, [orientation] * len(centres))
        pos, q = u.atoms.positions, np.asarray(u.atoms.charges, dtype=np.float64)
        inv = np.unique(u.atoms.molnums, return_inverse=True)[1]
        obs, _ = restatement.dielectric_planar_frame(pos, q, inv, pos, q, u.dimensions, 2, 1000, 0.01)
        n = len(centres)
        assert_allclose(obs["M_par"], np.multiply(M_par, n), rtol=0.1, atol=1e-5)
        assert_allclose(obs["M_perp"], np.multiply(M_perp, n), rtol=0.1, atol=1e-5)
        bv = obs["bin_volume"][0]
        assert_allclose(np.sum(obs["m_par"], axis=0) * bv, np.multiply(M_par, n), rtol=0.1, atol=1e-5)
        assert_allclose(np.sum(obs["m_perp"], axis=0) * bv, np.multiply(M_perp, n), rtol=0.1, atol=1e-5)
