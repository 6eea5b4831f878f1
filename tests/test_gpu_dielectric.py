"""Dielectric profile modules on the CUDA binning path against the NumPy restatement of the
reference's frame bodies and against the reference's own known answers for synthetic dipoles.

Tolerances: the device accumulates charges exactly (fixed point) while numpy adds them in input order, so the
comparison is relative to the scale of the quantity: 1e-10 of max|value| (measured ~1e-16); covariances cancel, so the
epsilon profiles are compared relative to pref * max|<mM>|.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal
from conftest import dipole_universe

import maicos_b200 as mb
from maicos_b200.lib import _dielectric, _hist
from oracle import restatement

pytestmark = pytest.mark.gpu

ORIENTATIONS = [([0, 0, 1], [0, 0], 1), ([1, 0, 0], [1, 0], 0), ([0, 1, 0], [0, 1], 0),
                ([1, 1, 1], [1 / np.sqrt(3), 1 / np.sqrt(3)], 1 / np.sqrt(3))]


def close(a, b, scale=None, tol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    assert_allclose(a, b, rtol=0, atol=tol * max(s, 1e-300))


@pytest.mark.parametrize("geometry", ["planar", "cylinder", "sphere", "scalar"])
def test_digitize_histogram_against_numpy(geometry):
    """mb200_profile_hist with MB200_BIN_DIGITIZE == bincount(digitize(x, edges[1:-1]), w): nothing dropped."""
    rng = np.random.default_rng(3)
    n, nb = 50_001, 173
    pos = rng.uniform(-8, 38, (n, 3)).astype(np.float32)  # box 0..30: values left and right of every range
    w = rng.normal(size=n)
    center = np.array([15.0, 15.0, 15.0])
    if geometry == "planar":
        edges = np.linspace(np.float64(0), np.float64(30.0), nb + 1)
        x = pos[:, 1].astype(np.float64)
        got = _dielectric.charge_per_bin(_hist.PLANAR, pos, w, edges, dim=1)
    elif geometry == "cylinder":
        edges = np.linspace(np.float64(0.5), np.float64(15.0), nb + 1)
        x = restatement.transform_cylinder(pos, center, 0)[:, 0]
        got = _dielectric.charge_per_bin(_hist.CYLINDER, pos, w, edges, dim=0, center=center)
    elif geometry == "sphere":
        edges = np.linspace(np.float64(0.0), np.float64(15.0), nb + 1)
        x = restatement.transform_sphere(pos, center)[:, 0]
        got = _dielectric.charge_per_bin(_hist.SPHERE, pos, w, edges, center=center)
    else:
        edges = np.linspace(np.float64(0), np.float64(30.0), nb + 1)
        x = pos[:, 2].astype(np.float64)
        x[:7] = [np.nan, -np.inf, np.inf, 0.0, 30.0, edges[5], np.nextafter(edges[5], 0)]
        hw, hc = _hist.profile_hist(_hist.SCALAR, x[None], 0, w[None], edges[None], digitize=True)
        got = hw[0]
        assert_array_equal(hc[0], np.bincount(np.digitize(x, edges[1:-1]), minlength=nb))
    ref = np.bincount(np.digitize(x, edges[1:-1]), weights=w, minlength=nb)
    close(got, ref, scale=np.abs(w).sum(), tol=1e-14)


@pytest.mark.parametrize("orientation,M_par,M_perp", ORIENTATIONS)
def test_single_dipole_orientations(orientation, M_par, M_perp):
    """tests/modules/test_dielectricplanar.py:104-137: total dipole moment and volume integral of the density."""
    u = dipole_universe([[5, 5, 5]], [orientation])
    eps = mb.DielectricPlanar(u.atoms, bin_width=0.01, vcutwidth=0.01).run()
    assert_allclose(eps._obs.M_par, M_par, rtol=0.1, atol=1e-6)
    assert_allclose(eps._obs.M_perp, M_perp, rtol=0.1, atol=1e-6)
    bin_volume = eps.means.bin_volume[0]
    assert_allclose(np.sum(eps._obs.m_par, axis=0) * bin_volume, M_par, rtol=0.1, atol=1e-6)
    assert_allclose(np.sum(eps._obs.m_perp, axis=0) * bin_volume, M_perp, rtol=0.1, atol=1e-6)


@pytest.mark.parametrize("selection", [1, 2])
@pytest.mark.parametrize("orientation,M_par,M_perp", ORIENTATIONS)
def test_multiple_dipole_orientations(selection, orientation, M_par, M_perp):
    """tests/modules/test_dielectricplanar.py:139-196: 5 x 5 x 5 dipoles, whole system or half of it selected."""
    xx, yy, zz = np.meshgrid(np.arange(0, 10, 2) + 1, np.arange(0, 10, 2) + 1, np.arange(0, 10, 2) + 1)
    pos = np.array([xx.flatten(), yy.flatten(), zz.flatten()]).T.tolist()
    n_dipoles = len(pos)
    ag = dipole_universe(pos, [orientation] * n_dipoles).atoms
    n = int(len(ag) / selection) + 1 if selection == 2 else len(ag)
    eps = mb.DielectricPlanar(ag[:n], bin_width=0.01, vcutwidth=0.01).run()
    assert_allclose(eps._obs.M_par, np.multiply(M_par, n_dipoles), rtol=0.1, atol=1e-5)
    assert_allclose(eps._obs.M_perp, np.multiply(M_perp, n_dipoles), rtol=0.1, atol=1e-5)
    bin_volume = eps.means.bin_volume[0]
    assert_allclose(np.sum(eps._obs.m_par, axis=0) * bin_volume, np.multiply(M_par, n_dipoles / selection), rtol=0.1, atol=1e-5)
    assert_allclose(np.sum(eps._obs.m_perp, axis=0) * bin_volume, np.multiply(M_perp, n_dipoles / selection), rtol=0.1, atol=1e-5)


def _transformed_frames(u, compound):
    """Positions as the analysis sees them: unwrap -> wrap by compound (core/base.py:436-448)."""
    out = []
    for f in range(u.trajectory.n_frames):
        u.trajectory[f]
        u.atoms.unwrap(compound=compound)
        u.atoms.wrap(compound=compound)
        out.append(u.atoms.positions.copy())
    return out


def test_dielectric_planar_against_restatement(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    u = mb.synthetic.spce_universe(700, 20261023, n_frames=3)
    q = np.asarray(u.atoms.charges, dtype=np.float64)
    inv = np.unique(u.atoms.molnums, return_inverse=True)[1]
    for dim in (2, 0):
        eps = mb.DielectricPlanar(u.atoms, dim=dim, bin_width=0.5, vcutwidth=0.1).run()
        stats = restatement.RunningStats()
        for p in _transformed_frames(u, "molecules"):
            obs, _ = restatement.dielectric_planar_frame(p, q, inv, p, q, u.dimensions, dim, eps.n_bins, 0.1)
            stats.update({k: v for k, v in obs.items() if k not in ("L", "bin_edges")})
        for k in ("m_perp", "mM_perp", "mm_perp", "cmM_perp", "cM_perp", "m_par", "mM_par", "mm_par", "cmM_par",
                  "cM_par", "M", "M_perp_2"):
            close(eps.means[k], stats.means[k])
            close(eps.sems[k], stats.sems[k], scale=np.max(np.abs(stats.means[k])))
        pref = eps._pref
        cov_perp = stats.means["mM_perp"] - stats.means["m_perp"] * stats.means["M_perp"]
        close(eps.results.eps_perp, -pref * cov_perp, scale=pref * np.abs(stats.means["mM_perp"]).max())  # cov cancels
        cov_par = 0.5 * (stats.means["mM_par"] - np.dot(stats.means["m_par"], stats.means["M_par"]))
        close(eps.results.eps_par, pref * cov_par, scale=pref * np.abs(stats.means["mM_par"]).max())
    eps.save()
    res_perp = np.loadtxt("eps_perp.dat")
    assert_allclose(eps.results.eps_perp, res_perp[:, 1])
    res_par = np.loadtxt("eps_par.dat")
    assert_allclose(eps.results.eps_par, res_par[:, 1])
    # sym and is_3d variants run and stay finite; symmetrised profiles are the mean of the profile and its mirror image
    e_sym = mb.DielectricPlanar(u.atoms, sym=True).run()
    e_3d = mb.DielectricPlanar(u.atoms, is_3d=True).run()
    e_0 = mb.DielectricPlanar(u.atoms).run()
    assert_allclose(e_sym.results.eps_par, (e_0.results.eps_par + e_0.results.eps_par[::-1]) / 2, rtol=1e-12, atol=1e-15)
    assert np.all(np.isfinite(e_3d.results.eps_perp))


def test_unsorted_atomgroup_takes_host_prologue():
    """tests/modules/test_dielectricplanar.py:208-223: shuffled atom order (LAMMPS) gives the same profiles."""
    u = mb.synthetic.spce_universe(300, 20261024, n_frames=2)
    permute = np.random.default_rng(1).permutation(u.atoms.n_atoms)
    e1 = mb.DielectricPlanar(u.atoms).run()
    e2 = mb.DielectricPlanar(u.atoms[permute]).run()
    assert e1._cuts.contiguous and not e2._cuts.contiguous
    close(e2.results.eps_par, e1.results.eps_par)
    close(e2.results.eps_perp, e1.results.eps_perp)


def test_dielectric_cylinder_and_sphere_against_restatement(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    u = mb.synthetic.spce_universe(700, 20261025, n_frames=3)
    q = np.asarray(u.atoms.charges, dtype=np.float64)
    inv = np.unique(u.atoms.molnums, return_inverse=True)[1]
    frames = _transformed_frames(u, "molecules")
    cyl = mb.DielectricCylinder(u.atoms, bin_width=0.25, vcutwidth=0.1).run()
    st = restatement.RunningStats()
    for p in frames:
        obs, _ = restatement.dielectric_cylinder_frame(p, q, inv, p, q, u.dimensions, 2, cyl.n_bins, 0.1)
        st.update({k: obs[k] for k in ("m_r", "m_r_tot", "M_r", "mM_r", "m_z", "M_z", "mM_z")})
    for k in ("m_r", "m_r_tot", "M_r", "mM_r", "m_z", "M_z", "mM_z"):
        close(cyl.means[k], st.means[k])
    pref = cyl._pref
    close(cyl.results.eps_z, pref * (st.means["mM_z"] - st.means["m_z"] * st.means["M_z"]), scale=pref * np.abs(st.means["mM_z"]).max())
    cyl.save()
    assert_allclose(np.loadtxt("eps_cyl_z.dat")[:, 1], cyl.results.eps_z)
    assert_allclose(np.loadtxt("eps_cyl_r.dat")[:, 1], cyl.results.eps_r)

    sph = mb.DielectricSphere(u.atoms, bin_width=0.25).run()
    st = restatement.RunningStats()
    for p in frames:
        obs, _ = restatement.dielectric_sphere_frame(p, q, p, q, u.dimensions, sph.n_bins)
        st.update({k: obs[k] for k in ("m_r", "m_r_tot", "M_r", "mM_r")})
    for k in ("m_r", "m_r_tot", "M_r", "mM_r"):
        close(sph.means[k], st.means[k])
    cov = st.means["mM_r"] - st.means["m_r"] * st.means["M_r"]
    close(sph.results.eps_rad, 1 - 4 * np.pi * sph.results.bin_pos**2 * sph._pref * cov,
          scale=np.max(4 * np.pi * sph.results.bin_pos**2 * sph._pref * np.abs(st.means["mM_r"])))
    sph.save()
    assert_allclose(np.loadtxt("eps_sph_rad.dat")[:, 1], sph.results.eps_rad)


def test_charge_neutrality_is_enforced():
    """lib/util.py:470-517: free charges in the group are an error (filter="error"), a charged universe a ValueError."""
    u = mb.synthetic.spce_universe(50, 1, n_frames=1)
    with pytest.raises(UserWarning, match="free charges"):
        mb.DielectricPlanar(u.select_atoms("element O")).run()
    q = u._attrs["charges"].copy()
    q[0] += 0.5  # an ion outside the (neutral) selection
    u._attrs["charges"] = q
    with pytest.raises(ValueError, match="non-neutral"):
        mb.DielectricPlanar(u.atoms[3:]).run()
