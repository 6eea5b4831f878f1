"""The reference's own module-level regression tests, run against the CUDA classes on the in-memory shim."""
import numpy as np
import pytest
from numpy.testing import assert_allclose
from conftest import water_universe
from reference_values import DENSITY_CYLINDER_MEAN, DENSITY_PLANAR_MEAN, DIPORDER_PLANAR, circle_of_water

import maicos_b200 as mb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order_parameter", ["P0", "cos_theta", "cos_2_theta"])
@pytest.mark.parametrize("dim", [0, 1, 2])
def test_diporder_planar_regression(water_gro, dim, order_parameter):
    """tests/modules/test_diporderplanar.py:65-82."""
    u = water_universe(water_gro["positions"], water_gro["dimensions"], water_gro["elements"])
    dip = mb.DiporderPlanar(u.atoms, bin_width=5, dim=dim, pdim=dim, refgroup=u.atoms,
                            order_parameter=order_parameter).run()
    assert_allclose(dip.results.profile.flatten(), DIPORDER_PLANAR[dim][order_parameter], rtol=1e-2)


@pytest.mark.parametrize("dens", ["mass", "number", "charge"])
def test_density_cylinder_single_frame(water_gro, dens):
    """tests/modules/test_densitycylinder.py:38-48."""
    u = water_universe(water_gro["positions"], water_gro["dimensions"], water_gro["elements"])
    d = mb.DensityCylinder(u.atoms, dens=dens).run()
    assert_allclose(d.results.profile.mean(), DENSITY_CYLINDER_MEAN[dens], atol=1e-4, rtol=1e-2)


def test_density_cylinder_regularly_spaced_molecules():
    """tests/modules/test_densitycylinder.py:50-70: ten molecules on a circle of radius 5, bin width 2."""
    ag, volume = circle_of_water(n_molecules=10, radius=5, bin_width=2)
    d = mb.DensityCylinder(ag, bin_width=2, dens="number", refgroup=ag).run()
    assert_allclose(d.results.profile, np.array([0, 0, 30, 0, 0]) / volume)


@pytest.mark.parametrize("dens", ["mass", "number", "charge"])
@pytest.mark.parametrize("dim", [0, 1, 2])
def test_density_planar_means(water_frames, water_gro, dens, dim):
    """tests/modules/test_densityplanar.py:99-107 (four frames of the same NPT trajectory)."""
    u = water_universe(water_frames["positions"], water_frames["dimensions"], water_gro["elements"])
    d = mb.DensityPlanar(u.atoms, dens=dens, dim=dim).run()
    assert_allclose(d.results.profile.mean(), DENSITY_PLANAR_MEAN[dens], rtol=1e-1, atol=1e-8)
    one = mb.DensityPlanar(u.atoms).run(stop=1)  # :109-115
    assert not np.isnan(one.results.profile).any()


def test_cli_end_to_end(tmp_path, monkeypatch, water_gro):
    """`maicos densityplanar -s ... -atomgroup ... -bin_width ...` (tests/test_main.py:24-42) through the shim readers."""
    from maicos_b200.__main__ import main

    monkeypatch.chdir(tmp_path)
    lines = ["water", str(len(water_gro["positions"]))]
    for i, (p, nm) in enumerate(zip(water_gro["positions"], water_gro["names"])):
        lines.append(f"{i // 3 + 1:5d}{'SOL':<5s}{nm:>5s}{i + 1:5d}{p[0] / 10:8.3f}{p[1] / 10:8.3f}{p[2] / 10:8.3f}")
    b = water_gro["dimensions"][:3] / 10
    lines.append(f"{b[0]:10.5f}{b[1]:10.5f}{b[2]:10.5f}")
    (tmp_path / "w.gro").write_text("\n".join(lines) + "\n")
    assert main(["densityplanar", "-s", "w.gro", "-atomgroup", "name OW", "-bin_width", "0.5", "-dens", "number",
                 "-output", "dens.dat"]) == 0
    out = np.loadtxt("dens.dat")
    assert out.shape[1] == 3 and abs(out[:, 1].mean() - 510 / np.prod(water_gro["dimensions"][:3])) < 2e-3
    assert "densityplanar" in (tmp_path / "dens.dat").read_text().lower()
