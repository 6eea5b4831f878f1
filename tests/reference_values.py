"""Regression values pinned by the reference's own module tests (produced there by MAICoS + MDAnalysis)."""
import numpy as np

#: tests/modules/test_diporderplanar.py:41-82 -- DiporderPlanar(water.gro, bin_width=5, dim=d, pdim=d, refgroup=ag,
#: order_parameter=op).results.profile, rtol 1e-2
DIPORDER_PLANAR = {
    0: {"P0": [6.33e-05, -9.26e-04, -8.31e-04, -4.92e-04, -9.56e-04],
        "cos_theta": [4.25e-3, -5.49e-2, -4.83e-2, -3.1188e-2, -5.8e-2],
        "cos_2_theta": [3.90e-1, 3.41e-1, 4.10e-1, 3.16e-1, 2.94e-1]},
    1: {"P0": [0.000344, 0.000734, 0.00092, 0.000335, 0.000667],
        "cos_theta": [0.022442, 0.043359, 0.056447, 0.021081, 0.040248],
        "cos_2_theta": [4.05e-1, 3.56e-1, 2.71e-1, 3.57e-1, 3.31e-1]},
    2: {"P0": [-1.12e-3, -1.61e-3, -1.22e-3, -1.44e-3, -1.29e-3],
        "cos_theta": [-6.96e-2, -9.65e-2, -7.41e-2, -8.76e-2, -8.26e-2],
        "cos_2_theta": [2.72e-1, 3.45e-1, 3.41e-1, 3.12e-1, 2.62e-1]},
}
#: tests/modules/test_densitycylinder.py:38-48 -- DensityCylinder(water.gro, dens=...).results.profile.mean(),
#: atol 1e-4, rtol 1e-2
DENSITY_CYLINDER_MEAN = {"mass": 0.561, "number": 0.095, "charge": 0.000609}
#: tests/modules/test_densityplanar.py:99-107 -- DensityPlanar(water.trr, dens=..., dim=0..2).results.profile.mean(),
#: rtol 1e-1, atol 1e-8
DENSITY_PLANAR_MEAN = {"mass": 0.588, "number": 0.097, "charge": 0.0}


def circle_of_water(n_molecules=10, radius=5.0, bin_width=2.0):
    """tests/util.py:122-182 on the in-memory shim: SPC/E molecules on a circle about the centre of a 20 A box."""
    from maicos_b200.mdshim import MemoryTrajectory, Universe
    from maicos_b200.synthetic import M_H, M_O, Q_H, Q_O

    mol = np.array([[0.0, 0.0, 0.0], [0.8164904, 0.5773590, 0.0], [-0.8164904, 0.5773590, 0.0]]) * 0.3
    pos = []
    for i in range(n_molecules):
        c = 10 + np.array([np.cos(2 * np.pi / n_molecules * i) * radius, np.sin(2 * np.pi / n_molecules * i) * radius, 0])
        pos.append(c + mol)
    pos = np.concatenate(pos).astype(np.float32)
    n = 3 * n_molecules
    res = np.repeat(np.arange(n_molecules), 3)
    names = np.tile(np.array(["OW", "HW1", "HW2"], dtype=object), n_molecules)
    u = Universe(n, MemoryTrajectory(pos[None], np.float32([20, 20, 20, 90, 90, 90])), names=names, types=names,
                 elements=np.tile(np.array(["O", "H", "H"], dtype=object), n_molecules),
                 masses=np.tile([M_O, M_H, M_H], n_molecules), charges=np.tile([Q_O, Q_H, Q_H], n_molecules),
                 resindices=res, molnums=res)
    n_bins = np.int32(10 / bin_width)
    edges = np.linspace(0, 10, n_bins + 1)
    return u.atoms, np.pi * np.diff(edges**2) * 20
