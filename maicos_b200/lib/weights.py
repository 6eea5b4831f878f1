"""Weight providers feeding the two hot paths.

Same call signatures and results as ``maicos.lib.weights``
(``/root/reference/src/maicos/lib/weights.py``: ``density_weights`` 43-96,
``temperature_weights`` 100-127, ``diporder_weights`` 131-178, ``velocity_weights`` 195-225).
They are O(N) per frame and run on the host from AtomGroup attributes.
"""

from __future__ import annotations

import numpy as np

from . import tables

_GROUPINGS = ["atoms", "residues", "segments", "molecules", "fragments"]


def _electrons(element: str) -> float:
    if element in tables.UNITED_ATOMS:
        heavy, n_h = tables.UNITED_ATOMS[element]
        return _electrons(heavy) + n_h * _electrons("H")
    try:
        return tables.electron_count[element.title()]
    except KeyError as e:
        raise KeyError(
            f"Element '{e.args[0]}' not found. Known elements are listed in the "
            "`maicos.lib.tables.elements` set."
        ) from e


def density_weights(atomgroup, grouping: str, dens: str) -> np.ndarray:
    """Per-compound weights of a density profile: ``number``, ``mass``, ``charge`` or ``electron``."""
    if grouping not in _GROUPINGS:
        raise ValueError(
            f"'{grouping}' grouping is not supported. "
            "Use `atoms`, `residues`, `segments`, `molecules` or `fragments`."
        )
    if dens == "number":
        n = len(np.unique(atomgroup.molnums)) if grouping == "molecules" else getattr(atomgroup, f"n_{grouping}")
        return np.ones(n)
    if dens == "mass":
        return atomgroup.masses if grouping == "atoms" else atomgroup.total_mass(compound=grouping)
    if dens == "charge":
        return atomgroup.charges if grouping == "atoms" else atomgroup.total_charge(compound=grouping)
    if dens == "electron":
        w = np.array([_electrons(el) for el in atomgroup.elements])
        return w if grouping == "atoms" else atomgroup.accumulate(w, compound=grouping)
    raise ValueError(f"'{dens}' density type is not supported. Use 'mass', 'number', 'charge' or 'electron'.")


def temperature_weights(atomgroup, grouping: str) -> np.ndarray:
    """Kinetic temperature contribution per atom (K), ``m v^2 / 2 k_B`` in MD units."""
    if grouping != "atoms":
        raise NotImplementedError(
            f"Temperature calculations of '{grouping}' is not supported. Use 'atoms' instead.'"
        )
    atomic_mass, boltzmann = 1.66053906660e-27, 1.380649e-23  # CODATA 2018 (scipy.constants)
    prefac = atomic_mass * 1e4 / boltzmann
    return (atomgroup.velocities**2).sum(axis=1) * atomgroup.atoms.masses / 2 * prefac


def diporder_weights(atomgroup, grouping: str, order_parameter: str, get_unit_vectors) -> np.ndarray:
    """Dipole projections per compound: ``P0`` = mu.e, ``cos_theta`` = mu.e/|mu|, ``cos_2_theta`` its square."""
    dipoles = atomgroup.dipole_vector(compound=grouping)
    unit_vectors = get_unit_vectors(atomgroup=atomgroup, grouping=grouping)
    if unit_vectors.shape != (3,) and unit_vectors.shape != (len(dipoles), 3):
        raise ValueError(
            f"Returned unit vectors have shape {unit_vectors.shape}. But only shape "
            f"(3,) or {(len(dipoles), 3)} is allowed."
        )
    if order_parameter == "P0":
        return np.sum(dipoles * unit_vectors, axis=1)
    if order_parameter in ("cos_theta", "cos_2_theta"):
        w = np.sum(dipoles / np.linalg.norm(dipoles, axis=1)[:, np.newaxis] * unit_vectors, axis=1)
        if order_parameter == "cos_2_theta":
            w *= w
        return w
    raise ValueError(f"'{order_parameter}' not supported. Use 'P0', 'cos_theta' or 'cos_2_theta'.")


def velocity_weights(atomgroup, grouping: str, vdim: int) -> np.ndarray:
    """Velocity component ``vdim`` per atom, or mass-weighted per compound."""
    v = atomgroup.velocities[:, vdim]
    if grouping == "atoms":
        return v
    m = atomgroup.atoms.masses
    return atomgroup.atoms.accumulate(v * m, compound=grouping) / atomgroup.atoms.accumulate(m, compound=grouping)
