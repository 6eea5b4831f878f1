"""Static lookup tables (Cromer-Mann X-ray form-factor coefficients, electron counts).

Mirror of ``maicos.lib.tables`` (``/root/reference/src/maicos/lib/tables.py:31-54``): the
same three public names ``CM_parameters`` (element -> ``a[4], b[4], c``), ``elements`` and
``electron_count`` (``sum(a) + c``, the q -> 0 limit of the form factor).  The coefficients
are loaded from ``share/cromer_mann.json`` (regenerated from the reference's data file by
``tools/make_cm_table.py``).
"""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class CMParameter:
    """Cromer-Mann parameters of one element: ``f(q) = sum_i a_i exp(-b_i (q/4pi)^2) + c``."""

    a: np.ndarray
    b: np.ndarray
    c: float


def _load() -> dict[str, CMParameter]:
    raw = json.loads((Path(__file__).resolve().parents[1] / "share" / "cromer_mann.json").read_text())
    return {
        name: CMParameter(a=np.array(v["a"], dtype=np.double), b=np.array(v["b"], dtype=np.double), c=float(v["c"]))
        for name, v in raw["elements"].items()
    }


#: Cromer-Mann X-ray scattering factors, keyed by element symbol (title case, ions as "Na1+").
CM_parameters: dict[str, CMParameter] = _load()

#: Set of elements with known coefficients.
elements = set(CM_parameters)

#: Number of electrons per element, from the q = 0 limit of the Cromer-Mann sum.
electron_count = {name: np.sum(p.a) + p.c for name, p in CM_parameters.items()}

#: United-atom aliases: heavy element + n hydrogens (``lib/math.py:671-684``).
UNITED_ATOMS = {
    "CH1": ("C", 1), "CH2": ("C", 2), "CH3": ("C", 3), "CH4": ("C", 4),
    "NH1": ("N", 1), "NH2": ("N", 2), "NH3": ("N", 3),
}


def cm_parameter_sets(element: str) -> list[tuple[float, CMParameter]]:
    """``[(multiplicity, parameters), ...]`` whose weighted form factors sum to ``f_element``.

    Raises the reference's ``ValueError`` text for unknown elements
    (``lib/math.py:686-690``, pinned by ``tests/modules/test_saxs.py:78-88``).
    """
    if element in UNITED_ATOMS:
        heavy, n_h = UNITED_ATOMS[element]
        return [(1.0, CM_parameters[heavy]), (float(n_h), CM_parameters["H"])]
    if element.title() not in CM_parameters:
        raise ValueError(
            f"Element '{element}' not found. Known elements are listed in the "
            "`maicos.lib.tables.elements` set."
        )
    return [(1.0, CM_parameters[element.title()])]


def cm_packed(element: str) -> np.ndarray:
    """``float64[2][10]`` block of ``mb200_saxs_frames``' ``cm_params`` argument for one group."""
    out = np.zeros((2, 10), dtype=np.float64)
    for i, (mult, p) in enumerate(cm_parameter_sets(element)):
        out[i, 0] = mult
        out[i, 1:5] = p.a
        out[i, 5:9] = p.b
        out[i, 9] = p.c
    return out
