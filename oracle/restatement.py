"""ORACLE (test infrastructure only): NumPy restatement of the reference's frame bodies.

Nothing under ``maicos_b200/`` imports this file.  Every function names the reference lines
(MAICoS v0.11.2, paths relative to ``/root/reference/src/maicos``) it restates; the structure-factor
loop itself is *not* restated here -- it is ``oracle.kernel.structure_factor`` (the reference's own
Cython kernel compiled into ``oracle/_ref`` or the C port ``oracle/sfactor_port.c``).

Parity status: PINNED by ``tests/test_oracle.py`` -- six-vector theta table
(``tests/lib/test_math.py:183-209``), ``Saxs(water.gro, qmax=20).scattering_intensities[0] ==
1.6047`` (``tests/modules/test_saxs.py:57-71``), hydrogen form-factor table
(``tests/lib/test_math.py:605-640``), ``_compute_histogram`` known answers
(``tests/core/test_planar.py:509-528``, ``test_cylinder.py:529-547``, ``test_sphere.py:531-549``),
running statistics == ``np.std/sqrt(n)`` (``tests/core/test_base.py:222-229``).
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from . import kernel

_REF_TABLE = Path("/root/reference/src/maicos/share/scatteringfactors.dat")
_JSON_TABLE = Path(__file__).resolve().parents[1] / "maicos_b200" / "share" / "cromer_mann.json"


def cromer_mann_table() -> dict:
    """lib/tables.py:31-43 -- element -> (a[4], b[4], c); read from the reference's data file when it
    is present, else from the JSON copy generated from it (``tools/make_cm_table.py``)."""
    table = {}
    if _REF_TABLE.exists():
        for line in _REF_TABLE.read_text().splitlines():
            if line and line[0] != "#":
                p = line.split()
                table[p[0]] = (np.array(p[1:5], dtype=np.double), np.array(p[5:9], dtype=np.double), float(p[9]))
    else:
        for k, v in json.loads(_JSON_TABLE.read_text())["elements"].items():
            table[k] = (np.array(v["a"]), np.array(v["b"]), float(v["c"]))
    return table


_CM = None


def atomic_form_factor(q, element: str):
    """lib/math.py:623-702."""
    global _CM
    if _CM is None:
        _CM = cromer_mann_table()
    united = {"CH1": ("C", 1), "CH2": ("C", 2), "CH3": ("C", 3), "CH4": ("C", 4), "NH1": ("N", 1),
              "NH2": ("N", 2), "NH3": ("N", 3)}
    if element in united:  # math.py:671-684
        heavy, n_h = united[element]
        h = atomic_form_factor(q, "H")
        return atomic_form_factor(q, heavy) + (h if n_h == 1 else n_h * h)
    if element.title() not in _CM:  # math.py:686-690
        raise ValueError(
            f"Element '{element}' not found. Known elements are listed in the "
            "`maicos.lib.tables.elements` set."
        )
    a, b, c = _CM[element.title()]
    q2 = np.asarray((q / (4 * np.pi)) ** 2)  # math.py:693
    flat = q2.flatten()
    ff = np.sum(a * np.exp(-b * flat[:, None]), axis=1) + c  # math.py:696-699
    return ff.reshape(q2.shape)


def wrap_atoms_f32(positions: np.ndarray, box: np.ndarray) -> np.ndarray:
    """``AtomGroup.wrap(compound="atoms")`` for an orthorhombic cell as the in-memory shim does it
    (``x - floor(x / L) * L`` in float32); identity for atoms already inside ``[0, L)``.
    MDAnalysis itself is third-party and absent: parity of this step is unpinned."""
    positions = np.asarray(positions, dtype=np.float32)
    box = np.asarray(box, dtype=np.float32)
    return positions - np.floor(positions / box) * box


def saxs_single_frame(positions, box, elements, qmin, qmax, thetamin, thetamax, n_bins, sfactor=None):
    """modules/saxs.py:175-241 (``bin_spectrum=True``).

    positions: float32 [n, 3] as MDAnalysis stores them, box: float32 [3], elements: per-atom symbols,
    angles in radians.  Returns ``(obs dict, correlation scalar)``."""
    sfactor = sfactor or kernel.structure_factor
    positions = np.asarray(positions, dtype=np.float32)
    box = np.asarray(box, dtype=np.float32)
    obs = {"structure_factors": np.zeros(n_bins), "scattering_intensities": np.zeros(n_bins)}  # :178-180
    last = None
    for element in np.unique(elements):  # saxs.py:158-164 groups, :191 loop
        p = positions[np.asarray(elements) == element]
        p = p - box * np.round(p / box)  # :193-195 (float32 arithmetic)
        q, s = sfactor(np.double(p), np.double(box), qmin, qmax, thetamin, thetamax, np.ones(len(p)))  # :197-205
        i = atomic_form_factor(q, str(element)) ** 2 * s  # :207-210
        q, s, i = q.flatten(), s.flatten(), i.flatten()  # :213-215
        nz = np.where(s != 0)[0]  # :216
        q, s, i = q[nz], s[nz], i[nz]
        kw = dict(a=q, bins=n_bins, range=(qmin, qmax))  # :222-226
        s_b, _ = np.histogram(weights=s, **kw)
        i_b, _ = np.histogram(weights=i, **kw)
        obs["bincount"], _ = np.histogram(weights=None, **kw)  # :233 (overwritten per group)
        obs["structure_factors"] += s_b  # :234
        obs["scattering_intensities"] += i_b  # :235
        last = s_b
    return obs, last.flatten()[-1]  # :241


def saxs_conclude(sums, sems, n_atoms, qmin, qmax, dq):
    """modules/saxs.py:243-300 (``bin_spectrum=True``)."""
    q = np.arange(qmin, qmax, dq) + 0.5 * dq  # :245-247
    with np.errstate(divide="ignore", invalid="ignore"):
        s = sums["structure_factors"] / sums["bincount"]  # :248
        i = sums["scattering_intensities"] / sums["bincount"]  # :249-251
    ds, di = sems["structure_factors"], sems["scattering_intensities"]  # :252-253
    keep = np.invert(np.isnan(s))  # :277
    return {
        "scattering_vectors": q[keep],
        "structure_factors": s[keep] / n_atoms,  # :286-290
        "scattering_intensities": i[keep] / n_atoms,
        "dstructure_factors": ds[keep] / n_atoms,
        "dscattering_intensities": di[keep] / n_atoms,
    }


def dipsf_single_frame(positions, weights3, box, qmin, qmax, n_bins, sfactor=None):
    """modules/diporderstructurefactor.py:93-155.  positions float64 [n,3]; weights3 float64 [3,n]."""
    sfactor = sfactor or kernel.structure_factor
    total = np.zeros(n_bins)  # :102
    for pdim in range(3):  # :105
        q, s = sfactor(np.double(positions), np.double(box), qmin, qmax, 0, np.pi,
                       np.ascontiguousarray(weights3[pdim]))  # :121-129
        q, s = q.flatten(), s.flatten()
        nz = np.where(s != 0)[0]  # :133
        q, s = q[nz], s[nz]
        kw = dict(a=q, bins=n_bins, range=(qmin, qmax))
        s_b, _ = np.histogram(weights=s, **kw)  # :141-143
        c_b, _ = np.histogram(weights=None, **kw)  # :144
        with np.errstate(invalid="ignore"):
            s_b /= c_b  # :146-147
        total += np.nan_to_num(s_b)  # :150
    total /= len(positions)  # :153
    return {"structure_factors": total}, total[-1]  # :155


def transform_cylinder(positions, origin, dim):
    """lib/math.py:705-744."""
    out = np.zeros(positions.shape)
    odims = np.roll(np.arange(3), -dim)[1:]
    c = positions - origin
    out[:, 0] = np.linalg.norm(c[:, odims], axis=1)
    np.arctan2(*c[:, odims].T, out=out[:, 1])
    out[:, 2] = np.copy(positions[:, dim])
    return out


def transform_sphere(positions, origin):
    """lib/math.py:747-778."""
    out = np.zeros(positions.shape)
    c = positions - origin
    out[:, 0] = np.linalg.norm(c, axis=1)
    np.arctan2(c[:, 1], c[:, 0], out=out[:, 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        np.arccos(c[:, 2] / out[:, 0], out=out[:, 2])
    return out


def hist_planar(positions, weights, dim, n_bins, zmin, zmax):
    """core/planar.py:235-243."""
    hist, _ = np.histogram(positions[:, dim], bins=n_bins, range=(np.float64(zmin), np.float64(zmax)), weights=weights)
    return hist


def hist_cylinder(positions, weights, dim, n_bins, rmin, rmax, zmin, zmax, box_center):
    """core/cylinder.py:229-243."""
    p = transform_cylinder(positions, box_center, dim)
    hist, _, _ = np.histogram2d(p[:, 0], p[:, 2], bins=(n_bins, 1),
                                range=((np.float64(rmin), np.float64(rmax)), (np.float64(zmin), np.float64(zmax))),
                                weights=weights)
    return hist[:, 0]


def hist_sphere(positions, weights, n_bins, rmin, rmax, box_center):
    """core/sphere.py:215-223."""
    r = transform_sphere(positions, origin=box_center)[:, 0]
    hist, _ = np.histogram(r, bins=n_bins, range=(np.float64(rmin), np.float64(rmax)), weights=weights)
    return hist


def planar_geometry(box_lengths, dim, n_bins, zmin_user=None, zmax_user=None):
    """core/planar.py:97-110 and 133-149 -> observables of one frame."""
    box_lengths = np.asarray(box_lengths, dtype=np.float64)
    center = box_lengths / 2
    zmin = np.float64(0 if zmin_user is None else center[dim] + zmin_user)
    zmax = np.float64(box_lengths[dim] if zmax_user is None else center[dim] + zmax_user)
    odims = np.roll(np.arange(3), -dim)[1:]
    obs = {"L": zmax - zmin, "box_center": center, "bin_edges": np.linspace(zmin, zmax, n_bins + 1)}
    obs["bin_width"] = obs["L"] / n_bins
    obs["bin_pos"] = obs["bin_edges"][1:] - obs["bin_width"] / 2
    obs["bin_area"] = np.ones(n_bins) * np.prod(box_lengths[odims])
    obs["bin_volume"] = obs["bin_area"] * obs["bin_width"]
    return zmin, zmax, obs


def profile_single_frame(hist_fn, positions, weights, normalization, bin_volume):
    """core/base.py:804-825: weighted pass, count pass, optional per-frame volume normalisation."""
    profile = hist_fn(positions, weights)  # :816
    bincount = hist_fn(positions, None)  # :818
    if normalization == "volume":
        profile = profile / bin_volume  # :820-821
    return {"profile": profile, "bincount": bincount}


class RunningStats:
    """core/base.py:453-494 with lib/math.py:325-459: per-key mean, sem (= sqrt(var_pop / n)), sum."""

    def __init__(self):
        self.n = 0
        self.means, self.sems, self.sums = {}, {}, {}

    def update(self, obs: dict):
        self.n += 1
        n = self.n
        if n == 1:  # base.py:482-494
            self.sums = {k: np.copy(v) for k, v in obs.items()}
            self.means = {k: np.copy(v) for k, v in obs.items()}
            self.sems = {k: np.zeros(np.shape(v)) for k, v in obs.items()}
            return
        for k, x in obs.items():  # base.py:460-480
            old_mean = self.means[k]
            old_var = self.sems[k] ** 2 * (n - 1)
            self.means[k] = ((n - 1) * old_mean + x) / n  # math.py:379
            s_new = old_var * (n - 1) + (x - old_mean) * (x - self.means[k])  # math.py:450-451
            s_new = np.where(s_new < 0, 0, s_new)  # math.py:453-457
            self.sems[k] = np.sqrt(s_new / n / n)
            self.sums[k] = self.sums[k] + x


def run_saxs(frames, boxes, elements, qmin=0, qmax=6, dq=0.1, thetamin=0, thetamax=180, pack=True, sfactor=None):
    """``Saxs(ag, ...).run()`` restated end to end for in-memory frames (saxs.py:133-300 on top of
    core/base.py:417-499).  ``frames`` float32 [F, n, 3], ``boxes`` float32 [F, 3]."""
    thetamin, thetamax = min(thetamin, thetamax) * np.pi / 180, max(thetamin, thetamax) * np.pi / 180  # :134-152
    n_bins = int(np.ceil((qmax - qmin) / dq))  # :167
    stats = RunningStats()
    series = []
    for pos, box in zip(frames, boxes):
        pos = np.asarray(pos, dtype=np.float32)
        if pack:  # base.py:446-448
            pos = wrap_atoms_f32(pos, box)
        obs, val = saxs_single_frame(pos, box, elements, qmin, qmax, thetamin, thetamax, n_bins, sfactor)
        stats.update(obs)
        series.append(val)
    res = saxs_conclude(stats.sums, stats.sems, len(elements), qmin, qmax, dq)
    return res, stats, np.array(series)
