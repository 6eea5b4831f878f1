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


def saxs_single_frame_unbinned(positions, box, elements, qmin, qmax, thetamin, thetamax, max_n, sfactor=None):
    """modules/saxs.py:175-241 with ``bin_spectrum=False``: the raw cubes of all element groups summed (:188-189,
    :238-239).  ``max_n`` is the class's float64 cube shape (:173); a kernel cube of another shape fails in ``+=`` as it
    does in the reference.  Returns ``(obs dict, correlation scalar)``."""
    sfactor = sfactor or kernel.structure_factor
    positions = np.asarray(positions, dtype=np.float32)
    box = np.asarray(box, dtype=np.float32)
    obs = {"structure_factors": np.zeros(max_n), "scattering_intensities": np.zeros(max_n)}  # :188-189
    s = None
    for element in np.unique(elements):
        p = positions[np.asarray(elements) == element]
        p = p - box * np.round(p / box)  # :193-195
        q, s = sfactor(np.double(p), np.double(box), qmin, qmax, thetamin, thetamax, np.ones(len(p)))  # :197-205
        obs["structure_factors"] += s  # :238
        obs["scattering_intensities"] += atomic_form_factor(q, str(element)) ** 2 * s  # :207-210, :239
    return obs, s.flatten()[-1]  # :241


def saxs_conclude_unbinned(means, n_atoms, box, max_n):
    """modules/saxs.py:255-300 with ``bin_spectrum=False``: Miller indices sorted by |q|."""
    factors = 2 * np.pi / np.asarray(box)  # :172
    miller = np.array(list(np.ndindex(tuple(max_n))))  # :256
    q = np.linalg.norm(miller * factors[np.newaxis, :], axis=1)  # :257-259
    order = np.argsort(q)  # :260
    q, miller = q[order], miller[order]
    s = means["structure_factors"].flatten()[order]  # :264-268
    i = means["scattering_intensities"].flatten()[order]
    keep = np.invert(np.isnan(s))  # :277
    return {"scattering_vectors": q[keep], "miller_indices": miller[keep], "structure_factors": s[keep] / n_atoms,
            "scattering_intensities": i[keep] / n_atoms}


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


def box_diagonal(dimensions):
    """``np.diag(mda.lib.mdamath.triclinic_vectors(ts.dimensions))`` (saxs.py:176, diporderstructurefactor.py:94).

    MDAnalysis (>= 2.8.0, pyproject.toml:63) is a third-party dependency that is not in /root/reference; its published
    construction puts a along x, b in the xy plane and c wherever the three angles require, in float64, and returns
    float32.  The diagonal in closed form: (lx, ly sin(gamma), lz sqrt(1 - cos(beta)^2 - cy^2)) with
    cy = (cos(alpha) - cos(beta) cos(gamma)) / sin(gamma); right angles use exact 0 / 1; an invalid cell gives zeros.
    """
    d = np.asarray(dimensions, dtype=np.float64)
    if d.shape[-1] == 3:
        return d.astype(np.float32)
    lx, ly, lz, al, be, ga = d[:6]
    if not (np.all(d[:6] > 0) and max(al, be, ga) < 180.0):
        return np.zeros(3, dtype=np.float32)
    if al == 90.0 and be == 90.0 and ga == 90.0:
        return d[:3].astype(np.float32)
    ca = 0.0 if al == 90.0 else np.cos(np.deg2rad(al))
    cb = 0.0 if be == 90.0 else np.cos(np.deg2rad(be))
    cg, sg = (0.0, 1.0) if ga == 90.0 else (np.cos(np.deg2rad(ga)), np.sin(np.deg2rad(ga)))
    cx, cy = lz * cb, lz * (ca - cb * cg) / sg
    with np.errstate(invalid="ignore"):
        czz = np.sqrt(lz * lz - cx ** 2 - cy ** 2)
    if not czz > 0.0:
        return np.zeros(3, dtype=np.float32)
    return np.array([lx, ly * sg, czz]).astype(np.float32)


def run_saxs(frames, boxes, elements, qmin=0, qmax=6, dq=0.1, thetamin=0, thetamax=180, pack=True, sfactor=None):
    """``Saxs(ag, ...).run()`` restated end to end for in-memory frames (saxs.py:133-300 on top of
    core/base.py:417-499).  ``frames`` float32 [F, n, 3], ``boxes`` float32 [F, 3]."""
    # saxs.py:136-137 assigns sequentially, so the second line sees the already replaced thetamin: swapped arguments
    # (150, 30) become (30, 30), not (30, 150).  Kept as the reference has it.
    thetamin = min(thetamin, thetamax)  # :136
    thetamax = max(thetamin, thetamax)  # :137
    for name, value in (("thetamin", thetamin), ("thetamax", thetamax)):  # :139-143
        if value < 0 or value > 180:
            raise ValueError(f"{name} ({value}°) has to between 0 and 180°.")
    if thetamin > thetamax:  # :145-148 (unreachable after the two lines above; kept)
        raise ValueError(f"thetamin ({thetamin}°) larger than thetamax ({thetamax}°).")
    thetamin, thetamax = thetamin * np.pi / 180, thetamax * np.pi / 180  # :151-152
    n_bins = int(np.ceil((qmax - qmin) / dq))  # :167
    stats = RunningStats()
    series = []
    for pos, box in zip(frames, boxes):
        pos = np.asarray(pos, dtype=np.float32)
        box = box_diagonal(box)  # saxs.py:176 (``boxes`` rows may be [lx, ly, lz] or full ``ts.dimensions``)
        if pack:  # base.py:446-448
            pos = wrap_atoms_f32(pos, box)
        obs, val = saxs_single_frame(pos, box, elements, qmin, qmax, thetamin, thetamax, n_bins, sfactor)
        stats.update(obs)
        series.append(val)
    res = saxs_conclude(stats.sums, stats.sems, len(elements), qmin, qmax, dq)
    return res, stats, np.array(series)


def compound_prologue(positions, box, compound_ptr, masses, charges, bin_method, unwrap, pack, weight_mode=0, pdim=0):
    """What the reference does per frame before a compound-grouped hot kernel, for one frame (orthorhombic cell,
    compounds = contiguous atom runs ``compound_ptr[c] .. compound_ptr[c + 1] - 1``):

    * ``atoms.unwrap(compound)`` (core/base.py:438-440): MDAnalysis makes every compound whole and puts its centre of
      mass into the primary cell; ``atoms.wrap(compound)`` (base.py:446-448) shifts whole compounds by lattice vectors so
      that the centre of mass lies in the cell.  MDAnalysis is absent here (third party): its bond walk is restated as
      the minimum image about the compound's first atom, valid for compounds smaller than half the cell (water);
    * ``get_center(ag, bin_method, compound)`` (lib/util.py:642-680): weighted mean with masses / |charges| / ones;
    * ``diporder_weights`` (lib/weights.py:131-178) on ``ag.dipole_vector(compound)`` = sum q_i (r_i - R_com):
      ``weight_mode`` 1 ``mu[pdim]``, 2 ``mu[pdim] / |mu|``, 3 its square, 4 all three components of ``mu / |mu|``.

    Coordinates stay float32 through the wrapping steps like ``Timestep.positions``; sums are float64.
    Returns ``(centres [n_c, 3], weights [n_c] | [3, n_c] | None)``.
    """
    pos = np.array(positions, dtype=np.float32)
    box = np.asarray(box, dtype=np.float32)
    n_c = len(compound_ptr) - 1
    sizes = np.diff(compound_ptr)
    ids = np.repeat(np.arange(n_c), sizes)
    masses = np.asarray(masses, dtype=np.float64)
    charges = np.asarray(charges, dtype=np.float64)
    if unwrap:
        anchor = pos[np.asarray(compound_ptr[:-1])][ids]
        delta = pos - anchor
        delta = delta - box * np.round(delta / box)
        pos = (anchor + delta).astype(np.float32)
    if unwrap or pack:
        msum = np.bincount(ids, weights=masses, minlength=n_c)
        com = np.stack([np.bincount(ids, weights=masses * pos[:, d], minlength=n_c) / msum for d in range(3)], 1)
        shift = (-np.floor(com / np.double(box)) * np.double(box)).astype(np.float32)
        pos = (pos + shift[ids]).astype(np.float32)
    cw = {"com": masses, "cog": np.ones(len(pos)), "coc": np.abs(charges)}[bin_method]  # util.py:668-679
    wsum = np.bincount(ids, weights=cw, minlength=n_c)
    centres = np.stack([np.bincount(ids, weights=cw * pos[:, d], minlength=n_c) / wsum for d in range(3)], 1)
    if not weight_mode:
        return centres, None
    msum = np.bincount(ids, weights=masses, minlength=n_c)
    com = np.stack([np.bincount(ids, weights=masses * pos[:, d], minlength=n_c) / msum for d in range(3)], 1)
    rel = np.double(pos) - com[ids]
    mu = np.stack([np.bincount(ids, weights=charges * rel[:, d], minlength=n_c) for d in range(3)], 1)
    if weight_mode == 1:  # weights.py:160-161
        return centres, mu[:, pdim]
    with np.errstate(invalid="ignore", divide="ignore"):
        unit = mu / np.linalg.norm(mu, axis=1)[:, None]  # weights.py:163-167
    if weight_mode == 4:
        return centres, np.ascontiguousarray(unit.T)
    w = unit[:, pdim]
    return centres, w * w if weight_mode == 3 else w


def unit_vectors_radial(centres, box, dim=None):
    """lib/util.py:710-756 (cylinder, ``pdim="r"``: ``dim`` = axis) and 760-786 (sphere, ``dim=None``): the direction
    from the box centre ``dimensions[:3] / 2`` (float32 halving) to every compound centre, axial component dropped for
    the cylinder, normalised."""
    rel = np.array(centres, dtype=np.float64)
    rel -= np.asarray(box, dtype=np.float32) / 2  # util.py:745 / 779
    if dim is not None:
        rel[:, dim] = 0  # util.py:750
    with np.errstate(invalid="ignore", divide="ignore"):
        rel /= np.linalg.norm(rel, axis=1)[:, np.newaxis]  # util.py:753 / 783
    return rel


def diporder_weights(dipoles, unit_vectors, order_parameter):
    """lib/weights.py:131-178 for per-compound unit vectors ``[n, 3]`` or one Cartesian vector ``[3]``."""
    if order_parameter == "P0":
        return np.sum(dipoles * unit_vectors, axis=1)  # :160-161
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.sum(dipoles / np.linalg.norm(dipoles, axis=1)[:, np.newaxis] * unit_vectors, axis=1)  # :163-167
    return w * w if order_parameter == "cos_2_theta" else w  # :168-169


def compound_dipoles(positions, compound_ptr, masses, charges):
    """``AtomGroup.dipole_vector(compound)``: sum_i q_i (r_i - R_com) per contiguous compound (whole molecules in)."""
    n_c = len(compound_ptr) - 1
    ids = np.repeat(np.arange(n_c), np.diff(compound_ptr))
    pos = np.double(positions)
    msum = np.bincount(ids, weights=masses, minlength=n_c)
    com = np.stack([np.bincount(ids, weights=masses * pos[:, d], minlength=n_c) / msum for d in range(3)], 1)
    rel = pos - com[ids]
    return np.stack([np.bincount(ids, weights=charges * rel[:, d], minlength=n_c) for d in range(3)], 1)


def velocity_weights(velocities, masses, compound_ptr, vdim):
    """lib/weights.py:195-225: ``v[:, vdim]`` per atom (``compound_ptr=None``) or its mass-weighted compound mean."""
    v = np.asarray(velocities, dtype=np.float32)[:, vdim]  # :213
    if compound_ptr is None:
        return v  # :215-216
    n_c = len(compound_ptr) - 1
    ids = np.repeat(np.arange(n_c), np.diff(compound_ptr))
    mass_vels = np.bincount(ids, weights=v * np.asarray(masses), minlength=n_c)  # :218-220
    group_mass = np.bincount(ids, weights=masses, minlength=n_c)  # :221-223
    return mass_vels / group_mass  # :224


def temperature_weights(velocities, masses):
    """lib/weights.py:101-127: ``(v**2).sum(axis=1) * m / 2 * prefac`` with float32 velocities."""
    atomic_mass, boltzmann = 1.66053906660e-27, 1.380649e-23  # scipy.constants (CODATA 2018)
    prefac = atomic_mass * 1e4 / boltzmann  # :126
    v = np.asarray(velocities, dtype=np.float32)
    return (v ** 2).sum(axis=1) * np.asarray(masses) / 2 * prefac  # :127


# ---------------------------------------------------------------------------------------------------------
# Dielectric profile modules: frame bodies with the reference's own digitize / bincount /
# cumsum / mean sequence, including the (virtual cuts x bins) table the CUDA path never materialises.
# Pinned by tests/test_oracle.py on the reference's known answers for synthetic dipoles
# (tests/modules/test_dielectricplanar.py:104-198: total dipole moment and volume integral of the dipole density).
# ---------------------------------------------------------------------------------------------------------
def _compound_centre(values, abs_charges, inverse_ix):
    """``atomgroup.center(weights=|q|, compound)`` / ``accumulate(v |q|) / accumulate(|q|)`` per compound."""
    n_c = int(inverse_ix.max()) + 1
    num = np.bincount(inverse_ix, weights=abs_charges * values, minlength=n_c)
    den = np.bincount(inverse_ix, weights=abs_charges, minlength=n_c)
    return num / den


def dielectric_planar_frame(pos_ag, charges_ag, inverse_ix, pos_all, charges_all, dimensions, dim, n_bins, vcutwidth,
                            zmin_user=None, zmax_user=None):
    """modules/dielectricplanar.py:152-241 on top of core/planar.py:97-110, 133-149.

    ``pos_*`` float32 ``[n, 3]`` as MDAnalysis stores them, ``dimensions`` float32 ``[lx, ly, lz, ...]``.
    """
    pos_ag = np.asarray(pos_ag, dtype=np.float32)
    pos_all = np.asarray(pos_all, dtype=np.float32)
    dimensions = np.asarray(dimensions, dtype=np.float32)
    box = dimensions[:3].astype(np.float64)
    odims = np.roll(np.arange(3), -dim)[1:]
    zmin = np.float64(0 if zmin_user is None else box[dim] / 2 + zmin_user)
    zmax = np.float64(box[dim] if zmax_user is None else box[dim] / 2 + zmax_user)
    o = {}
    o["L"] = zmax - zmin
    o["bin_edges"] = np.linspace(zmin, zmax, n_bins + 1)
    o["bin_width"] = o["L"] / n_bins
    o["bin_pos"] = o["bin_edges"][1:] - o["bin_width"] / 2
    o["bin_area"] = np.ones(n_bins) * np.prod(box[odims])
    o["bin_volume"] = o["bin_area"] * o["bin_width"]

    o["M"] = np.dot(charges_all, pos_all)                                        # 155-157
    o["M_perp"] = o["M"][dim]
    o["M_perp_2"] = o["M"][dim] ** 2
    o["M_par"] = o["M"][odims]
    zbins = np.digitize(pos_ag[:, dim], o["bin_edges"][1:-1])                   # 170-172
    curQ = np.bincount(zbins, weights=charges_ag, minlength=n_bins)              # 174-176
    o["m_perp"] = -np.cumsum(curQ / o["bin_area"])                               # 178
    o["mM_perp"] = o["m_perp"] * o["M_perp"]
    o["mm_perp"] = o["m_perp"] ** 2 * o["bin_volume"]
    o["cmM_perp"] = o["m_perp"] * (o["M_perp"] - o["m_perp"] * o["bin_volume"])
    o["cM_perp"] = o["M_perp"] - o["m_perp"] * o["bin_volume"]

    testpos = _compound_centre(pos_ag[:, dim].astype(np.float64), np.abs(charges_ag), inverse_ix)[inverse_ix]  # 192-194
    o["m_par"] = np.zeros((n_bins, 2))
    for j, direction in enumerate(odims):                                        # 197-219
        Lx = dimensions[direction]
        Ax = dimensions[odims[1 - j]] * o["bin_width"]
        vbinsx = np.ceil(Lx / vcutwidth).astype(int)
        x_bin_edges = (np.arange(vbinsx)) * (Lx / vbinsx)
        zpos = np.digitize(testpos, o["bin_edges"][1:-1])
        xbins = np.digitize(pos_ag[:, direction], x_bin_edges[1:])
        curQx = np.bincount(zpos + n_bins * xbins, weights=charges_ag, minlength=vbinsx * n_bins).reshape(vbinsx, n_bins)
        o["m_par"][:, j] = -np.cumsum(curQx / Ax, axis=0).mean(axis=0)
    bin_volume = o["bin_volume"][0]
    o["mM_par"] = np.dot(o["m_par"], o["M_par"])                                 # 226-233
    o["mm_par"] = (o["m_par"] * o["m_par"]).sum(axis=1) * bin_volume
    o["cmM_par"] = (o["m_par"] * (o["M_par"] - o["m_par"] * bin_volume)).sum(axis=1)
    o["cM_par"] = o["M_par"] - o["m_par"] * bin_volume
    return o, np.linalg.norm(o["M_par"])


def dielectric_cylinder_frame(pos_ag, charges_ag, inverse_ix, pos_all, charges_all, dimensions, dim, n_bins, vcutwidth,
                              rmin=0.0, rmax=None):
    """modules/dielectriccylinder.py:131-199 on top of core/cylinder.py:90-109, 132-145 (zmin / zmax = whole box)."""
    pos_ag = np.asarray(pos_ag, dtype=np.float32)
    pos_all = np.asarray(pos_all, dtype=np.float32)
    box = np.asarray(dimensions, dtype=np.float32)[:3].astype(np.float64)
    odims = np.roll(np.arange(3), -dim)[1:]
    center = box / 2
    rmin = np.float64(rmin)
    rmax = np.float64(box[odims].min() / 2 if rmax is None else rmax)
    o = {}
    o["L"] = np.float64(box[dim]) - np.float64(0)
    o["R"] = rmax - rmin
    o["bin_edges"] = np.linspace(rmin, rmax, n_bins + 1, endpoint=True)
    o["bin_width"] = o["R"] / n_bins
    o["bin_pos"] = o["bin_edges"][1:] - o["bin_width"] / 2
    o["bin_area"] = np.pi * np.diff(o["bin_edges"] ** 2)
    o["bin_volume"] = o["bin_area"] * o["L"]
    cyl_ag = transform_cylinder(pos_ag, center, dim)
    cyl_all = transform_cylinder(pos_all, center, dim)
    curQ_r = np.bincount(np.digitize(cyl_ag[:, 0], o["bin_edges"][1:-1]), weights=charges_ag, minlength=n_bins)   # 134-141
    o["m_r"] = -np.cumsum(curQ_r) / 2 / np.pi / o["L"] / o["bin_pos"]                                            # 148
    curQ_r_tot = np.bincount(np.digitize(cyl_all[:, 0], o["bin_edges"][1:-1]), weights=charges_all, minlength=n_bins)
    o["m_r_tot"] = -np.cumsum(curQ_r_tot) / 2 / np.pi / o["L"] / o["bin_pos"]                                    # 155-157
    o["M_r"] = np.sum(o["m_r_tot"] * o["bin_width"])                                                             # 162
    o["mM_r"] = o["m_r"] * o["M_r"]
    nbinsz = np.ceil(o["L"] / vcutwidth).astype(int)                                                             # 168
    testpos = _compound_centre(cyl_ag[:, 0], np.abs(charges_ag), inverse_ix)[inverse_ix]                         # 172-178
    rbins = np.digitize(testpos, o["bin_edges"][1:-1])
    z = (np.arange(nbinsz)) * (o["L"] / nbinsz)
    zbins = np.digitize(cyl_ag[:, 2], z[1:])
    curQz = np.bincount(rbins + n_bins * zbins, weights=charges_ag, minlength=n_bins * nbinsz).reshape(nbinsz, n_bins)
    curqz = np.cumsum(curQz, axis=0) / (o["bin_area"])[np.newaxis, :]                                            # 188
    o["m_z"] = -curqz.mean(axis=0)
    o["M_z"] = np.dot(charges_all, cyl_all[:, 2])                                                                # 192
    o["mM_z"] = o["m_z"] * o["M_z"]
    return o, o["M_z"]


def dielectric_sphere_frame(pos_ag, charges_ag, pos_all, charges_all, dimensions, n_bins, rmin=0.0, rmax=None):
    """modules/dielectricsphere.py:109-141 on top of core/sphere.py:85-104, 125-137."""
    pos_ag = np.asarray(pos_ag, dtype=np.float32)
    pos_all = np.asarray(pos_all, dtype=np.float32)
    box = np.asarray(dimensions, dtype=np.float32)[:3].astype(np.float64)
    center = box / 2
    rmin = np.float64(rmin)
    rmax = np.float64(box.min() / 2 if rmax is None else rmax)
    o = {}
    o["R"] = rmax - rmin
    o["bin_edges"] = np.linspace(rmin, rmax, n_bins + 1, endpoint=True)
    o["bin_width"] = o["R"] / n_bins
    o["bin_pos"] = o["bin_edges"][1:] - o["bin_width"] / 2
    r_ag = transform_sphere(pos_ag, center)[:, 0]
    r_all = transform_sphere(pos_all, center)[:, 0]
    curQ = np.bincount(np.digitize(r_ag, o["bin_edges"][1:-1]), weights=charges_ag, minlength=n_bins)
    o["m_r"] = -np.cumsum(curQ) / 4 / np.pi / o["bin_pos"] ** 2
    curQ_tot = np.bincount(np.digitize(r_all, o["bin_edges"][1:-1]), weights=charges_all, minlength=n_bins)
    o["m_r_tot"] = -np.cumsum(curQ_tot) / 4 / np.pi / o["bin_pos"] ** 2
    o["M_r"] = np.sum(o["m_r_tot"] * o["bin_width"])
    o["mM_r"] = o["m_r"] * o["M_r"]
    return o, o["M_r"]
