"""Host wrapper of ``mb200_profile_hist``: fused weighted + count slab / shell histograms.

Replaces the two ``np.histogram`` / ``np.histogram2d`` passes behind
``ProfileBase._single_frame`` (``/root/reference/src/maicos/core/base.py:815-818``;
``core/planar.py:235-243``, ``core/cylinder.py:229-243``, ``core/sphere.py:215-223``) for a batch
of frames.  Bin semantics are numpy's (values outside ``[edges[0], edges[-1]]`` dropped, right edge
inclusive, ``edges[i] <= x < edges[i+1]``), counts are exact.
"""

from __future__ import annotations

import numpy as np

from .. import _cabi

PLANAR, CYLINDER, SPHERE, SCALAR = 0, 1, 2, 3
DIGITIZE = 0x10  # or-ed into the geometry: np.digitize(x, edges[1:-1]) semantics (nothing dropped, ends clamp)


def profile_hist(geometry: int, positions: np.ndarray | None, dim: int, weights: np.ndarray | None, edges: np.ndarray,
                 center: np.ndarray | None = None, zrange: np.ndarray | None = None, device: int | None = None,
                 device_positions: tuple | None = None, device_weights: tuple | None = None, digitize: bool = False,
                 window: tuple | None = None, pack_boxes: np.ndarray | None = None, selection=None):
    """Histogram ``positions [F, n, 3]`` (float32 or float64) into ``edges [F, n_bins + 1]``.

    ``geometry = SCALAR``: ``positions [F, n]`` is the binned coordinate itself.  ``digitize=True``: the bin is
    ``np.digitize(x, edges[1:-1])`` (dielectric modules) instead of ``np.histogram``'s.

    ``weights``: ``None`` (counts only), ``[1, n]`` (same for all frames) or ``[F, n]``.
    ``device_positions = (pointer, F, n)``: float64 ``[F, n, 3]`` already in HBM (output of the
    compound prologue); ``device_weights = (pointer, per_frame)`` likewise.
    ``window = (first, n)``: ``positions`` holds whole frames ``[F, n_total, 3]`` and the group is atoms
    ``first .. first + n`` of each (``mb200_profile_hist_window``); weights are per group atom.
    ``selection = (handle, n)``: ``positions`` holds whole frames and the group is the ``n`` atoms listed in the device
    index list ``handle`` (``mb200_groups_create`` with one group), gathered on the device.
    ``pack_boxes``: float32 ``[F, 3]`` cell lengths -- float32 positions are wrapped into the primary cell on the device
    first (``AtomGroup.wrap(compound="atoms")``: ``x -= floor(x / L) * L`` in float32).
    Returns ``(hist_w float64 [F, n_bins] or None, hist_c int64 [F, n_bins])``.

    Accuracy of the weighted sums: every weight is quantised to ``Q = 2^(k - 52)`` with ``2^k >= max|w|`` of the frame
    (of the array, for static weights) and the sums are exact integers, so a bin's sum carries an ABSOLUTE error of at
    most ``count * max|w| * 2^-53`` -- what float64 accumulation has once its running sum reaches ``max|w|``, and
    independent of the summation order (bitwise reproducible).  A bin that holds only weights many orders of magnitude
    below ``max|w|`` therefore has a larger *relative* error than numpy's float64 sum of that bin alone (e.g. 1e-12
    relative needs ``|w| >= 1e-4 max|w|`` in the bin); counts are always exact.  Frames with non-finite weights and
    histograms with more than 4096 bins are accumulated in float64.
    """
    if device_positions is not None:
        pos_ptr, F, n = device_positions
        pos_is_f32, pos_on_dev = 0, 1
    else:
        positions = np.ascontiguousarray(positions)
        if geometry == SCALAR:
            if positions.ndim != 2:
                raise ValueError("a scalar coordinate must have shape (n_frames, n)")
        elif positions.ndim != 3 or positions.shape[2] != 3:
            raise ValueError("positions must have shape (n_frames, n, 3)")
        if positions.dtype not in (np.float32, np.float64):
            raise ValueError("positions must be float32 or float64")
        F, n = positions.shape[0], positions.shape[1]
        pos_ptr, pos_is_f32, pos_on_dev = _cabi.ptr(positions), int(positions.dtype == np.float32), 0
    n_total, first = n, 0
    sel_handle = None
    if selection is not None:
        sel_handle, n = selection[0], int(selection[1])
    elif window is not None:
        first, n = int(window[0]), int(window[1])
        if first < 0 or first + n > n_total:
            raise ValueError(f"window [{first}, {first + n}) does not fit frames of {n_total} atoms")
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    if edges.shape[0] != F:
        raise ValueError("edges must have one row per frame")
    n_bins = edges.shape[1] - 1
    per_frame, w_ptr, w_on_dev = 0, None, 0
    if device_weights is not None:
        w_ptr, per_frame, w_on_dev = device_weights[0], int(bool(device_weights[1]) and F > 1), 1
        weights = True  # weighted histogram requested
    elif weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        if weights.shape not in ((1, n), (F, n)):
            raise ValueError(f"weights must have shape (1, {n}) or ({F}, {n})")
        per_frame = int(weights.shape[0] == F and F > 1)
        w_ptr = _cabi.ptr(weights)
    if center is not None:
        center = np.ascontiguousarray(center, dtype=np.float64)
    if zrange is not None:
        zrange = np.ascontiguousarray(zrange, dtype=np.float64)
    if pack_boxes is not None:
        pack_boxes = np.ascontiguousarray(pack_boxes, dtype=np.float32)
        if pack_boxes.shape != (F, 3):
            raise ValueError("pack_boxes must have shape (n_frames, 3)")
    hist_c = np.zeros((F, n_bins), dtype=np.int64)
    hist_w = np.zeros((F, n_bins), dtype=np.float64) if weights is not None else None
    ctx = _cabi.context(device)
    _cabi.check(
        _cabi.load().mb200_profile_hist_window(
            ctx.handle, int(geometry) | (DIGITIZE if digitize else 0), pos_ptr, pos_is_f32, pos_on_dev, F, n_total, first, n,
            sel_handle, _cabi.ptr(pack_boxes), int(dim),
            w_ptr, per_frame, w_on_dev, _cabi.ptr(edges), n_bins, _cabi.ptr(center), _cabi.ptr(zrange),
            _cabi.ptr(hist_w), _cabi.ptr(hist_c),
        )
    )
    return hist_w, hist_c
