"""Path-2 measurement: fused weighted + count slab histogram (mb200_profile_hist) at config-4 size.

1M atoms (float32 xyz), 2154 bins (dz = 0.1 A on Lz = 215.37 A), static mass weights.
Reports frames/s and the HBM roofline fraction (algorithmic bytes: 12 B per atom per frame)
with the frames resident in HBM, the same through host pinned buffers (H2D inside), and numpy's two
np.histogram passes on one host core (planar.py:239-241, base.py:815-818) as the CPU baseline.
Prints one JSON line.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from maicos_b200 import _cabi  # noqa: E402

N, NB, F, LZ = 1_000_000, 2154, 32, 215.37
rng = np.random.default_rng(20261021)
pos = _cabi.pinned_empty((F, N, 3), np.float32, slot="bp")
pos[...] = rng.uniform(0, LZ, (F, N, 3)).astype(np.float32)
w = rng.uniform(1, 16, N)
edges = np.tile(np.linspace(np.float64(0), np.float64(LZ), NB + 1), (F, 1))
hw = np.zeros((F, NB)); hc = np.zeros((F, NB), dtype=np.int64)
ctx = _cabi.context(0)
lib = _cabi.load()
d_pos = torch.from_numpy(pos).cuda()
d_w = torch.from_numpy(w).cuda()
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(int(lib.mb200_ctx_stream(ctx.handle)), device=0)


def call(p, on_dev, wptr, w_on_dev):
    _cabi.check(lib.mb200_profile_hist(ctx.handle, 0, p, 1, on_dev, F, N, 2, wptr, 0, w_on_dev, _cabi.ptr(edges), NB,
                                       None, None, _cabi.ptr(hw), _cabi.ptr(hc)))


def timed(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, (time.perf_counter() - t0) / reps


def kernel_ms(mode, weighted=True, reps=10):
    """CUDA-event time of the histogram kernel alone (mb200_ctx_stats out[5]), median over calls."""
    ctx.set_hist_mode(mode)
    ctx.set_timing(True)
    ts = []
    for _ in range(reps + 3):
        call(d_pos.data_ptr(), 1, d_w.data_ptr() if weighted else None, 1)
        ts.append(ctx.stats()["contract_ns"] * 1e-6)
    ctx.set_timing(False)
    return float(np.median(ts[3:]))


variants = {}
for mode in ("atomics", "sorted", "chunks"):
    ctx.set_hist_mode(mode)
    s_call, _ = timed(lambda: call(d_pos.data_ptr(), 1, d_w.data_ptr(), 1), 20)
    k = kernel_ms(mode)
    variants[mode] = {"ms_per_call": s_call * 1e3, "kernel_ms": k, "kernel_gbs": 12.0 * N * F / k / 1e6,
                      "call_gbs": 12.0 * N * F / s_call / 1e9}
variants["counts_only"] = {"kernel_ms": kernel_ms("chunks", weighted=False)}
variants["counts_only"]["kernel_gbs"] = 12.0 * N * F / variants["counts_only"]["kernel_ms"] / 1e6
ctx.set_hist_mode("chunks")
dev_s, dev_wall = timed(lambda: call(d_pos.data_ptr(), 1, d_w.data_ptr(), 1), 20)
ref_c = np.histogram(pos[0, :, 2], NB, (np.float64(0), np.float64(LZ)))[0]  # float64 range as planar.py:107-110
assert np.array_equal(hc[0], ref_c)
host_s, host_wall = timed(lambda: call(_cabi.ptr(pos), 0, _cabi.ptr(w), 0), 5)
t0 = time.perf_counter()
for f in range(2):
    np.histogram(pos[f, :, 2], bins=NB, range=(np.float64(0), np.float64(LZ)), weights=w)
    np.histogram(pos[f, :, 2], bins=NB, range=(np.float64(0), np.float64(LZ)))
cpu = (time.perf_counter() - t0) / 2
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
peak = peaks.get("hbm_gbs", 6650.0)
gbs = 12.0 * N * F / dev_s / 1e9
print(json.dumps({
    "metric": "profile_atoms_frames_per_s", "workload": f"DensityPlanar-like slab histogram, {N} atoms, {NB} bins, {F} frames/call",
    "resident": {"frames_per_s": F / dev_s, "atoms_per_s": N * F / dev_s, "ms_per_call": dev_s * 1e3,
                 "wall_ms_per_call": dev_wall * 1e3},
    "e2e_host_pinned": {"frames_per_s": F / host_wall, "h2d_gbs": 12.0 * N * F / host_wall / 1e9},
    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                 "algorithmic_bytes_per_atom_frame": 12,
                 "kernel_only_frac": variants["chunks"]["kernel_gbs"] / peak},
    "variants": variants,
    "cpu_baseline": {"frames_per_s": 1 / cpu, "cores": 1, "kind": "numpy np.histogram weighted + count (reference call)"},
}))
