#!/usr/bin/env python
"""Headline benchmark: Saxs atom*q evaluations/s on a synthetic 100k-atom SPC/E box (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

* A *step* is one pass of the hot path over one batch of ``--frames-per-step`` frames
  (``mb200_saxs_frames``: wrap, phase tables, FP64 DMMA contraction for the O and the H group,
  f(q)^2, three |q| histograms).  Default K * B = 25 * 40 = the configuration's 1000 frames.
* ``value``: whole-job atom*q/s with the frames already resident in HBM (a pool of distinct frames
  larger than L2; every step reads a different batch).  Timed with CUDA events on the library's
  stream, barrier + synchronize on both sides, max over ranks.
* ``e2e``: the same metric through the public API ``maicos_b200.Saxs(atomgroup, qmax=2).run()`` on
  an in-memory trajectory: host staging, H2D from pinned memory and D2H of the binned results are
  inside the timed region (host wall clock bracketed by device synchronisation, max over ranks).
* ``roofline``: the dominant kernel ``k_sf_contract`` -- 8 algorithmic FP64 flop per atom*q
   over its CUDA-event duration, against the FP64 DMMA peak measured on this
  pool's B200 by ``tools/microbench/fp64_peaks`` (MEASURED_PEAKS.json carries no FP64 figure).
* ``cpu_baseline`` / ``--impl reference``: the reference's own Cython kernel (``oracle/_ref``,
  compiled from /root/reference; else the C port) + the NumPy restatement of Saxs._single_frame on
  the host cores, on a bounded sample (frames of the same workload).
Under torchrun (N > 1) frames are sharded over ranks (weak scaling: every rank gets K steps) and the
binned accumulators are merged by one allreduce at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_MOL = 33_333          # 99 999 atoms
BOX = 99.97             # Å, SPC/E liquid density
QMAX, DQ = 2.0, 0.1     # 20 nm^-1
SEED = 20261019
FP64_DMMA_PEAK_TFLOPS = 37.0   # tools/microbench/fp64_peaks on this pool's B200 (profiles/r01_fp64_peaks.log)
FLOP_PER_ATOMQ = 8.0           # one complex multiply-accumulate per atom*q
I8_PEAK_TOPS = 4359.0          # tcgen05 kind::i8 dense, M=128 N=256, tools/microbench/umma_i8 (profiles/r01_umma_i8.log)
I8_OPS_PER_ATOMQ = 120.0       # 4 real MACs x 15 digit products (s + t <= 6 of 5 x 5) x 2 ops
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the contraction kernel (one `ncu --set full` capture of
# `bench.py --steps 2 --warmup 1`, 40 frames per launch): profiles/r01_tc_ncu_details.txt, r01_contract_v2_ncu_details.txt
NCU_TRAFFIC_BYTES = {"i8x5": 5.238e9 + 0.256e9, "fp64": 17.944e9 + 0.093e9}
NCU_TRAFFIC_FRAMES = 40


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=40)
    ap.add_argument("--pool-steps", type=int, default=7, help="distinct batches resident in HBM (pool > L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", default="i8x5", choices=["i8x5", "fp64"],
                    help="arithmetic of the contraction for the headline numbers (i8x5 = the library default: split "
                         "precision on the tcgen05 tensor cores; fp64 = FP64 DMMA); the other mode is always measured too "
                         "and reported under other_precision_mode")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def workload():
    from maicos_b200.lib.math import sfactor_maxn
    from oracle import kernel

    dims = np.double(np.float32([BOX, BOX, BOX]))
    q, _ = kernel.port_structure_factor(np.zeros((1, 3)), dims, 0.0, QMAX, 0.0, np.pi, np.ones(1))
    n_q = int(np.count_nonzero(q))
    return {"n_atoms": 3 * N_MOL, "n_q": n_q, "maxn": [int(x) for x in sfactor_maxn(dims, QMAX)],
            "atomq_per_frame": 3 * N_MOL * n_q}


def config_dict(args, wl, n_gpus):
    return {
        "workload": "Saxs on synthetic 100k-atom SPC/E water box (33333 molecules, L=99.97 A), qmax=2.0/A "
                    "(20/nm), dq=0.1/A, element groups O+H (BASELINE.json configs[1])",
        "n_atoms": wl["n_atoms"], "n_q_valid": wl["n_q"], "maxn": wl["maxn"], "frames_per_step": args.frames_per_step,
        "frames_total": args.steps * args.frames_per_step * n_gpus, "parallelism": f"frames sharded over {n_gpus} GPU(s)",
        "l2": f"inputs larger than L2: pool of {args.pool_steps * args.frames_per_step} distinct frames "
              f"({args.pool_steps * args.frames_per_step * wl['n_atoms'] * 12 / 1e6:.0f} MB) resident, every step reads a "
              "different batch; phase tables (>1 GB per step) are rewritten every step",
        "precision": PRECISION_TEXT[args.precision],
    }


PRECISION_TEXT = {
    "i8x5": "split precision (library default): FP64 phase tables cut into five int8 digits (40-bit fixed point), exact int32 "
            "accumulation on tcgen05 kind::i8 (TMEM), FP64 recombination; measured rel err <= 5e-9 per S cell, 3e-13 binned "
            "(stated tolerance 1e-6)",
    "fp64": "FP64 tables, FP64 DMMA accumulation (rel err ~1e-12 vs the reference kernel; stated tolerance 1e-6)",
}


def run_reference(args, wl, rank):
    """--impl reference: the reference's CPU implementation of the same path on the host cores."""
    if rank != 0:
        return
    from maicos_b200.synthetic import spce_topology, spce_water_frames
    from oracle import kernel, restatement

    n_fr = args.steps + args.warmup
    frames = spce_water_frames(N_MOL, BOX, SEED, n_frames=min(n_fr, 8))
    el = spce_topology(N_MOL)["elements"].astype(str)
    box = np.float32([BOX, BOX, BOX])
    n_bins = int(np.ceil(QMAX / DQ))
    stats = restatement.RunningStats()
    t0 = None
    for k in range(n_fr):
        if k == args.warmup:
            t0 = time.perf_counter()
        pos = restatement.wrap_atoms_f32(frames[k % len(frames)], box)
        obs, _ = restatement.saxs_single_frame(pos, box, el, 0.0, QMAX, 0.0, np.pi, n_bins)
        stats.update(obs)
    dt = time.perf_counter() - t0
    res = restatement.saxs_conclude(stats.sums, stats.sems, wl["n_atoms"], 0.0, QMAX, DQ)
    assert np.all(np.isfinite(res["structure_factors"]))
    value = args.steps * wl["atomq_per_frame"] / dt
    cores = kernel.n_threads()
    line = {
        "impl": "reference", "metric": "saxs_atom_q_evals_per_s", "value": value, "unit": "atom*q/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, wl, 1),
        "frames_per_s": args.steps / dt,
        "cpu_baseline": {"value": value, "unit": "atom*q/s", "cores": cores, "kind": kernel.kind(),
                         "sample": f"{args.steps} frame(s) of the workload (1 frame = O + H group, 1.8e9 atom*q) "
                                   "per step through the Cython kernel + NumPy Saxs frame body"},
        "e2e": {"value": value, "unit": "atom*q/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = workload()
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return

    import torch

    import maicos_b200 as mb
    from maicos_b200 import _cabi, parallel
    from maicos_b200.lib import tables
    from maicos_b200.mdshim import MemoryTrajectory, Universe
    from maicos_b200.synthetic import spce_topology, spce_water_frames

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: maicos_b200 has no CPU fallback")
    os.environ.setdefault("MAICOS_B200_DEVICE", str(local))
    torch.cuda.set_device(local)
    parallel.init(mode="frames")
    ctx = _cabi.context(local)
    ctx.set_timing(True)
    ctx.set_precision(args.precision)
    lib = _cabi.load()

    B, K, W, P = args.frames_per_step, args.steps, args.warmup, args.pool_steps
    n = wl["n_atoms"]
    n_bins = int(np.ceil(QMAX / DQ))
    topo = spce_topology(N_MOL)
    pool = spce_water_frames(N_MOL, BOX, SEED + 1000 * rank, n_frames=P * B)  # [P*B, n, 3] float32
    d_pool = torch.from_numpy(pool).cuda()
    boxes = np.tile(np.float32([BOX, BOX, BOX]), (B, 1))
    elements = topo["elements"].astype(str)
    groups = [np.nonzero(elements == e)[0].astype(np.int32) for e in np.unique(elements)]
    gptr = np.concatenate([[0], np.cumsum([len(g) for g in groups])]).astype(np.int64)
    gflat = np.ascontiguousarray(np.concatenate(groups))
    cm = np.ascontiguousarray(np.stack([tables.cm_packed(e) for e in np.unique(elements)]))
    G = len(groups)
    s_b = np.zeros((B, G, n_bins)); i_b = np.zeros((B, G, n_bins)); c_b = np.zeros((B, G, n_bins), dtype=np.int64)
    acc = np.zeros(3 * n_bins)

    def step(k):
        ptr_dev = d_pool.data_ptr() + (k % P) * B * n * 12
        _cabi.check(lib.mb200_saxs_frames(ctx.handle, ptr_dev, 1, B, n, _cabi.ptr(boxes), _cabi.ptr(gflat),
                                          _cabi.ptr(gptr), G, _cabi.ptr(cm), 0.0, QMAX, 0.0, float(np.pi), n_bins, 1,
                                          0, 1, _cabi.ptr(s_b), _cabi.ptr(i_b), _cabi.ptr(c_b)))
        acc[:n_bins] += s_b.sum((0, 1)); acc[n_bins:2 * n_bins] += i_b.sum((0, 1)); acc[2 * n_bins:] += c_b[:, -1].sum(0)
        return ctx.stats()

    stream = torch.cuda.ExternalStream(int(lib.mb200_ctx_stream(ctx.handle)), device=local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(W):
        step(k)
    acc[:] = 0.0  # accumulate the timed steps only
    parallel.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    launches0 = ctx.launch_count()
    kern_ns, kern_launches = 0, 0
    ev0.record(stream)
    for k in range(K):
        st = step(W + k)
        kern_ns += st["contract_ns"]; kern_launches += 1
    parallel.allreduce_sum(acc)  # the single end-of-run merge of the binned accumulators
    ev1.record(stream)
    torch.cuda.synchronize(); parallel.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    t_all = np.array([ms]);
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms = float(tt.item())
    atomq_total = float(K) * B * wl["atomq_per_frame"] * world
    value = atomq_total / (ms * 1e-3)

    # ---- the other arithmetic mode, device resident (same steps, same frames) -------------------
    other = "i8x5" if args.precision == "fp64" else "fp64"
    ctx.set_precision(other)
    ref_acc = acc.copy()
    acc[:] = 0.0
    for k in range(W):
        step(k)
    acc[:] = 0.0
    torch.cuda.synchronize()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record(stream)
    o_ns = 0
    for k in range(K):
        o_ns += step(W + k)["contract_ns"]
    o1.record(stream)
    torch.cuda.synchronize()
    o_ms = o0.elapsed_time(o1)
    parallel.allreduce_sum(acc)
    nzb = ref_acc[:n_bins] != 0
    other_mode = {
        "precision": other, "value": float(K) * B * wl["atomq_per_frame"] / (o_ms * 1e-3), "unit": "atom*q/s (this rank, resident)",
        "ms_per_step": o_ms / K, "kernel_ms_per_launch": o_ns * 1e-6 / K,
        "max_rel_diff_binned_S_vs_headline_mode": float(np.max(np.abs(acc[:n_bins][nzb] / ref_acc[:n_bins][nzb] - 1))),
        "bincount_equal": bool(np.array_equal(acc[2 * n_bins:], ref_acc[2 * n_bins:])),
    }
    acc[:] = ref_acc
    ctx.set_precision(args.precision)

    # ---- end to end through the public API (host trajectory -> Saxs.run()) ----------------------
    e2e = None
    if not args.no_e2e:
        def universe(n_frames, offset):
            def src(f):  # a frame of the in-memory pool as a read-only view (what MDAnalysis' MemoryReader hands out):
                v = pool[(offset + f) % len(pool)].view()  # the one host copy of a frame is trajectory -> pinned staging
                v.flags.writeable = False
                return v
            traj = MemoryTrajectory(dimensions=np.float32([BOX, BOX, BOX, 90, 90, 90]), n_frames=n_frames,
                                    frame_source=src, n_atoms=n)
            return Universe(n, traj, **topo)

        mb.Saxs._batch_size = B
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mb.Saxs(universe(W * B * world, 0).atoms, qmax=QMAX, dq=DQ).run()  # warm-up: W steps per rank
            u = universe(K * B * world, 17)
            parallel.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            saxs = mb.Saxs(u.atoms, qmax=QMAX, dq=DQ).run()   # K steps per rank + one merging allreduce
            torch.cuda.synchronize(); parallel.barrier()
            dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            dt = float(tt.item())
        assert saxs._index == K * B * world and np.all(np.isfinite(saxs.results.scattering_intensities))
        e2e = {"value": atomq_total / dt, "unit": "atom*q/s", "h2d_bytes_per_step": int(B * n * 12 + B * 12 + n * 4),
               "d2h_bytes_per_step": int(B * G * n_bins * 24), "frames_per_s": K * B * world / dt,
               "api": "maicos_b200.Saxs(atomgroup, qmax=2.0, dq=0.1).run()", "ms_per_step": dt / K * 1e3}

    if rank != 0:
        parallel.shutdown()
        return
    kern_s = kern_ns * 1e-9
    atomq_per_launch = B * wl["atomq_per_frame"]
    peaks_file = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    common = {"kernel_ms_per_launch": kern_s / max(kern_launches, 1) * 1e3, "kernel_share_of_step": kern_s / (ms * 1e-3),
              "traffic": NCU_TRAFFIC_BYTES[args.precision] * B / NCU_TRAFFIC_FRAMES,
              "traffic_note": "DRAM bytes per launch from the committed ncu capture (not measured in this run): the tables are "
                              "re-read from HBM by every row tile; the kernel is compute-side bound (DRAM < 10 % of peak)",
              "hbm_gbs_measured_peak": peaks_file.get("hbm_gbs"),
              "fp64_equivalent_tflops": FLOP_PER_ATOMQ * atomq_per_launch * kern_launches / kern_s / 1e12 if kern_s > 0 else None,
              "fp64_dmma_peak_tflops": FP64_DMMA_PEAK_TFLOPS}
    if args.precision == "fp64":
        achieved = common["fp64_equivalent_tflops"]
        roofline = {
            "kernel": "k_sf_contract", "bound": "tensor", "pipe": "FP64 DMMA (mma.sync.m8n8k4.f64)",
            "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
            "frac": achieved / FP64_DMMA_PEAK_TFLOPS if achieved else None,
            "peak_source": "measured FP64 DMMA burst on this pool's B200 (tools/microbench/fp64_peaks.cu, "
                           "profiles/r01_fp64_peaks.log); MEASURED_PEAKS.json has no FP64 entry",
            "algorithmic_flop_per_atomq": FLOP_PER_ATOMQ, **common}
    else:
        achieved = I8_OPS_PER_ATOMQ * atomq_per_launch * kern_launches / kern_s / 1e12 if kern_s > 0 else None
        roofline = {
            "kernel": "k_sf_contract_tc", "bound": "tensor", "pipe": "tcgen05.mma kind::i8 (UTCIMMA), int32 accumulators in TMEM",
            "achieved": achieved, "peak": I8_PEAK_TOPS, "unit": "TFLOP/s",
            "unit_note": "int8 tensor ops (multiply-add = 2), i.e. TOP/s; MEASURED_PEAKS.json carries only the bf16 figure "
                         f"({peaks_file.get('bf16_tflops')} TFLOP/s), the int8 peak is our own probe of the same pipe",
            "frac": achieved / I8_PEAK_TOPS if achieved else None,
            "peak_source": "measured tcgen05 kind::i8 dense rate on this pool's B200, 148 CTAs x 128x256x32 MMAs from shared "
                           "memory (tools/microbench/umma_i8.cu, profiles/r01_umma_i8.log); nominal 4500",
            "algorithmic_int8_ops_per_atomq": I8_OPS_PER_ATOMQ,
            "kernel_variant": "k_sf_contract_tc<true>: generated A operand written to tensor memory (tcgen05.st), l tile <= 32",
            "co_limiters": "no pipe is saturated (ncu: tensor 41 %, FP64 22 %, shared-memory pipe 42 %, issue slots 52 %): a K-step "
                           "is paced by the MMA warp's serial loop -- ~620 cycles issuing six UTCIMMA (480 cycles of tensor work at "
                           "N = 64 columns per digit slice) + ~550 cycles in two mbarrier try_waits; traces and the rejected "
                           "variants in profiles/r01_tc_notes.md (session 4)",
            **common}
    line = {
        "metric": "saxs_atom_q_evals_per_s", "value": value, "unit": "atom*q/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i8x5->s32 (40-bit fixed point, FP64 tables and recombination)" if args.precision == "i8x5" else "f64",
        "data": "synthetic", "config": config_dict(args, wl, world),
        "frames_per_s": K * B * world / (ms * 1e-3), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "precision": args.precision, "other_precision_mode": other_mode,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], ref_obs, ref_frame = cpu_baseline(wl)
        # parity next to the throughput: the same frame through the C ABI vs the reference kernel's result
        s1 = np.zeros((1, G, n_bins)); i1 = np.zeros((1, G, n_bins)); c1 = np.zeros((1, G, n_bins), dtype=np.int64)
        fr = np.ascontiguousarray(ref_frame[None])
        _cabi.check(lib.mb200_saxs_frames(ctx.handle, _cabi.ptr(fr), 0, 1, n, _cabi.ptr(boxes[:1].copy()), _cabi.ptr(gflat),
                                          _cabi.ptr(gptr), G, _cabi.ptr(cm), 0.0, QMAX, 0.0, float(np.pi), n_bins, 1,
                                          0, 1, _cabi.ptr(s1), _cabi.ptr(i1), _cabi.ptr(c1)))
        nz = ref_obs["structure_factors"] != 0
        line["parity"] = {
            "checked": "frame 0 of the cpu_baseline sample: binned S, I, bincount vs reference Cython kernel + NumPy body",
            "max_rel_err_S_binned": float(np.max(np.abs(s1[0].sum(0)[nz] / ref_obs["structure_factors"][nz] - 1))),
            "max_rel_err_I_binned": float(np.max(np.abs(i1[0].sum(0)[nz] / ref_obs["scattering_intensities"][nz] - 1))),
            "bincount_equal": bool(np.array_equal(c1[0, -1], ref_obs["bincount"])),
            "tolerance": "rel <= 1e-6, counts exact",
        }
    if world == 1 and not args.no_e2e:
        line["path2_profile_hist"] = path2_summary(ctx, lib, torch)
    print(json.dumps(line))
    parallel.shutdown()


def path2_summary(ctx, lib, torch):
    """Secondary hot path (slab histogram, BASELINE.json configs[3] size): frames/s and HBM roofline fraction."""
    from maicos_b200 import _cabi

    n, nb, F, lz = 1_000_000, 2154, 16, 215.37
    rng = np.random.default_rng(20261021)
    d_pos = torch.from_numpy(rng.uniform(0, lz, (F, n, 3)).astype(np.float32)).cuda()
    d_w = torch.from_numpy(rng.uniform(1, 16, n)).cuda()
    edges = np.tile(np.linspace(np.float64(0), np.float64(lz), nb + 1), (F, 1))
    hw = np.zeros((F, nb)); hc = np.zeros((F, nb), dtype=np.int64)
    stream = torch.cuda.ExternalStream(int(lib.mb200_ctx_stream(ctx.handle)))
    torch.cuda.synchronize()

    def call():
        _cabi.check(lib.mb200_profile_hist(ctx.handle, 0, d_pos.data_ptr(), 1, 1, F, n, 2, d_w.data_ptr(), 0, 1,
                                           _cabi.ptr(edges), nb, None, None, _cabi.ptr(hw), _cabi.ptr(hc)))
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        call()
    e1.record(stream)
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 10 * 1e-3
    # the histogram kernel alone: CUDA events around it inside the call (mb200_ctx_stats out[5])
    lib.mb200_ctx_set_timing(ctx.handle, 1)
    ks = []
    for _ in range(5):
        call()
        ks.append(ctx.stats()["contract_ns"] * 1e-9)
    lib.mb200_ctx_set_timing(ctx.handle, 0)
    ksec = float(np.median(ks))
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    gbs = 12.0 * n * F / sec / 1e9
    return {"workload": f"weighted + count slab histogram, {n} atoms x {F} frames resident, {nb} bins",
            "frames_per_s": F / sec, "atoms_per_s": n * F / sec,
            "roofline": {"kernel": "k_hist_sorted", "bound": "hbm", "achieved": 12.0 * n * F / ksec / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": 12.0 * n * F / ksec / 1e9 / peak, "kernel_ms_per_launch": ksec * 1e3,
                         "whole_call": {"achieved": gbs, "frac": gbs / peak, "ms_per_call": sec * 1e3,
                                        "note": "call = edges upload + max|w| + histogram + merge kernels + per-frame results to host"},
                         "co_limiter": "shared-memory atomic unit: one ATOMS lane per clock and SM (profiles/r01_hist_notes.md)",
                         "algorithmic_bytes_per_atom_frame": 12,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}}


def cpu_baseline(wl):
    """Reference CPU path on the host cores, bounded sample (about 10-30 s of CPU work)."""
    from maicos_b200.synthetic import spce_topology, spce_water_frames
    from oracle import kernel, restatement

    frames = spce_water_frames(N_MOL, BOX, SEED, n_frames=3)
    el = spce_topology(N_MOL)["elements"].astype(str)
    box = np.float32([BOX, BOX, BOX])
    n_bins = int(np.ceil(QMAX / DQ))
    restatement.saxs_single_frame(restatement.wrap_atoms_f32(frames[0][:3000], box), box, el[:3000], 0.0, QMAX, 0.0,
                                  np.pi, n_bins)  # warm the OpenMP pool
    t0 = time.perf_counter()
    done = 0
    first_obs = None
    while done < 8 and (time.perf_counter() - t0 < 12.0 or done < 2):
        pos = restatement.wrap_atoms_f32(frames[done % 3], box)
        obs, _ = restatement.saxs_single_frame(pos, box, el, 0.0, QMAX, 0.0, np.pi, n_bins)
        if first_obs is None:
            first_obs = obs
        done += 1
    dt = time.perf_counter() - t0
    return {"value": done * wl["atomq_per_frame"] / dt, "unit": "atom*q/s", "cores": kernel.n_threads(),
            "kind": kernel.kind(), "frames_per_s": done / dt,
            "sample": f"{done} full frame(s) of the workload (O + H groups, {wl['atomq_per_frame']:.3e} atom*q each) "
                      f"through the Cython kernel + NumPy Saxs frame body in {dt:.1f} s"}, first_obs, frames[0]


if __name__ == "__main__":
    main()
