#!/usr/bin/env python
"""Headline benchmark: Saxs atom*q evaluations/s on a synthetic 100k-atom SPC/E box (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-configs]

* A *step* is one pass of the hot path over one batch of ``--frames-per-step`` frames (``mb200_saxs_frames``: wrap,
  phase tables, contraction for the O and the H group, f(q)^2, three |q| histograms).  Default K * B = 25 * 40 = the
  configuration's 1000 frames.
* ``value``: whole-job atom*q/s with the frames already resident in HBM (a pool of distinct frames larger than L2;
  every step reads a different batch).  CUDA events on the library's stream, barrier + synchronize on both sides,
  max over ranks.
* ``e2e``: the same metric through the public API ``maicos_b200.Saxs(atomgroup, qmax=2).run()`` on an in-memory
  trajectory: host staging, H2D from pinned memory and D2H of the binned results are inside the timed region.
* ``roofline``: the dominant kernel of the selected arithmetic against the ceiling of its pipe **probed in this run**
  (``mb200_probe_peaks``: tcgen05 kind::i8 dense rate, FP64 DMMA rate, HBM copy rate), next to MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the reference's own Cython kernel (``oracle/_ref``, compiled from
  /root/reference; else the C port) + the NumPy restatement of ``Saxs._single_frame`` on ALL host cores
  (``torch.distributed.run`` exports OMP_NUM_THREADS=1; the pool is resized explicitly), one frame per step.
* ``configs``: every other BASELINE.json configuration (C1 shipped water, C3 DiporderStructureFactor on 300k atoms at
  qmax 2 and 6, C4 DensityPlanar + DiporderPlanar on a 1M-atom slab, C5 one 1M-atom frame at qmax 3, and north_star's
  1M-atom / qmax 2 target) as bounded runs through the public classes from host memory, each with a parity figure
  against the oracle next to its time.  With ``--gpus N > 1``: C5 sharded over q-blocks (strong scaling, device and
  wall time) and C3 sharded over frames.
Under torchrun (N > 1) the headline frames are sharded over ranks (weak scaling: every rank gets K steps) and the binned
accumulators are merged by one allreduce at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_MOL = 33_333          # 99 999 atoms
BOX = 99.97             # Å, SPC/E liquid density
QMAX, DQ = 2.0, 0.1     # 20 nm^-1
SEED = 20261019
FLOP_PER_ATOMQ = 8.0           # one complex multiply-accumulate per atom*q
I8_OPS_PER_ATOMQ = 120.0       # 4 real MACs x 15 digit products (s + t <= 6 of 5 x 5) x 2 ops
# fallbacks when the in-run probe fails (profiles/r01_fp64_peaks.log, profiles/r01_umma_i8.log)
FP64_DMMA_PEAK_TFLOPS_FALLBACK = 37.0
I8_PEAK_TOPS_FALLBACK = 4359.0
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the contraction kernel from the committed `ncu --set full`
# capture of `bench.py --steps 2 --warmup 1` (40 frames per launch); see profiles/ for the round's capture
NCU_TRAFFIC_BYTES = {"i8x5": 5.228e9 + 0.263e9, "fp64": 17.944e9 + 0.093e9}
NCU_TRAFFIC_FRAMES = 40
NCU_TRAFFIC_SOURCE = {"i8x5": "profiles/r02_tc_atm_ncu_details.txt (round-2 build)", "fp64": "profiles/r01_contract_v2_ncu_details.txt"}


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
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded runs of the other BASELINE configurations")
    ap.add_argument("--configs", default="", help="comma-separated subset of C1,C3,C4,C5,T1M (default: all)")
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
            self._stop_evt.wait(0.05)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def n_valid_q(dims, qmax):
    """Accepted Miller vectors of a box (the reference's own acceptance test, through the oracle's C port)."""
    from oracle import kernel

    q, _ = kernel.port_structure_factor(np.zeros((1, 3)), np.double(np.float32(dims)), 0.0, qmax, 0.0, np.pi, np.ones(1))
    return int(np.count_nonzero(q))


def workload():
    from oracle import kernel  # metadata only (maxn, number of accepted q); the product library is not touched here

    dims = np.double(np.float32([BOX, BOX, BOX]))
    n_q = n_valid_q(dims, QMAX)
    return {"n_atoms": 3 * N_MOL, "n_q": n_q, "maxn": [int(x) for x in kernel.port_maxn(dims, QMAX)],
            "atomq_per_frame": 3 * N_MOL * n_q}


def config_dict(args, wl, n_gpus):
    """The workload, identical for both arms (what a *step* of each arm covers is stated under ``batch`` / ``sample``)."""
    return {
        "workload": "Saxs on synthetic 100k-atom SPC/E water box (33333 molecules, L=99.97 A), qmax=2.0/A "
                    "(20/nm), dq=0.1/A, element groups O+H (BASELINE.json configs[1])",
        "n_atoms": wl["n_atoms"], "n_q_valid": wl["n_q"], "maxn": wl["maxn"],
        "parallelism": f"frames sharded over {n_gpus} GPU(s)",
        "l2": f"inputs larger than L2: pool of {args.pool_steps * args.frames_per_step} distinct frames "
              f"({args.pool_steps * args.frames_per_step * wl['n_atoms'] * 12 / 1e6:.0f} MB) resident, every step reads a "
              "different batch; phase tables (>1 GB per step) are rewritten every step",
    }


PRECISION_TEXT = {
    "i8x5": "split precision (library default): FP64 phase tables cut into five int8 digits (40-bit fixed point), exact int32 "
            "accumulation on tcgen05 kind::i8 (TMEM), FP64 recombination; measured rel err <= 5e-9 per S cell, 3e-13 binned "
            "(stated tolerance 1e-6)",
    "fp64": "FP64 tables, FP64 DMMA accumulation (rel err ~1e-12 vs the reference kernel; stated tolerance 1e-6)",
}


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU implementation of the same path on all host cores, one frame per step."""
    if rank != 0:
        return
    from maicos_b200.synthetic import spce_topology, spce_water_frames   # input generator only (pure NumPy)
    from oracle import kernel, restatement

    cores = kernel.set_threads()  # torch.distributed.run exports OMP_NUM_THREADS=1: size the pool explicitly
    n_fr = args.steps + args.warmup
    frames = spce_water_frames(N_MOL, BOX, SEED, n_frames=min(n_fr, 8))
    el = spce_topology(N_MOL)["elements"].astype(str)
    box = np.float32([BOX, BOX, BOX])
    n_bins = int(np.ceil(QMAX / DQ))
    stats = restatement.RunningStats()
    t0 = time.perf_counter()
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
    sample = (f"1 frame per step ({args.steps} timed frames, {args.warmup} warm-up; 1 frame = O + H group = "
              f"{wl['atomq_per_frame']:.3e} atom*q) through the reference's Cython kernel + the NumPy Saxs frame body on "
              f"{cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": "saxs_atom_q_evals_per_s", "value": value, "unit": "atom*q/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, wl, world),
        "sample": {"frames_per_step": 1, "frames_total": args.steps, "note": sample},
        "frames_per_s": args.steps / dt,
        "cpu_baseline": {"value": value, "unit": "atom*q/s", "cores": cores, "kind": kernel.kind(), "sample": sample},
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
        run_reference(args, wl, rank, world)
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

    # the element groups go to the device once, as in Saxs.run() (mb200_groups_create; the array form of the call would
    # fingerprint the 100k indices in every step)
    import ctypes

    groups_handle = ctypes.c_void_p()
    _cabi.check(lib.mb200_groups_create(ctx.handle, _cabi.ptr(gflat), _cabi.ptr(gptr), G, n, ctypes.byref(groups_handle)))

    # a step is submitted behind the one that is running and collected afterwards (mb200_saxs_frames_submit / _collect, the
    # pair Saxs.run() uses): every step's kernels, results and host-side accumulation lie inside the timed region
    def submit(k):
        ptr_dev = d_pool.data_ptr() + (k % P) * B * n * 12
        ticket = ctypes.c_int32(-1)
        _cabi.check(lib.mb200_saxs_frames_submit(ctx.handle, ptr_dev, 1, B, _cabi.ptr(boxes), groups_handle, _cabi.ptr(cm),
                                                 0.0, QMAX, 0.0, float(np.pi), n_bins, 1, 0, 1, ctypes.byref(ticket)))
        return int(ticket.value)

    def collect(ticket):
        _cabi.check(lib.mb200_saxs_frames_collect(ctx.handle, ticket, _cabi.ptr(s_b), _cabi.ptr(i_b), _cabi.ptr(c_b)))
        acc[:n_bins] += s_b.sum((0, 1)); acc[n_bins:2 * n_bins] += i_b.sum((0, 1)); acc[2 * n_bins:] += c_b[:, -1].sum(0)
        return ctx.stats()

    def step(k):
        return collect(submit(k))

    def steps(first, count):
        """`count` steps, each queued behind its predecessor; yields the statistics of every step."""
        waiting = None
        for k in range(first, first + count):
            ticket = submit(k)
            if waiting is not None:
                yield collect(waiting)
            waiting = ticket
        if waiting is not None:
            yield collect(waiting)

    stream = torch.cuda.ExternalStream(int(lib.mb200_ctx_stream(ctx.handle)), device=local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(W):
        step(k)
    # ceilings of the pipes, probed in this process on this GPU, warm, right before the timed region
    probe_sampler = ClockSampler(local); probe_sampler.start()
    try:
        peaks = ctx.probe_peaks()
    except _cabi.MB200Error as err:  # keep the run alive: the fallbacks are labelled as such
        peaks = {"i8_tops": 0.0, "fp64_dmma_tflops": 0.0, "hbm_copy_gbs": 0.0, "error": str(err)}
    peaks["clocks_during_probe"] = probe_sampler.stop()
    step(0)
    acc[:] = 0.0  # accumulate the timed steps only
    parallel.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    launches0 = ctx.launch_count()
    kern_ns, kern_launches = 0, 0
    ev0.record(stream)
    mma_count = 0
    for st in steps(W, K):
        kern_ns += st["contract_ns"]; kern_launches += 1
        mma_count += st["dmma_steps"]
    parallel.allreduce_sum(acc)  # the single end-of-run merge of the binned accumulators
    ev1.record(stream)
    torch.cuda.synchronize(); parallel.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
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
    for st in steps(W, K):
        o_ns += st["contract_ns"]
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
        # the in-memory trajectory of this rank (K * B frames, the pool tiled) in PAGE-LOCKED host memory -- frames go
        # trajectory -> PCIe -> HBM with no staging copy -- and the same frames served from ordinary pageable memory (one
        # host copy per frame into the pinned batch buffer).  Rank r analyses frames r, r + W, ...: its local frame j is
        # global frame j * W + r.
        n_local = max(K, W) * B
        traj_pinned = _cabi.pinned_empty((n_local, n, 3), np.float32, slot="bench:traj")
        for j0 in range(0, n_local, len(pool)):
            np.copyto(traj_pinned[j0:j0 + len(pool)], pool[:min(len(pool), n_local - j0)])

        def universe(n_frames, pinned):
            def src(f):  # a frame as a read-only view (what MDAnalysis' MemoryReader hands out)
                v = (traj_pinned[f // world] if pinned else pool[(f // world) % len(pool)]).view()
                v.flags.writeable = False
                return v
            rows = (lambda f: (traj_pinned, f // world)) if pinned else None
            traj = MemoryTrajectory(dimensions=np.float32([BOX, BOX, BOX, 90, 90, 90]), n_frames=n_frames,
                                    frame_source=src, n_atoms=n, pinned_rows=rows)
            return Universe(n, traj, **topo)

        def run_e2e(pinned):
            mb.Saxs(universe(W * B * world, pinned).atoms, qmax=QMAX, dq=DQ).run()  # warm-up: W steps per rank
            u = universe(K * B * world, pinned)
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
            return dt, saxs

        mb.Saxs._batch_size = B
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dt, saxs = run_e2e(True)
            dt_pg, saxs_pg = run_e2e(False)
        e2e = {"value": atomq_total / dt, "unit": "atom*q/s", "h2d_bytes_per_step": int(B * n * 12 + B * 12 + n * 4),
               "d2h_bytes_per_step": int(B * G * n_bins * 24), "frames_per_s": K * B * world / dt,
               "api": "maicos_b200.Saxs(atomgroup, qmax=2.0, dq=0.1).run()", "ms_per_step": dt / K * 1e3,
               "input": "in-memory trajectory in page-locked host memory (mdshim.MemoryTrajectory(pinned_rows=...)): every step's "
                        "frames are copied host -> device inside the timed region, straight from the trajectory",
               "host_timings_s": saxs.timings,
               "pageable_trajectory": {"value": atomq_total / dt_pg, "frames_per_s": K * B * world / dt_pg,
                                       "ms_per_step": dt_pg / K * 1e3, "host_timings_s": saxs_pg.timings,
                                       "note": "same run with the trajectory in ordinary pageable memory: one host copy per "
                                               "frame into the pinned batch buffer, then the same H2D"}}
        mb.Saxs._batch_size = 32
        _cabi.release_pinned("bench:traj")

    # ---- the other BASELINE configurations (bounded, with parity) -- every rank takes part in the sharded ones ----
    configs = None
    if not args.no_configs:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            configs = run_configs(args, ctx, lib, torch, rank, world, peaks)

    if rank != 0:
        parallel.shutdown()
        return
    kern_s = kern_ns * 1e-9
    atomq_per_launch = B * wl["atomq_per_frame"]
    peaks_file = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    i8_peak = peaks["i8_tops"] if peaks.get("i8_tops", 0) > 0 else I8_PEAK_TOPS_FALLBACK
    f64_peak = peaks["fp64_dmma_tflops"] if peaks.get("fp64_dmma_tflops", 0) > 0 else FP64_DMMA_PEAK_TFLOPS_FALLBACK
    common = {"kernel_ms_per_launch": kern_s / max(kern_launches, 1) * 1e3, "kernel_share_of_step": kern_s / (ms * 1e-3),
              "traffic": NCU_TRAFFIC_BYTES[args.precision] * B / NCU_TRAFFIC_FRAMES,
              "traffic_note": f"DRAM bytes per launch from the committed ncu capture {NCU_TRAFFIC_SOURCE[args.precision]} (not "
                              "measured in this run).  It equals the size of the phase tables + Z image of the batch (5.2 GB for 40 "
                              "frames x 100k atoms x 1296 B): k_sf_tables writes them once, the contraction reads every byte once "
                              "from HBM (the seven row tiles of an atom range run concurrently and share it through L2, hit rate "
                              "73 %); the 48 MB of positions are the algorithmic input.  DRAM is at 9 % of its peak during the "
                              "contraction: the kernel is bound on the SM side",
              "hbm_gbs_measured_peak": peaks_file.get("hbm_gbs"),
              "fp64_equivalent_tflops": FLOP_PER_ATOMQ * atomq_per_launch * kern_launches / kern_s / 1e12 if kern_s > 0 else None,
              "fp64_dmma_peak_tflops": f64_peak,
              "peaks_probed_in_this_run": peaks,
              "peaks_file": {k: peaks_file.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")}}
    if args.precision == "fp64":
        achieved = common["fp64_equivalent_tflops"]
        roofline = {
            "kernel": "k_sf_contract", "bound": "tensor", "pipe": "FP64 DMMA (mma.sync.m8n8k4.f64)",
            "achieved": achieved, "peak": f64_peak, "unit": "TFLOP/s",
            "frac": achieved / f64_peak if achieved else None,
            "peak_source": "FP64 DMMA rate probed in this run (mb200_probe_peaks); MEASURED_PEAKS.json has no FP64 entry"
                           if peaks.get("fp64_dmma_tflops", 0) > 0 else "fallback: profiles/r01_fp64_peaks.log",
            "algorithmic_flop_per_atomq": FLOP_PER_ATOMQ, **common}
    else:
        achieved = I8_OPS_PER_ATOMQ * atomq_per_launch * kern_launches / kern_s / 1e12 if kern_s > 0 else None
        roofline = {
            "kernel": "k_sf_contract_tc", "bound": "tensor", "pipe": "tcgen05.mma kind::i8 (UTCIMMA), int32 accumulators in TMEM",
            "achieved": achieved, "peak": i8_peak, "unit": "TFLOP/s",
            "unit_note": "int8 tensor ops (multiply-add = 2), i.e. TOP/s; MEASURED_PEAKS.json carries only the bf16 figure "
                         f"({peaks_file.get('bf16_tflops')} TFLOP/s; 2x that = {2 * (peaks_file.get('bf16_tflops') or 0):.0f} would be "
                         "the int8 rate if the pipe kept its nominal 2:1 ratio), the int8 ceiling is probed on the same pipe",
            "frac": achieved / i8_peak if achieved else None,
            "frac_vs_2x_bf16_burst": achieved / (2 * peaks_file["bf16_tflops"]) if achieved and peaks_file.get("bf16_tflops") else None,
            "peak_source": "tcgen05 kind::i8 dense rate probed in this run (mb200_probe_peaks: 148 CTAs x 128x256x32 MMAs from "
                           "shared memory); nominal 4500" if peaks.get("i8_tops", 0) > 0 else "fallback: profiles/r01_umma_i8.log",
            "algorithmic_int8_ops_per_atomq": I8_OPS_PER_ATOMQ,
            "shared_pipe_bound": shared_pipe_bound(mma_count, kern_s, wl, peaks, i8_peak),
            "algorithmic_unit_note": "SURVEY 8(d) counts 8 FP64 flop per atom*q; fp64_equivalent_tflops restates the kernel in "
                                     "that unit against the FP64 DMMA ceiling (a ratio > 1 is what the split-precision path buys)",
            **common}
    line = {
        "metric": "saxs_atom_q_evals_per_s", "value": value, "unit": "atom*q/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i8x5->s32 (40-bit fixed point, FP64 tables and recombination)" if args.precision == "i8x5" else "f64",
        "data": "synthetic", "config": config_dict(args, wl, world),
        "batch": {"frames_per_step": B, "frames_total": K * B * world, "precision": PRECISION_TEXT[args.precision]},
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
        line["path2_profile_hist"] = path2_summary(ctx, lib, torch, peaks)
    if configs is not None:
        line["configs"] = configs
    print(json.dumps(line))
    parallel.shutdown()


# ----------------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configurations: bounded runs through the public classes, host memory -> results, with parity
# ----------------------------------------------------------------------------------------------------------------------
def run_configs(args, ctx, lib, torch, rank, world, peaks):
    import maicos_b200 as mb
    from maicos_b200 import _cabi, parallel
    from maicos_b200.lib.math import sfactor_abs_error_bound, structure_factor
    from maicos_b200.mdshim import MemoryTrajectory, Universe
    from maicos_b200.synthetic import confined_slab_frames, cubic_box_for, spce_topology, spce_water_frames
    from oracle import kernel, restatement

    # N > 1: the two configurations that shard (C3 over frames, C5 over q-blocks); the single-GPU ones belong to N = 1
    want = [c.strip().upper() for c in args.configs.split(",") if c.strip()] or (
        ["C1", "C3", "C4", "C5", "T1M"] if world == 1 else ["C3", "C5"])
    kernel.set_threads()
    cores = kernel.n_threads()
    stream = torch.cuda.ExternalStream(int(lib.mb200_ctx_stream(ctx.handle)))
    out = {"note": "bounded runs of the other BASELINE.json configurations through the public classes (.run() from host "
                   "memory: staging, H2D, kernels, D2H inside the timed region); best of two passes after one warm-up pass; "
                   "parity = the same inputs through the oracle (reference Cython kernel / NumPy restatement) on "
                   f"{cores} host cores", "precision": args.precision}
    i8_peak = peaks["i8_tops"] if peaks.get("i8_tops", 0) > 0 else I8_PEAK_TOPS_FALLBACK

    def max_over_ranks(x):
        if parallel.world() <= 1:  # also inside parallel.local_only()
            return float(x)
        tt = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        return float(tt.item())

    def timed(fn, passes=2):
        """(result, wall seconds, device milliseconds between two events on the library's stream), best wall of `passes`."""
        best = None
        for _ in range(passes):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            parallel.barrier(); torch.cuda.synchronize()
            e0.record(stream)
            t0 = time.perf_counter()
            r = fn()
            e1.record(stream)
            torch.cuda.synchronize(); parallel.barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            dev = max_over_ranks(e0.elapsed_time(e1))
            if best is None or dt < best[1]:
                best = (r, dt, dev)
        return best

    def water_universe(frames, dims, n_mol, pinned=False):
        dims = np.asarray(dims, dtype=np.float32)
        return Universe(3 * n_mol, MemoryTrajectory(frames, dims, pinned=pinned), **spce_topology(n_mol))

    def slab_parity(pos, box, w, qmax, frac=0.05):
        """Raw S cube of the CUDA kernel against the oracle's C port on the h-slab [0, h1) that holds >= `frac` of the
        accepted q vectors (full-size parity on a slab); every cell of the slab is compared."""
        box = np.double(box)
        q_gpu, s_gpu = structure_factor(np.double(pos), box, 0.0, qmax, 0.0, np.pi, w)
        per_h = np.count_nonzero(q_gpu.reshape(q_gpu.shape[0], -1), axis=1)
        h1 = int(np.searchsorted(np.cumsum(per_h), frac * per_h.sum()) + 1)
        t0 = time.perf_counter()
        q_ref, s_ref = kernel.port_structure_factor_slab(np.double(pos), box, 0.0, qmax, 0.0, np.pi, w, 0, h1)
        dt = time.perf_counter() - t0
        nz = s_ref[:h1] != 0
        rel = np.abs(s_gpu[:h1][nz] / s_ref[:h1][nz] - 1)
        bound = sfactor_abs_error_bound(len(pos), float(np.max(np.abs(w))) if len(w) else 1.0)
        return {"cells_compared": int(nz.sum()), "fraction_of_q": float(nz.sum() / max(per_h.sum(), 1)), "h_slab": [0, h1],
                "q_values_equal": bool(np.array_equal(q_gpu[:h1], q_ref[:h1])),
                "mask_equal": bool(np.array_equal(s_gpu[:h1] != 0, s_ref[:h1] != 0)),
                "max_rel_err_S": float(rel.max()) if rel.size else 0.0, "median_rel_err_S": float(np.median(rel)) if rel.size else 0.0,
                "max_abs_err_amplitude": float(np.max(np.abs(np.sqrt(s_gpu[:h1][nz]) - np.sqrt(s_ref[:h1][nz])))) if rel.size else 0.0,
                "abs_bound_amplitude": bound, "oracle": "oracle/sfactor_port.c (C restatement, pinned on the reference kernel)",
                "oracle_seconds": dt, "oracle_atomq_per_s": float(nz.sum()) * len(pos) / dt, "tolerance": "rel <= 1e-6 per cell"}

    # ---- C1: Saxs on the shipped water box (configs[0]) ---------------------------------------------------------
    def c1():
        g = np.load(ROOT / "tests/golden/water_trr_frames.npz")
        el = np.load(ROOT / "tests/golden/water_gro.npz")["elements"].astype(object)
        n_rep = 25
        pos, dims = np.tile(g["positions"], (n_rep, 1, 1)), np.tile(g["dimensions"], (n_rep, 1))
        n_mol = pos.shape[1] // 3
        topo = dict(spce_topology(n_mol)); topo["elements"] = el
        u = Universe(pos.shape[1], MemoryTrajectory(pos, dims), **topo)
        mb.Saxs(u.atoms).run(stop=8)
        s, dt, dev = timed(lambda: mb.Saxs(u.atoms).run())
        nq = [n_valid_q(d[:3], 6.0) for d in g["dimensions"]]
        atomq = float(sum(nq)) * n_rep * pos.shape[1]
        # parity: EVERY cell of the raw cube of every distinct frame and group, then the class result
        worst, mask_ok = 0.0, True
        for f in range(len(g["positions"])):
            box = np.double(g["dimensions"][f][:3])
            for e in np.unique(el.astype(str)):
                p = np.double(g["positions"][f][el.astype(str) == e])
                q1, s1 = kernel.structure_factor(p, box, 0, 6.0, 0, np.pi, np.ones(len(p)))
                q2, s2 = structure_factor(p, box, 0, 6.0, 0, np.pi, np.ones(len(p)))
                mask_ok &= bool(np.array_equal(q1, q2) and np.array_equal(s1 != 0, s2 != 0))
                worst = max(worst, float(np.max(np.abs(s2[s1 != 0] / s1[s1 != 0] - 1))))
        ref, stats, _ = restatement.run_saxs(g["positions"], g["dimensions"][:, :3], el.astype(str))
        got = mb.Saxs(Universe(pos.shape[1], MemoryTrajectory(g["positions"], g["dimensions"]), **topo).atoms).run()
        out["C1_saxs_shipped_water"] = {
            "workload": "maicos.Saxs on the shipped 1530-atom SPC/E box, default q range (qmax 6, dq 0.1): the 4 golden NPT "
                        "frames of examples/water.trr (tests/golden) tiled to 100 frames",
            "frames": int(s.n_frames), "n_atoms": int(pos.shape[1]), "n_q_valid": nq, "e2e_seconds": dt, "device_ms": dev,
            "frames_per_s": s.n_frames / dt, "atomq_per_s": atomq / dt,
            "parity": {"checked": "all cells of the raw S cube of all 4 frames x 2 element groups vs the reference kernel, and "
                                  "Saxs(...).run() on the 4 frames vs the NumPy restatement",
                       "cells_mask_and_q_equal": mask_ok, "max_rel_err_S_cell": worst,
                       "max_rel_err_I_binned": float(np.max(np.abs(got.results.scattering_intensities / ref["scattering_intensities"] - 1))),
                       "bincount_equal": bool(np.array_equal(got.sums.bincount, stats.sums["bincount"])),
                       "max_rel_err_dI": float(np.max(np.abs(got.results.dscattering_intensities / ref["dscattering_intensities"] - 1))),
                       "tolerance": "rel <= 1e-6, counts exact"}}

    # ---- C3: DiporderStructureFactor on 300k-atom water (configs[2]) ---------------------------------------------
    def c3():
        n_mol = 100_000
        L = cubic_box_for(n_mol)
        per2 = 32 if world == 1 else 16
        F = per2 * world  # frame-sharded over the ranks: per2 frames per rank at qmax 2, 4 per rank at qmax 6
        frames = spce_water_frames(n_mol, L, 20261020, n_frames=F)
        u = water_universe(frames, [L, L, L, 90, 90, 90], n_mol, pinned=True)
        dims = np.float32([L, L, L])
        for qmax, per_rank in ((2.0, per2), (6.0, 4)):
            nq = n_valid_q(dims, qmax)
            run = lambda: mb.DiporderStructureFactor(u.atoms, qmax=qmax, dq=0.01).run(stop=per_rank * world)  # noqa: E731
            run()
            a, dt, dev = timed(run)
            entry = {
                "workload": f"DiporderStructureFactor on synthetic 300k-atom SPC/E water (100k molecules, L={L:.2f} A), "
                            f"qmax={qmax:g}/A, dq=0.01/A, unwrap + pack + centre of mass + mu/|mu| on the device, "
                            f"{per_rank} frames per GPU, frames sharded over {world} GPU(s)",
                "frames": int(a.n_frames), "n_compounds": n_mol, "n_q_valid": nq, "maxn": [int(x) for x in kernel.port_maxn(np.double(dims), qmax)],
                "e2e_seconds": dt, "device_ms": dev, "frames_per_s": a.n_frames / dt,
                "atomq_per_s": 3.0 * n_mol * nq * a.n_frames / dt, "device_atomq_per_s": 3.0 * n_mol * nq * a.n_frames / (dev * 1e-3),
                "frac_of_i8_peak_device": I8_OPS_PER_ATOMQ * 3.0 * n_mol * nq * a.n_frames / world / (dev * 1e-3) / 1e12 / i8_peak,
                "host_timings_s": a.timings}
            if rank == 0:
                centres, unit = restatement.compound_prologue(frames[0], dims, np.arange(0, 3 * n_mol + 1, 3), u.atoms.masses,
                                                              u.atoms.charges, "com", True, True, weight_mode=4)
                if qmax == 2.0:  # one whole frame through the reference kernel
                    t0 = time.perf_counter()
                    obs, _ = restatement.dipsf_single_frame(centres, unit, np.double(dims), 0, qmax, a.n_bins)
                    t_ref = time.perf_counter() - t0
                    one = mb.DiporderStructureFactor(water_universe(frames[:1], [L, L, L, 90, 90, 90], n_mol).atoms, qmax=qmax,
                                                     dq=0.01)
                    with parallel.local_only():
                        one.run()
                    nz = obs["structure_factors"] != 0
                    entry["parity"] = {
                        "checked": "frame 0, all three components, every |q| bin: class result vs oracle compound prologue + "
                                   "reference Cython kernel + NumPy body",
                        "max_rel_err_S_binned": float(np.max(np.abs(one.means.structure_factors[nz] / obs["structure_factors"][nz] - 1))),
                        "nonzero_bins_equal": bool(np.array_equal(one.means.structure_factors != 0, nz)),
                        "tolerance": "rel <= 1e-6", "oracle_seconds": t_ref,
                        "oracle_atomq_per_s": 3.0 * n_mol * nq / t_ref, "oracle_cores": cores}
                    entry["vs_cpu_reference_e2e"] = entry["atomq_per_s"] / entry["parity"]["oracle_atomq_per_s"]
                else:            # 1.39M q vectors: an h-slab with >= 5 % of them, x component of mu/|mu|
                    entry["parity"] = slab_parity(centres, dims, np.ascontiguousarray(unit[0]), qmax)
            out[f"C3_dipsf_300k_qmax{qmax:g}"] = entry
        del frames, u

    # ---- C4: DensityPlanar + DiporderPlanar on a 1M-atom graphene-confined slab (configs[3]) --------------------
    def c4():
        n_c, n_w, F = 100_000, 300_000, 96
        box = [215.0, 215.0, 215.4]
        fr, topo = confined_slab_frames(n_w, n_c, *box, 200.0, 20261021, n_frames=F)
        u_pg = Universe(fr.shape[1], MemoryTrajectory(fr, np.float32([*box, 90, 90, 90])), **topo)
        u = Universe(fr.shape[1], MemoryTrajectory(fr, np.float32([*box, 90, 90, 90]), pinned=True), **topo)
        run_d = lambda: mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False).run()  # noqa: E731
        run_pg = lambda: mb.DensityPlanar(u_pg.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False).run()  # noqa: E731
        run_pg()
        a_pg, dt_pg, dev_pg = timed(run_pg)
        run_d()
        a, dt, dev = timed(run_d, passes=3)  # PCIe-bound: the boxes show occasional 2-4x slower runs (profiles/r02_notes.md)
        # steady state: the same run over the first third of the frames -- the slope removes the per-run constants
        # (weights upload, pipeline fill with a quarter and a half batch, conclude), which a 10k-frame run does not see
        a_s, dt_s, _ = timed(lambda: mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False).run(stop=F // 3))
        zmin, zmax, geo = restatement.planar_geometry(np.float32(box), 2, a.n_bins)
        ref_c = restatement.hist_planar(fr[0], None, 2, a.n_bins, zmin, zmax)
        ref_w = restatement.hist_planar(fr[0], topo["masses"], 2, a.n_bins, zmin, zmax)
        one = mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False).run(stop=1)
        nzw = ref_w != 0
        out["C4_densityplanar_1M"] = {
            "workload": "DensityPlanar (mass) on a synthetic 1M-atom graphene-confined SPC/E slab, dz=0.01 nm (bin_width 0.1 A, "
                        f"{a.n_bins} bins), {F} frames",
            "frames": int(a.n_frames), "n_atoms": int(fr.shape[1]), "n_bins": int(a.n_bins), "e2e_seconds": dt, "device_ms": dev,
            "frames_per_s": a.n_frames / dt, "atoms_per_s": a.n_frames * fr.shape[1] / dt,
            "steady_state_frames_per_s": (a.n_frames - a_s.n_frames) / max(dt - dt_s, 1e-9),
            "pcie_bound_frames_per_s": 50e9 / (fr.shape[1] * 12), "host_timings_s": a.timings,
            "input": "trajectory in page-locked host memory (MemoryTrajectory(pinned=True)): no staging copy, H2D inside the timed region",
            "pageable_trajectory": {"frames_per_s": a_pg.n_frames / dt_pg, "e2e_seconds": dt_pg, "host_timings_s": a_pg.timings},
            "parity": {"checked": "frame 0 through the class vs np.histogram (core/base.py:815-818 restated): counts and "
                                  "mass-weighted sums / bin volume",
                       "bincount_equal": bool(np.array_equal(one.sums.bincount, ref_c)),
                       "max_rel_err_profile": float(np.max(np.abs(one.means.profile[nzw] * geo["bin_volume"][nzw] / ref_w[nzw] - 1))),
                       "tolerance": "counts exact, weighted sums rel <= 1e-12"}}
        # the reference's default arguments (pack=True: `universe.atoms.wrap()` on the host in every frame, base.py:446-448):
        # the float32 wrap runs on the device in front of the histogram
        run_dp = lambda: mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1).run()  # noqa: E731
        run_dp()
        a_dp, dt_dp, _ = timed(run_dp, passes=3)
        wr0 = restatement.wrap_atoms_f32(fr[0], np.float32(box))
        one_dp = mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1).run(stop=1)
        out["C4_densityplanar_1M"]["default_arguments_pack_true"] = {
            "frames_per_s": a_dp.n_frames / dt_dp, "e2e_seconds": dt_dp, "host_timings_s": a_dp.timings,
            "parity": {"checked": "frame 0 vs np.histogram of the host-wrapped frame (x - floor(x / L) L in float32)",
                       "bincount_equal": bool(np.array_equal(one_dp.sums.bincount, restatement.hist_planar(wr0, None, 2, a.n_bins, zmin, zmax)))}}
        # C4 as BASELINE.json words it -- DensityPlanar + DiporderPlanar over the same trajectory -- in ONE frame loop
        # (AnalysisCollection, core/base.py:602-718): the members (the whole system, and the water as a window of the whole
        # frames) read the frames where the page-locked trajectory holds them and share one host -> device copy per batch
        water = u.select_atoms("element O H")

        def members():
            return [mb.DensityPlanar(u.atoms, dens="mass", bin_width=0.1, unwrap=False, pack=False),
                    mb.DensityPlanar(water, dens="charge", bin_width=0.1, unwrap=False, pack=False),
                    mb.DiporderPlanar(water, bin_width=0.1, unwrap=False, pack=False)]

        def pstat():
            o = np.zeros(3, dtype=np.int64)
            _cabi.check(lib.mb200_ctx_prefetch_stats(ctx.handle, _cabi.ptr(o)))
            return o

        mb.AnalysisCollection(*members()).run(stop=F // 3)
        p0 = pstat()
        coll, dt_c, _ = timed(lambda: mb.AnalysisCollection(*members()).run(), passes=3)
        p1 = pstat()
        solo, dt_solo, _ = timed(lambda: [m.run() for m in members()], passes=3)
        out["C4_collection_density_diporder_1M"] = {
            "workload": "AnalysisCollection(DensityPlanar mass of all 1M atoms, DensityPlanar charge of the water, DiporderPlanar of "
                        f"the water), {F} frames, one frame loop",
            "frames_per_s": F / dt_c, "e2e_seconds": dt_c, "three_solo_runs_seconds": dt_solo,
            "speedup_vs_solo_runs": dt_solo / dt_c,
            "h2d_copies_over_the_three_timed_passes": {"issued": int(p1[0] - p0[0]), "saved_by_sharing": int(p1[1] - p0[1]),
                                                       "not_prefetched": int(p1[2] - p0[2])},
            "parity": {"checked": "every member of the collection vs its solo run (means.profile)",
                       "bitwise_equal": bool(all(np.array_equal(a.means.profile, b.means.profile, equal_nan=True)
                                                 for a, b in zip(coll._analysis_instances, solo)))}}
        run_p = lambda: mb.DiporderPlanar(water, bin_width=0.1, unwrap=False, pack=False).run()  # noqa: E731
        run_p()
        a, dt, dev = timed(run_p, passes=3)
        off = n_c
        c_ref, w_ref = restatement.compound_prologue(fr[0][off:], np.float32(box), np.arange(0, 3 * n_w + 1, 3),
                                                     topo["masses"][off:], topo["charges"][off:], "com", False, False,
                                                     weight_mode=1, pdim=2)
        ref_c = restatement.hist_planar(c_ref, None, 2, a.n_bins, zmin, zmax)
        ref_w = restatement.hist_planar(c_ref, w_ref, 2, a.n_bins, zmin, zmax)
        one = mb.DiporderPlanar(water, bin_width=0.1, unwrap=False, pack=False).run(stop=1)
        scale = np.max(np.abs(ref_w))
        out["C4_diporderplanar_300k_mol"] = {
            "workload": "DiporderPlanar (P0, pdim z) on the water of the same slab: 300k molecules, centre of mass + dipole per "
                        f"molecule formed on the device, {a.n_bins} bins, {F} frames",
            "frames": int(a.n_frames), "n_compounds": n_w, "n_bins": int(a.n_bins), "e2e_seconds": dt, "device_ms": dev,
            "frames_per_s": a.n_frames / dt, "host_timings_s": a.timings,
            "parity": {"checked": "frame 0 through the class vs oracle compound prologue + np.histogram",
                       "bincount_equal": bool(np.array_equal(one.sums.bincount, ref_c)),
                       "max_abs_err_profile_rel_to_max": float(np.max(np.abs(one.means.profile * geo["bin_volume"] - ref_w)) / scale),
                       "tolerance": "counts exact, dipole sums <= 1e-10 of the largest bin (float32 coordinates in, float64 sums)"}}
        del fr, u

    # ---- C5: Saxs on one 1M-atom frame, qmax 3 (configs[4]); q-blocks over the ranks ----------------------------
    def c5():
        n_mol = 333_333
        L = cubic_box_for(n_mol)
        frames = spce_water_frames(n_mol, L, 20261022, n_frames=1)  # the same frame on every rank
        dims = np.float32([L, L, L])
        nq = n_valid_q(dims, 3.0)
        u = water_universe(frames, [L, L, L, 90, 90, 90], n_mol, pinned=True)
        run5 = lambda: mb.Saxs(u.atoms, qmax=3.0).run()  # noqa: E731
        with parallel.local_only():            # every rank on its own: the unsharded time and result
            run5()
            a1, dt1, dev1 = timed(run5)
        entry = {
            "workload": f"Saxs on a single synthetic 1M-atom SPC/E frame (333333 molecules, L={L:.2f} A), qmax=3.0/A (30/nm), "
                        "q-vector blocks sharded over the GPUs",
            "n_atoms": 3 * n_mol, "n_q_valid": nq, "maxn": [int(x) for x in kernel.port_maxn(np.double(dims), 3.0)], "n_gpus": world}
        if world == 1:
            entry.update({"e2e_seconds": dt1, "device_ms": dev1, "atomq_per_s": 3.0 * n_mol * nq / dt1,
                          "device_atomq_per_s": 3.0 * n_mol * nq / (dev1 * 1e-3),
                          "frac_of_i8_peak_device": I8_OPS_PER_ATOMQ * 3.0 * n_mol * nq / (dev1 * 1e-3) / 1e12 / i8_peak,
                          "host_timings_s": a1.timings})
        else:
            parallel.set_mode("qblocks")
            try:
                run5()
                aq, dtq, devq = timed(run5, passes=3)
            finally:
                parallel.set_mode("frames")
            nzq = a1.results.scattering_intensities != 0
            entry.update({
                "scaling": "strong (one frame, q-blocks)", "e2e_seconds": dtq, "device_ms": devq,
                "atomq_per_s": 3.0 * n_mol * nq / dtq, "device_atomq_per_s": 3.0 * n_mol * nq / (devq * 1e-3),
                "host_timings_s": aq.timings,
                "unsharded_on_one_gpu_of_this_run": {"e2e_seconds": dt1, "device_ms": dev1},
                "speedup_e2e": dt1 / dtq, "speedup_device": dev1 / devq,
                "efficiency_e2e": dt1 / dtq / world, "efficiency_device": dev1 / devq / world,
                "sharded_vs_unsharded": {
                    "max_rel_diff_I": float(np.max(np.abs(aq.results.scattering_intensities[nzq] / a1.results.scattering_intensities[nzq] - 1))),
                    "bincount_equal": bool(np.array_equal(aq.sums.bincount, a1.sums.bincount))}})
        if rank == 0:
            el = spce_topology(n_mol)["elements"].astype(str)
            p = frames[0][el == "O"]
            p = p - dims * np.round(p / dims)  # saxs.py:193-195
            entry["parity"] = slab_parity(p, dims, np.ones(len(p)), 3.0)
            entry["parity"]["checked"] = ("oxygen group (333333 atoms) of the frame: raw S cube of the CUDA kernel vs the oracle on "
                                          "an h-slab with >= 5 % of the 582k accepted q vectors, every cell")
        out["C5_saxs_1M_single_frame_qmax3"] = entry
        del frames, u

    # ---- north_star's quoted target: Saxs on 1M atoms at qmax 20/nm, vs the host-CPU reference ------------------
    def t1m():
        n_mol = 333_333
        L = cubic_box_for(n_mol)
        F = 4
        frames = spce_water_frames(n_mol, L, 20261023, n_frames=F)
        dims = np.float32([L, L, L])
        nq = n_valid_q(dims, 2.0)
        u = water_universe(frames, [L, L, L, 90, 90, 90], n_mol, pinned=True)
        runt = lambda: mb.Saxs(u.atoms, qmax=2.0).run()  # noqa: E731
        runt()
        a, dt, dev = timed(runt)
        # CPU reference on a bounded sample: the first 16 667 molecules of frame 0 in the SAME box (the kernel's cost is
        # linear in the atom count), through the Cython kernel + NumPy body; the same sub-system through the class
        m_sub = 16_667
        el_sub = spce_topology(m_sub)["elements"].astype(str)
        sub = frames[:1, :3 * m_sub]
        t0 = time.perf_counter()
        ref, stats, _ = restatement.run_saxs(sub, dims[None], el_sub, qmax=2.0, dq=0.1)
        t_ref = time.perf_counter() - t0
        got = mb.Saxs(water_universe(sub, [L, L, L, 90, 90, 90], m_sub).atoms, qmax=2.0).run()
        cpu_rate = 3.0 * m_sub * nq / t_ref
        out["T1M_saxs_1M_qmax2"] = {
            "workload": f"north_star target: Saxs on a synthetic 1M-atom SPC/E box (L={L:.2f} A), qmax=2.0/A (20/nm), {F} frames, 1 GPU",
            "frames": F, "n_atoms": 3 * n_mol, "n_q_valid": nq, "maxn": [int(x) for x in kernel.port_maxn(np.double(dims), 2.0)],
            "e2e_seconds": dt, "device_ms": dev, "frames_per_s": F / dt, "atomq_per_s": 3.0 * n_mol * nq * F / dt,
            "device_atomq_per_s": 3.0 * n_mol * nq * F / (dev * 1e-3),
            "frac_of_i8_peak_device": I8_OPS_PER_ATOMQ * 3.0 * n_mol * nq * F / (dev * 1e-3) / 1e12 / i8_peak,
            "cpu_reference": {"value": cpu_rate, "unit": "atom*q/s", "cores": cores, "kind": kernel.kind(),
                              "sample": f"{3 * m_sub} atoms of frame 0 in the same box and q range ({3.0 * m_sub * nq:.2e} atom*q) in {t_ref:.1f} s"},
            "speedup_vs_cpu_reference_per_gpu": 3.0 * n_mol * nq * F / dt / cpu_rate, "target": ">= 100x per GPU",
            "parity": {"checked": "the CPU sample's sub-system through Saxs(...).run() vs the reference kernel + NumPy body",
                       "max_rel_err_I_binned": float(np.max(np.abs(got.results.scattering_intensities / ref["scattering_intensities"] - 1))),
                       "bincount_equal": bool(np.array_equal(got.sums.bincount, stats.sums["bincount"])),
                       "tolerance": "rel <= 1e-6, counts exact"}}
    for name, fn, sharded in (("C1", c1, False), ("C3", c3, True), ("C4", c4, False), ("C5", c5, True), ("T1M", t1m, False)):
        if name not in want:
            continue
        if sharded:
            fn()                    # every rank takes part
        elif rank == 0:
            with parallel.local_only():  # a single-process job while the other ranks wait at the next barrier
                fn()
    return out


def shared_pipe_bound(mma_count, kern_s, wl, peaks, i8_peak):
    """The bound that applies to k_sf_contract_tc: on a B200 SM the FP64 pipe (the generators that form the A operand) and the
    tensor pipe do not run at the same time (tools/microbench/fp64_vs_umma.cu, profiles/r02_fp64_vs_umma.log: combined time
    = sum of the two), so a K-step of 16 atoms costs at least its MMA time plus its FP64 time."""
    if not mma_count or kern_s <= 0:
        return None
    k_steps = mma_count / 15.0                      # fifteen digit products per K-step
    np_cols = 2 * ((wl["maxn"][2] + 7) // 8 * 8)    # accumulator columns per digit group (re / im of the l tile)
    int8_ops = mma_count * 2.0 * 128 * np_cols * 32  # issued, dead cells of the tiles included
    fp64_lane_instr = k_steps * 8 * 64 * 32          # 8 generator tasks x 64 FP64 warp instructions per K-step
    f64 = peaks.get("fp64_fma_tflops") or 33.0
    t_tensor = int8_ops / (i8_peak * 1e12)
    t_fp64 = fp64_lane_instr * 2.0 / (f64 * 1e12)
    return {"note": "FP64 FMA and tcgen05 MMA work of an SM serialise (measured: profiles/r02_fp64_vs_umma.log); bound = issued int8 "
                    "ops / probed kind::i8 rate + generator FP64 instructions / probed FP64 FMA rate",
            "issued_int8_ops_total": int8_ops, "generator_fp64_flop_equiv_total": fp64_lane_instr * 2.0,
            "fp64_fma_tflops_probed": peaks.get("fp64_fma_tflops"),
            "tensor_seconds": t_tensor, "fp64_seconds": t_fp64, "measured_seconds": kern_s,
            "frac_of_serialised_bound": (t_tensor + t_fp64) / kern_s}


def path2_summary(ctx, lib, torch, peaks=None):
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
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = pk.get("hbm_gbs", 6650.0)
    gbs = 12.0 * n * F / sec / 1e9
    return {"workload": f"weighted + count slab histogram, {n} atoms x {F} frames resident, {nb} bins",
            "frames_per_s": F / sec, "atoms_per_s": n * F / sec,
            "roofline": {"kernel": "k_hist_chunk", "bound": "hbm", "achieved": 12.0 * n * F / ksec / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": 12.0 * n * F / ksec / 1e9 / peak, "kernel_ms_per_launch": ksec * 1e3,
                         "whole_call": {"achieved": gbs, "frac": gbs / peak, "ms_per_call": sec * 1e3,
                                        "note": "call = parameter upload + max|w| + histogram + merge kernels + per-frame results to host"},
                         "co_limiter": "four non-returning shared-memory atomics per element at ~9.7 lanes per clock and SM (profiles/r02_smem_atomics.log)",
                         "algorithmic_bytes_per_atom_frame": 12,
                         "hbm_copy_gbs_probed_in_this_run": (peaks or {}).get("hbm_copy_gbs"),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if pk else "fallback 6650 GB/s"}}


def cpu_baseline(wl):
    """Reference CPU path on the host cores, bounded sample (about 10-30 s of CPU work)."""
    from maicos_b200.synthetic import spce_topology, spce_water_frames
    from oracle import kernel, restatement

    cores = kernel.set_threads()
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
    return {"value": done * wl["atomq_per_frame"] / dt, "unit": "atom*q/s", "cores": cores,
            "kind": kernel.kind(), "frames_per_s": done / dt,
            "sample": f"{done} full frame(s) of the workload (O + H groups, {wl['atomq_per_frame']:.3e} atom*q each) "
                      f"through the Cython kernel + NumPy Saxs frame body in {dt:.1f} s on {cores} threads"}, first_obs, frames[0]


if __name__ == "__main__":
    main()
