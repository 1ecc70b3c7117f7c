#!/usr/bin/env python
"""Benchmark of the embedding-space hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline line (ONE JSON line on stdout, rank 0): OCELoss fwd+bwd px/s on BASELINE
config #2 -- offsets (8, 2, 496, 496) fp32, 8 x 702,367 (anchor, reference) pairs,
int64 coordinate lists as the reference's DataLoader delivers them, kappa 10,
density 0.1, T 10, w 1e-5.  A "step" is one pass of the fused gather + loss +
backward over one such batch; with N GPUs every rank runs its own batch (the path
shards by batch, no data-path collective: weak scaling).  The same line carries

  roofline      achieved algorithmic HBM GB/s of the fused kernel vs the measured peak
  cpu_baseline  the oracle port of the reference path timed on this box's host cores
  e2e           the same metric through the public API from pinned HOST buffers
  multi_gpu     (every N) wall-clock jobs that exercise the multi-GPU inference paths: ONE volume split by
                seed over the ranks (two all-gathers, N-rank result asserted equal to the 1-rank result) and
                BASELINE configs[4] blockwise inference (scan blocks dealt round-robin)
  detect        second half of BASELINE's metric, LAST in the line: mean-shift detection Mpx/s on
                BASELINE configs[2] (128 x 256 x 256 volume, 3-D embeddings) with its own roofline
                (distance tests of the hill-climb kernel vs the measured FP64 FMA peak), cpu_baseline, e2e
                and a label-for-label comparison with the reference's own output (tests/golden/)

`--impl reference` times the oracle port (torch CPU, all host threads) on rank 0; at N = 1 it also runs the
reference detect path (oracle port of mean_shift_segmentation, scikit-learn underneath) ONCE on the full
configs[2] volume (~4 min).  Nothing here reads /root/reference.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# ---- BASELINE config #2 ---------------------------------------------
B, D, OUT = 8, 2, (496, 496)
KAPPA, DENSITY, TEMP, REGW = 10.0, 0.1, 10.0, 1e-5
N_ANCHORS = int(DENSITY * (OUT[0] - 2 * KAPPA) * (OUT[1] - 2 * KAPPA))  # 22 657
N_REFS = int(DENSITY * KAPPA**2 * np.pi)  # 31
P = N_ANCHORS * N_REFS  # 702 367
N_PX = B * OUT[0] * OUT[1]  # 1 968 128
# algorithmic bytes per step: both int64 coordinate lists once, offsets once, dense gradient once
ALGO_BYTES = B * P * D * 8 * 2 + N_PX * D * 4 + N_PX * D * 4


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the
    same kernel (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep), or None."""
    try:
        rec = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))[kernel_key]
        return int(rec["dram_bytes_read"] + rec["dram_bytes_write"])
    except Exception:
        return None

WORKLOAD = "configs[1]: OCELoss fwd+bwd, offsets (8,2,496,496) f32, 8x702367 pairs, int64 coords, kappa=10, density=0.1"

# ---- BASELINE config #3 (secondary: detect) ----------------------------------------------
DET_SHAPE, DET_OBJECTS, DET_RADIUS, DET_BW, DET_THR, DET_RP = (128, 256, 256), 400, 10.0, 7.0, 0.5, 0.1


def measured_peak_gbs():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not any(a - 0.05 <= ts <= b + 0.15 for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # timed regions shorter than the sampling period: fall back to every sample taken
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    smax.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- reference arm
def cpu_loss_step_time(batch, steps, warmup, seed=0):
    """Oracle port of the reference loss slice on the host CPU (torch, all threads)."""
    from cellulus_b200 import synthetic
    from oracle import oce_loss as oloss
    from oracle import sampler as osampler

    np.random.seed(seed)
    pairs = [osampler.sample_coordinates(OUT, KAPPA, DENSITY, D) for _ in range(batch)]
    anchors = torch.from_numpy(np.stack([p[0] for p in pairs])).long()
    refs = torch.from_numpy(np.stack([p[1] for p in pairs])).long()
    offsets = torch.from_numpy(synthetic.loss_offsets(batch, D, OUT, seed=seed))
    for _ in range(warmup):
        oloss.loss_step(offsets, anchors, refs, TEMP, REGW)
    t0 = time.perf_counter()
    for _ in range(steps):
        oloss.loss_step(offsets, anchors, refs, TEMP, REGW)
    return (time.perf_counter() - t0) / max(steps, 1)


def cpu_detect(shape, objects, what):
    """Oracle port of `mean_shift_segmentation` (cellulus/utils/mean_shift.py:6-45; scikit-learn MeanShift, hill
    climb on ONE core as shipped, `predict` on all cores) on a scene of the bench family."""
    from cellulus_b200 import synthetic
    from oracle import mean_shift as oms

    emb, _, _ = synthetic.blob_scene(shape, objects, radius=DET_RADIUS, seed=0)
    emb64 = emb.astype(np.float64)
    np.random.seed(0)
    t0 = time.perf_counter()
    labels = oms.mean_shift_segmentation(emb64[np.newaxis, :3].copy(), emb64[3], DET_BW, 0, DET_RP, DET_THR, None)
    dt = time.perf_counter() - t0
    n = int(np.prod(shape))
    return {"value": n / dt / 1e6, "unit": "Mpx/s", "cores": 1, "kind": "port", "seconds": dt,
            "centres": int(labels.max()),
            "sample": f"{what}: {shape[0]}x{shape[1]}x{shape[2]}, {objects} balls, {int((labels > 0).sum())} fg voxels, "
                      f"{dt:.1f} s; sklearn hill climb is single-core (n_jobs=None), predict uses OpenMP"}, labels


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 per rank; the reference arm is allowed every host core
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cores = torch.get_num_threads()
    t_one = cpu_loss_step_time(1, 1, 1)
    budget = 100.0 / max(args.steps + args.warmup, 1)  # whole run within a couple of minutes
    batch = int(max(1, min(B, budget / max(t_one, 1e-6))))
    t = cpu_loss_step_time(batch, args.steps, args.warmup)
    px = batch * OUT[0] * OUT[1]
    value = px / t
    sample = f"{batch}/{B} samples of the batch per step ({batch * P} pairs), {args.steps} steps"
    detect = None
    if args.gpus == 1 and not args.skip_detect:
        # the reference detect path on the FULL configs[2] volume, once (a few minutes of one host core)
        detect, labels = cpu_detect(DET_SHAPE, DET_OBJECTS, "full configs[2] volume")
        detect.update({"metric": "detect Mpx/s (mean-shift)", "impl": "reference",
                       "labels_equal_committed_golden": golden_labels_equal(labels)})
    line = {
        "impl": "reference", "metric": "OCELoss fwd+bwd px/s", "value": value, "unit": "px/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "px/s", "cores": cores, "kind": "port", "sample": sample,
                         "pairs_per_s": batch * P / t},
        "e2e": {"value": value, "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "detect": detect,
    }
    print(json.dumps(line), flush=True)


def golden_labels_equal(labels):
    """Label-for-label comparison with the REFERENCE's own output on the configs[2] bench volume
    (tests/golden/config2_labels.npz, written by tests/golden/make_golden_config2.py); None without the file."""
    path = os.path.join(REPO, "tests", "golden", "config2_labels.npz")
    if not os.path.exists(path):
        return None
    gold = np.load(path)["labels"]
    got = labels.cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
    return bool(np.array_equal(got.reshape(gold.shape).astype(np.int64), gold.astype(np.int64)))


# ------------------------------------------------------------------------------- B200 arm: detect
def bench_detect(dev, windows, with_cpu):
    """Second half of the metric: threshold -> labels on BASELINE configs[2], inputs resident in HBM."""
    from cellulus_b200 import kernels as K
    from cellulus_b200 import synthetic
    from cellulus_b200.detect import detect_embeddings
    from cellulus_b200.utils.mean_shift import segment_embeddings_device

    emb_np, _, ids = synthetic.blob_scene(DET_SHAPE, DET_OBJECTS, radius=DET_RADIUS, seed=0)
    emb = torch.from_numpy(emb_np).to(dev)
    n_vox = int(np.prod(DET_SHAPE))
    fg = int((ids > 0).sum())
    kw = dict(bandwidth=DET_BW, threshold=DET_THR, reduction_probability=DET_RP, rng="philox")
    # (a) parity at full size: the fit subset drawn like the reference (np.random under seed 0), labels compared
    #     one for one with the reference's own output on this volume
    np.random.seed(0)
    labels_np_rng, _, _ = detect_embeddings(emb.double(), DET_BW, DET_THR, 1, DET_RP, rng="numpy", label_dtype=torch.int32)
    equal_golden = golden_labels_equal(labels_np_rng[0])
    # (b) the hill-climb kernel alone, step-by-step path: its time, its distance tests, its share
    for _ in range(2):
        labels, info = segment_embeddings_device(emb, DET_BW, DET_THR, DET_RP, rng="philox", method="grid",
                                                 label_dtype=torch.uint16)
    n_fit, n_seeds = int(info["n_fit"]), int(info["n_seeds"])
    grid = info["grid"]
    pts, _, n, _ = K.fg_compact(emb, DET_THR)
    fit_pts, n_fit2 = K.select_points(pts, n, K.bernoulli_flags(n, DET_RP, 0, dev))
    sorted_pts, cell_start, _ = K.grid_build(fit_pts, n_fit2, grid)
    def time_climb(fn):
        """Device time of one hill climb (all of its launches), replayed from a CUDA graph so that the host side of the
        Python wrapper (allocations, ctypes) is not what is measured; its distance tests and window evaluations."""
        seeds = fit_pts.clone()
        fn(sorted_pts, n_fit2, grid, cell_start, seeds, n_fit2, DET_BW)  # warm-up outside the capture
        torch.cuda.synchronize(dev)
        # two graphs: (seed reset + climb) and (seed reset) alone; the difference of their replay times is the climb
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            seeds.copy_(fit_pts)
            fn(sorted_pts, n_fit2, grid, cell_start, seeds, n_fit2, DET_BW)
        copy_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(copy_graph):
            seeds.copy_(fit_pts)

        def replay_ms(g):
            times = []
            for _ in range(6):
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize(dev)
                times.append(e0.elapsed_time(e1))
            return float(np.median(times[1:]))

        ms = replay_ms(graph) - replay_ms(copy_graph)
        return ms, K.grid_modes_distance_tests(), K.grid_modes_climb_steps()

    # every seed to convergence (what scikit-learn does: the algorithmic work) ...
    full_ms, full_tests, full_steps = time_climb(K.ms_grid_modes)
    # ... and what cb200_detect_volume runs: one evaluation per seed, then one representative per distinct unfinished mean
    k_ms, tests, steps = time_climb(K.ms_grid_modes_distinct)
    flop_per_test = 3 * 3 + 2  # D sub, D mul/fma, 1 compare, D + 1 predicated adds
    fp64_peak = K.fma_peak_tflops(torch.float64, dev)
    fp32_peak = K.fma_peak_tflops(torch.float32, dev)
    achieved_tf = tests * flop_per_test / (k_ms * 1e-3) / 1e12
    # (c) the timed path: one C-ABI call per volume (its scratch arena grows during the warm-up)
    for _ in range(3):
        detect_embeddings(emb, **kw)
    torch.cuda.synchronize(dev)
    reps = 10
    c0 = K.launch_counter["calls"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(reps):
        labels, _, _ = detect_embeddings(emb, **kw)
    e1.record()
    torch.cuda.synchronize(dev)
    windows.append((w0, time.time()))
    ms = e0.elapsed_time(e1) / reps
    calls = (K.launch_counter["calls"] - c0) // reps
    # (d) e2e: pinned host volume in, uint16 labels out, wall clock per volume (median of 10)
    host = torch.from_numpy(emb_np).pin_memory()
    out_host = torch.empty((1, *DET_SHAPE), dtype=torch.uint16).pin_memory()
    e2e = []
    for _ in range(12):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        d = host.to(dev, non_blocking=True)
        labels, _, _ = detect_embeddings(d, **kw)
        out_host.copy_(labels, non_blocking=True)
        torch.cuda.synchronize(dev)
        e2e.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e[2:]))
    cpu = None
    if with_cpu:  # bounded sample of the same scene family (the full volume runs in the --impl reference arm)
        cpu, _ = cpu_detect((40, 128, 128), 31, "sub-volume at the object density of configs[2]")
        cpu.pop("centres")
    # (e) the alternative clustering of detect.py:162-192 (`clustering = "greedy"`, utils/greedy_cluster.py) on the same
    #     volume: one persistent cooperative kernel per volume instead of ~10 eager kernels + a host sync per object
    fg_mask = (emb[3] < DET_THR).to(torch.uint8)
    K.greedy_cluster(emb, fg_mask, DET_BW, 10)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        inst, n_obj, n_tried = K.greedy_cluster(emb, fg_mask, DET_BW, 10)
    torch.cuda.synchronize(dev)
    greedy_s = (time.perf_counter() - t0) / 3
    greedy = {"ms_per_volume": greedy_s * 1e3, "Mpx/s": n_vox / greedy_s / 1e6, "objects": int(n_obj),
              "seeds_tried": int(n_tried), "note": "public call, wall clock incl. the compaction and the result read"}
    if with_cpu:
        from oracle import greedy as ogreedy

        sub, _, _ = synthetic.blob_scene((40, 128, 128), 31, radius=DET_RADIUS, seed=0)
        t0 = time.perf_counter()
        _, n_sub, _ = ogreedy.greedy_cluster(sub, sub[3] < DET_THR, DET_BW, 10)
        dt = time.perf_counter() - t0
        greedy["cpu_baseline"] = {"Mpx/s": sub[0].size / dt / 1e6, "kind": "port", "cores": torch.get_num_threads(),
                                  "sample": f"40x128x128 sub-volume, {int(n_sub)} objects, {dt:.2f} s (torch CPU, the "
                                            "reference's loop)"}
    return {
        "metric": "detect Mpx/s (mean-shift)", "value": n_vox / ms / 1e3, "unit": "Mpx/s", "ms_per_volume": ms,
        "config": {"workload": "configs[2]: 128x256x256 volume, 3-D embeddings, 400 balls r=10, bw=7, threshold=0.5, "
                               "reduction_probability=0.1, seeds=all fit points",
                   "foreground_voxels": fg, "fit_points": n_fit, "seeds": n_seeds, "centres": int(info["k"]),
                   "method": "grid hash (cells of edge >= bandwidth), one warp per seed"},
        "abi_calls_per_volume": int(calls),
        "greedy_clustering": greedy,
        "labels_equal_reference_golden": equal_golden,
        "roofline": {"bound": "fp64-pipe",
                     "kernel": "ms_grid_modes_kernel<3>, two launches (cb200_ms_grid_modes_distinct: one window evaluation per "
                               "seed, then the distinct unfinished means) incl. the merge pass between them",
                     "kernel_ms": k_ms,
                     "every_seed_to_convergence": {"kernel_ms": full_ms, "distance_tests": full_tests, "climb_steps": full_steps,
                                                   "achieved": full_tests * flop_per_test / (full_ms * 1e-3) / 1e12,
                                                   "frac": full_tests * flop_per_test / (full_ms * 1e-3) / 1e12 / fp64_peak,
                                                   "note": "cb200_ms_grid_modes, one launch: the work scikit-learn does per "
                                                           "seed; the shipped form does 47 % fewer tests in 20 % less time "
                                                           "(its second pass is a latency tail of a few long climbs)"},
                     "kernel_share_of_volume": k_ms / ms,
                     "distance_tests_per_launch": tests, "flop_per_test": flop_per_test,
                     "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                     "peak_source": "measured in this run (cb200_fma_peak, DFMA chains, CUDA events)",
                     "fp32_fma_peak_tflops": fp32_peak,
                     "distance_tests_per_s": tests / (k_ms * 1e-3),
                     "climb_steps": steps, "brute_force_tests": steps * n_fit2,
                     "pruning_ratio": steps * n_fit2 / max(tests, 1),
                     "gathered_GBps": tests * 3 * 8 / (k_ms * 1e-3) / 1e9,
                     "traffic": None,
                     "note": "arithmetic is float64 without FMA contraction (the in/out decisions are the reference "
                             "KD-tree's); the gathered cell ranges are L1/L2 hits (DRAM ~0: profiles/)"},
        "cpu_baseline": cpu,
        "e2e": {"value": n_vox / e2e_s / 1e6, "unit": "Mpx/s", "ms_per_volume": e2e_s * 1e3,
                "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 2,
                "note": "median wall clock of 10 volumes: pinned host volume in, uint16 labels out"},
    }


def bench_tta(dev, windows):
    """TTA aggregate (models/unet.py:90-98) of T = 32 passes over a 496 x 496 scan block and a 3-D block."""
    from cellulus_b200.models import tta_aggregate

    out = {}
    peak, _ = measured_peak_gbs()
    for name, shape in [("2d_32x2x496x496", (32, 2, 496, 496)), ("3d_32x3x20x212x212", (32, 3, 20, 212, 212))]:
        stacks = [torch.randn(shape, device=dev) for _ in range(3 if name.startswith("2d") else 2)]
        nbytes = stacks[0].numel() * 4 + (shape[1] + 1) * int(np.prod(shape[2:])) * 4
        # rotate over enough distinct stacks that none is L2-resident when re-read (66 MB each in 2-D); the
        # calls are replayed from a CUDA graph so that the host's launch rate is not what is measured
        for i in range(6):
            tta_aggregate(stacks[i % len(stacks)])
        torch.cuda.synchronize(dev)
        per_graph = 2 * len(stacks)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            keep = [tta_aggregate(stacks[i % len(stacks)]) for i in range(per_graph)]
        graph.replay()
        torch.cuda.synchronize(dev)
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for i in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        windows.append((w0, time.time()))
        us = e0.elapsed_time(e1) / (reps * per_graph) * 1e3
        del graph, keep
        out[name] = {"us": us, "GB/s": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak,
                     "Mpx/s": float(np.prod(shape[2:])) / us, "algorithmic_bytes": nbytes}
        del stacks
    return out


def _label_scene(shape, n_blobs, radius, seed):
    """Label image of filled discs + a raw image with a bright shell / dim core per object (uint16)."""
    rng = np.random.default_rng(seed)
    seg = np.zeros(shape, np.int32)
    raw = rng.random(shape) * 0.3
    yy, xx = np.mgrid[: shape[0], : shape[1]]
    for k in range(n_blobs):
        cy, cx, r = rng.uniform(0, shape[0]), rng.uniform(0, shape[1]), rng.uniform(0.6 * radius, radius)
        y0, y1, x0, x1 = int(max(cy - r, 0)), int(min(cy + r + 1, shape[0])), int(max(cx - r, 0)), int(min(cx + r + 1, shape[1]))
        d = np.sqrt((yy[y0:y1, x0:x1] - cy) ** 2 + (xx[y0:y1, x0:x1] - cx) ** 2)
        m = d < r
        seg[y0:y1, x0:x1][m] = k + 1
        raw[y0:y1, x0:x1][m] = np.where(d[m] > 0.45 * r, 0.9, 0.3) + 0.1 * rng.random(int(m.sum()))
    return seg, (raw / raw.max() * 40000).astype(np.uint16)


def bench_post(dev, with_cpu):
    """The rows after the path: segment() post-processing and evaluate() counts on a 2048^2 image
    with ~1200 objects; CPU legs (oracle = scipy / the reference's loops) on bounded samples."""
    from cellulus_b200 import kernels as K
    from cellulus_b200.evaluate import compute_pairwise_IoU
    from cellulus_b200.segment import nucleus

    shape, blobs = (2048, 2048), 1200
    seg_np, raw_np = _label_scene(shape, blobs, 26, seed=3)
    seg = torch.from_numpy(seg_np).to(dev)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / reps

    n = float(np.prod(shape))
    t_gs = timed(lambda: K.grow_shrink_(seg.clone(), 3, 6))
    t_sf = timed(lambda: K.size_filter_(seg.clone(), 25))
    t_nuc = timed(lambda: nucleus(seg_np, raw_np), reps=3)  # numpy in / numpy out
    gt = np.roll(seg_np, (3, -2), axis=(0, 1)).astype(np.uint16)
    t_ev = timed(lambda: compute_pairwise_IoU(seg_np.astype(np.uint16), gt), reps=3)
    out = {
        "workload": f"{shape[0]}x{shape[1]} label image, {int(seg_np.max())} objects, uint16 raw",
        "grow_shrink": {"ms": t_gs * 1e3, "Mpx/s": n / t_gs / 1e6, "note": "device tensor in place (segment.py:46-50)"},
        "size_filter": {"ms": t_sf * 1e3, "Mpx/s": n / t_sf / 1e6, "note": "device tensor (utils/misc.py:11-25)"},
        "nucleus": {"ms": t_nuc * 1e3, "Mpx/s": n / t_nuc / 1e6, "note": "numpy in/out incl. the host<->device copies (per-instance Otsu runs on the device; the Otsu rule "
                            "is the restated scikit-image threshold_otsu: parity unpinned, DESIGN 2)"},
        "evaluate_tables": {"ms": t_ev * 1e3, "Mpx/s": n / t_ev / 1e6, "note": "numpy in/out, IoU table of all id pairs"},
    }
    if with_cpu:
        from oracle import evaluate as oeval
        from oracle import post_process as opost

        small, small_raw = seg_np[:1024, :1024].copy(), raw_np[:1024, :1024]
        t0 = time.perf_counter()
        opost.grow_shrink(small.copy(), 3, 6)
        t_c = time.perf_counter() - t0
        out["grow_shrink"]["cpu_baseline"] = {"Mpx/s": small.size / t_c / 1e6, "kind": "port",
                                              "sample": "1024x1024 corner, scipy distance_transform_edt x2, 1 core"}
        tiny, tiny_raw = seg_np[:512, :512].copy(), raw_np[:512, :512]
        t0 = time.perf_counter()
        opost.nucleus(tiny, tiny_raw)
        t_c = time.perf_counter() - t0
        out["nucleus"]["cpu_baseline"] = {"Mpx/s": tiny.size / t_c / 1e6, "kind": "port",
                                          "sample": f"512x512 corner, {len(np.unique(tiny)) - 1} objects, 1 core"}
        tiny_gt = gt[:384, :384]
        t0 = time.perf_counter()
        oeval.compute_pairwise_IoU(seg_np[:384, :384].astype(np.uint16), tiny_gt)
        t_c = time.perf_counter() - t0
        out["evaluate_tables"]["cpu_baseline"] = {"Mpx/s": tiny_gt.size / t_c / 1e6, "kind": "port",
                                                  "sample": f"384x384 corner, {len(np.unique(tiny_gt)) - 1} ground-truth "
                                                            "objects (cost grows with the product of the id counts)"}
    return out


def run_b200(args):
    from cellulus_b200 import kernels as K
    from cellulus_b200.criterions import GraphedLossCycle, GraphedLossStep, oce_loss_fused, oce_loss_fused_sampled
    from cellulus_b200.datasets import PairListStager

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    else:
        dist = None
    if args.gpus != world:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    if world > 1:  # torchrun pins OMP_NUM_THREADS=1; the host-side list conversion of the e2e leg shares the cores evenly
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))

    sampler = ClockSampler(local) if rank == 0 else None
    windows = []

    # inputs resident in HBM: every rank owns its own batches (sharded by batch).  N_SETS distinct input
    # sets are visited round-robin so that no step finds its 180 MB of coordinate lists in the 126 MB L2.
    N_SETS = 3
    torch.manual_seed(rank)

    # Headline pair lists: the reference's own sampler (zarr_dataset.py:177-251, restated in the dataset module of this
    # package on numpy's global generator) under np.random.seed, samples drawn b = 0..7 in order -- SURVEY 8d.
    from cellulus_b200.datasets.zarr_dataset import sample_coordinates

    host_lists = []
    for i in range(N_SETS):
        np.random.seed(N_SETS * rank + i)
        pairs = [sample_coordinates(OUT, KAPPA, N_ANCHORS, N_REFS, D) for _ in range(B)]
        host_lists.append((torch.from_numpy(np.stack([p[0] for p in pairs])).to(dev),
                           torch.from_numpy(np.stack([p[1] for p in pairs])).to(dev)))
    assert tuple(host_lists[0][0].shape) == (B, P, D) and host_lists[0][0].dtype == torch.int64

    def make_steps(memory_format, dtype=torch.float32, coord_dtype=torch.int64, reference_lists=False):
        steps = []
        for i in range(N_SETS):
            off = torch.randn(B, D, *OUT, device=dev).to(dtype).contiguous(memory_format=memory_format)
            if reference_lists:
                anc, ref = host_lists[i]
            else:
                anc, ref = K.sample_pairs(B, (OUT[1], OUT[0]), KAPPA, N_ANCHORS, N_REFS, seed=1234 + 17 * rank + i,
                                          device=dev, dtype=coord_dtype)
            steps.append(GraphedLossStep(off, anc, ref, TEMP, REGW))
        return steps

    def timed(steps, n_steps, n_warm):
        """EXACTLY n_steps steps, visiting the N_SETS input sets round-robin.  Whole rounds are replayed as ONE graph
        of N_SETS steps (kernel -> kernel edges inside; a graph launch per step leaves ~1.5 us of front-end gap on
        the stream), the remainder as single-step graphs.  Every step does its full work: its own gradient
        zero-fill and its own fused kernel on its own 211 MB of inputs."""
        cycle = GraphedLossCycle(steps)

        def run(n):
            for _ in range(n // N_SETS):
                cycle.replay()
            for i in range(n % N_SETS):
                steps[i].replay()

        run(max(n_warm, N_SETS))
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        calls0 = K.launch_counter["calls"]
        run(n_steps)
        timed.calls = K.launch_counter["calls"] - calls0  # C-ABI loss calls inside the timed region (= n_steps)
        e1.record()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        windows.append((w0, time.time()))
        t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        del cycle
        return t_ms.item() / n_steps

    warm = max(args.warmup, 3)
    # (1) headline: channels-last offsets (the layout this framework keeps the U-Net output in)
    steps_cl = make_steps(torch.channels_last, reference_lists=True)
    ms_per_step = timed(steps_cl, args.steps, warm)
    launches = timed.calls * 2  # zero-fill + fused kernel per step
    value = world * N_PX / (ms_per_step * 1e-3)
    loss_value = steps_cl[0].loss.item()
    loss_oracle = None
    if rank == 0 and world == 1 and not args.skip_cpu:  # the same batch through the float64 oracle arithmetic
        from oracle import oce_loss as oloss

        loss_oracle = oloss.loss_step_float64(steps_cl[0].offsets.cpu().contiguous(), steps_cl[0].anchors.cpu(),
                                              steps_cl[0].refs.cpu(), TEMP, REGW)[0].item()
    # (2) same op on the planar NCHW tensor the reference's model emits
    steps_pl = make_steps(torch.contiguous_format, reference_lists=True)
    ms_planar = timed(steps_pl, args.steps, warm)
    # (2b) bf16 offsets (BASELINE configs[1] trains the U-Net in bf16): bf16 storage, fp32 arithmetic and gradient
    steps_bf = make_steps(torch.channels_last, torch.bfloat16)
    ms_bf16 = timed(steps_bf, args.steps, warm)
    del steps_bf
    algo_bf16 = B * P * D * 8 * 2 + N_PX * D * 2 + N_PX * D * 4
    # (2c) int16 coordinate lists, the format the device pair sampler hands to the training loop (train.py)
    steps_i16 = make_steps(torch.channels_last, torch.float32, torch.int16)
    ms_i16 = timed(steps_i16, args.steps, warm)
    del steps_i16
    algo_i16 = B * P * D * 2 * 2 + N_PX * D * 4 * 2
    # (2d) fused-sampling mode (north_star: "the loss fuses sampling, neighbour gather and the per-pair terms
    # into one pass"): the pairs are drawn inside the kernel, no coordinate list exists; SURVEY 8d: 31.5 MB
    def make_sampled(memory_format):
        return [GraphedLossStep(torch.randn(B, D, *OUT, device=dev).contiguous(memory_format=memory_format), None, None,
                                TEMP, REGW, sampled=dict(kappa=KAPPA, num_anchors=N_ANCHORS, num_references=N_REFS,
                                                         seed=4321 + 17 * rank + i, extent_xyz=(OUT[1], OUT[0])))
                for i in range(N_SETS)]

    steps_s = make_sampled(torch.channels_last)
    ms_sampled = timed(steps_s, args.steps, warm)
    steps_s = make_sampled(torch.contiguous_format)
    ms_sampled_planar = timed(steps_s, args.steps, warm)
    del steps_s
    algo_sampled = N_PX * D * 4 * 2
    peak, peak_src = measured_peak_gbs()
    achieved = ALGO_BYTES / (ms_per_step * 1e-3) / 1e9
    achieved_planar = ALGO_BYTES / (ms_planar * 1e-3) / 1e9
    offsets, anchors, refs = steps_pl[0].offsets, steps_pl[0].anchors, steps_pl[0].refs
    del steps_pl

    # (3) end to end through the public API from pinned HOST buffers: every step copies that step's offsets and
    # both int64 pair lists (the format the reference's DataLoader delivers, train.py:162-166) host -> device,
    # runs loss + backward and reads the loss back.  Double-buffered: the copy of step i + 1 runs on a second
    # stream while step i computes; the pinned buffers are allocated after this rank's CPU affinity has been
    # narrowed to the NUMA node of its GPU (first touch places them there).
    numa = bind_to_gpu_numa_node(local)
    h_off = offsets.detach().cpu().contiguous().pin_memory()
    h_anc, h_ref = anchors.cpu().pin_memory(), refs.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_run(lists, n_steps, narrow=False):
        """`lists`: pinned (anchors, refs) or None (pairs drawn inside the kernel).  Returns seconds per step.
        `narrow`: the int64 host lists are converted to int16 on the HOST, into pinned staging (`PairListStager`:
        all host threads, two slots), and only the int16 copy crosses PCIe."""
        bufs = []
        stager = PairListStager(lists[0].shape, dev) if narrow else None
        for _ in range(2):
            bufs.append([torch.empty_like(h_off, device=dev)] +
                        ([torch.empty_like(t, device=dev, dtype=torch.int16 if narrow else t.dtype) for t in lists]
                         if lists else []))
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        main_stream = torch.cuda.current_stream(dev)

        def upload(i):
            slot = i % 2
            staged = stager.narrow(lists[0], lists[1], slot) if narrow else lists
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[slot])  # the step that last used this slot has finished
                bufs[slot][0].copy_(h_off, non_blocking=True)
                if lists:
                    bufs[slot][1].copy_(staged[0], non_blocking=True)
                    bufs[slot][2].copy_(staged[1], non_blocking=True)
                ready[slot].record(copy_stream)
                if narrow:
                    stager.copied(slot, copy_stream)

        for ev in done:
            ev.record(main_stream)
        upload(0)
        t0 = None
        for i in range(n_steps + 2):
            if i == 2:  # two untimed warm-up steps
                torch.cuda.synchronize(dev)
                if dist is not None:
                    dist.barrier()
                t0 = time.perf_counter()
            upload(i + 1)
            slot = i % 2
            main_stream.wait_event(ready[slot])
            o = bufs[slot][0].requires_grad_(True)
            if lists:
                loss, _, _ = oce_loss_fused(o, bufs[slot][1], bufs[slot][2], TEMP, REGW)
            else:
                loss, _, _ = oce_loss_fused_sampled(o, KAPPA, N_ANCHORS, N_REFS, 77 + rank, i, TEMP, REGW,
                                                    extent_xyz=(OUT[1], OUT[0]))
            loss.backward()
            done[slot].record(main_stream)
            bufs[slot][0] = o.detach()
            loss.item()  # device -> host read of the step's result (train.py:180)
        torch.cuda.synchronize(dev)
        sec = torch.tensor([(time.perf_counter() - t0) / n_steps], device=dev)
        if dist is not None:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return sec.item()

    w0 = time.time()
    e2e_s = e2e_run((h_anc, h_ref), e2e_steps)
    windows.append((w0, time.time()))
    e2e_value = world * N_PX / e2e_s
    h2d_bytes = int(h_off.numel() * 4 + h_anc.numel() * 8 + h_ref.numel() * 8)
    # (3b) the same step with the lists narrowed to int16 by the DataLoader workers (ZarrDataset(coordinate_dtype=
    # "int16"): a quarter of the PCIe bytes) and (3c) with the pairs drawn inside the kernel: only the offsets cross
    h_anc16, h_ref16 = h_anc.to(torch.int16).pin_memory(), h_ref.to(torch.int16).pin_memory()
    e2e_i16_s = e2e_run((h_anc16, h_ref16), e2e_steps)
    e2e_narrow_s = e2e_run((h_anc, h_ref), e2e_steps, narrow=True)
    e2e_sampled_s = e2e_run(None, e2e_steps)
    del h_anc16, h_ref16

    # (4) multi-GPU inference jobs, wall clock, every rank takes part (also run at N = 1: the scaling baseline)
    multi = None
    if not args.skip_detect and not args.skip_multi:
        multi = bench_multi_gpu(dev, rank, world, dist)

    detect = None
    cpu_base = None
    if rank == 0:
        if not args.skip_detect:
            tta = bench_tta(dev, windows)
            post = bench_post(dev, with_cpu=(world == 1 and not args.skip_cpu))
            detect = bench_detect(dev, windows, with_cpu=(world == 1 and not args.skip_cpu))
        if world == 1 and not args.skip_cpu:
            t_cpu = cpu_loss_step_time(B, 3, 1)
            cpu_base = {"value": N_PX / t_cpu, "unit": "px/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"full configs[1] batch ({B * P} pairs), 3 steps after 1 warm-up, {t_cpu * 1e3:.0f} ms/step",
                        "pairs_per_s": B * P / t_cpu}
    if dist is not None:
        dist.barrier()
    if rank == 0:
        clocks = sampler.stop(windows)
        line = {
            "metric": "OCELoss fwd+bwd px/s", "value": value, "unit": "px/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": B * P, "px_per_step_per_gpu": N_PX,
                       "offsets_layout": "channels_last (B,H,W,2 in memory; same logical (8,2,496,496) tensor)",
                       "l2": f"inputs larger than L2: {N_SETS} distinct input sets of 211 MB visited round-robin "
                             "(126 MB L2), no explicit flush",
                       "step": "zero-fill + fused gather/loss/backward kernel (programmatic dependent launch); replayed from CUDA "
                               "graphs of 3 steps (one per input set), the remainder of K as single-step graphs",
                       "pairs": "int64 lists from the reference's sampler (zarr_dataset.py:177-251 restated on numpy's global "
                                "generator, np.random.seed(3 * rank + set)); the bf16 / int16 variants use the device pair stream",
                       "sharding": "by batch, one batch per rank, no data-path collective"},
            "pairs_per_s": world * B * P / (ms_per_step * 1e-3),
            "loss_value_check": {"kernel": loss_value, "oracle_float64": loss_oracle,
                                 "rel_err": abs(loss_value - loss_oracle) / abs(loss_oracle) if loss_oracle else None},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("oce_loss_fused_kernel<2,int64,f32,bwd,channels_last>"),
                         "peak_source": peak_src,
                         "kernel": "oce_loss_fused_kernel<2,int64,f32,bwd,channels_last> (+ zero_fill_kernel)",
                         "kernel_ms": ms_per_step, "algorithmic_bytes_per_launch": ALGO_BYTES,
                         "note": "duration is the whole step: the 15.7 MB gradient zero-fill is included"},
            "planar": {"value": world * N_PX / (ms_planar * 1e-3), "unit": "px/s", "ms_per_step": ms_planar,
                       "offsets_layout": "planar NCHW (what the reference's model emits); ONE kernel per step: it clears "
                                         "the gradient and gathers from its own channels-last copy of the offsets "
                                         "(15.7 MB staging scratch written inside the same launch)",
                       "roofline": {"bound": "hbm", "achieved": achieved_planar, "peak": peak, "unit": "GB/s",
                                    "frac": achieved_planar / peak,
                                    "traffic": ncu_traffic("oce_loss_staged_kernel<int64,f32,bwd>"),
                                    "kernel": "oce_loss_staged_kernel<int64,f32,bwd> (one launch per step)",
                                    "algorithmic_bytes_per_launch": ALGO_BYTES}},
            "bf16_offsets": {"value": world * N_PX / (ms_bf16 * 1e-3), "unit": "px/s", "ms_per_step": ms_bf16,
                             "offsets_layout": "channels_last, bf16 storage, fp32 arithmetic, fp32 gradient",
                             "roofline": {"bound": "hbm", "achieved": algo_bf16 / (ms_bf16 * 1e-3) / 1e9,
                                          "peak": peak, "unit": "GB/s",
                                          "frac": algo_bf16 / (ms_bf16 * 1e-3) / 1e9 / peak,
                                          "algorithmic_bytes_per_launch": algo_bf16}},
            "i16_pairs": {"value": world * N_PX / (ms_i16 * 1e-3), "unit": "px/s", "ms_per_step": ms_i16,
                          "note": "channels_last fp32 offsets, int16 coordinate lists (the reference's lists are int64)",
                          "roofline": {"bound": "hbm", "achieved": algo_i16 / (ms_i16 * 1e-3) / 1e9, "peak": peak,
                                       "unit": "GB/s", "frac": algo_i16 / (ms_i16 * 1e-3) / 1e9 / peak,
                                       "algorithmic_bytes_per_launch": algo_i16}},
            "fused_sampling": {"value": world * N_PX / (ms_sampled * 1e-3), "unit": "px/s", "ms_per_step": ms_sampled,
                               "pairs_per_s": world * B * P / (ms_sampled * 1e-3),
                               "planar_ms_per_step": ms_sampled_planar,
                               "note": "cb200_oce_loss_sampled: pairs drawn inside the kernel (device pair stream), no "
                                       "coordinate list in HBM; what train() runs per step",
                               "roofline": {"bound": "hbm", "achieved": algo_sampled / (ms_sampled * 1e-3) / 1e9,
                                            "peak": peak, "unit": "GB/s",
                                            "frac": algo_sampled / (ms_sampled * 1e-3) / 1e9 / peak,
                                            "algorithmic_bytes_per_launch": algo_sampled,
                                            "note": "31.5 MB: at this size the kernel is bound by the L1 -> L2 request "
                                                    "rate of its scattered 8-byte gathers (profiles/), not by HBM"}},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": "px/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "ms_per_step": e2e_s * 1e3,
                    "h2d_GBps_per_rank": h2d_bytes / e2e_s / 1e9, "numa": numa,
                    "note": "int64 lists as the reference's DataLoader delivers them; double-buffered uploads; the step "
                            "is the host -> device copy (PCIe / host-memory bound), the kernels are 1-2 % of it",
                    "int16_host_lists": {"value": world * N_PX / e2e_i16_s, "unit": "px/s", "ms_per_step": e2e_i16_s * 1e3,
                                         "h2d_bytes_per_step": int(h_off.numel() * 4 + h_anc.numel() * 2 * 2)},
                    "int64_host_lists_narrowed_on_host": {
                        "value": world * N_PX / e2e_narrow_s, "unit": "px/s", "ms_per_step": e2e_narrow_s * 1e3,
                        "h2d_bytes_per_step": int(h_off.numel() * 4 + h_anc.numel() * 2 * 2),
                        "host_threads": torch.get_num_threads(),
                        "note": "the same int64 lists, converted to int16 into pinned staging by the host threads "
                                "(cellulus_b200.datasets.PairListStager) while the previous step computes"},
                    "pairs_drawn_in_kernel": {"value": world * N_PX / e2e_sampled_s, "unit": "px/s",
                                              "ms_per_step": e2e_sampled_s * 1e3,
                                              "h2d_bytes_per_step": int(h_off.numel() * 4)}},
            "gpu_launches": int(launches * world),
            "clocks": clocks,
            "multi_gpu": multi,
            "detect": None,
        }
        if detect is not None:
            detect = dict(detect)
            detect["tta_aggregate"] = tta
            detect["post_processing"] = post
            # the detect core numbers go LAST so that they are what a tail of this line shows
            core = {k: detect.pop(k) for k in ["config", "labels_equal_reference_golden", "cpu_baseline", "e2e", "roofline",
                                               "metric", "unit", "ms_per_volume", "value"]}
            ordered = {k: detect[k] for k in ["tta_aggregate", "post_processing", "greedy_clustering", "abi_calls_per_volume"]}
            ordered.update(core)
            line["detect"] = ordered
            # the same detect headline as flat top-level keys (a parser that keeps scalars only still sees them)
            line["detect_value_mpx_s"] = core["value"]
            line["detect_ms_per_volume"] = core["ms_per_volume"]
            line["detect_e2e_mpx_s"] = core["e2e"]["value"]
            line["detect_roofline_frac"] = core["roofline"]["frac"]
            line["detect_labels_equal_reference_golden"] = core["labels_equal_reference_golden"]
            if core["cpu_baseline"]:
                line["detect_cpu_baseline_mpx_s"] = core["cpu_baseline"]["value"]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(local_rank):
    """Narrow this process to the CPUs of the NUMA node its GPU hangs off, so that pinned staging buffers are
    first-touched there.  Returns a short description (or why it was not done)."""
    try:
        bus = torch.cuda.get_device_properties(local_rank)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read())
        if node < 0:
            return f"gpu {pci}: no NUMA information"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return f"gpu {pci} on node {node}, {len(allowed)} cpus"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


# ------------------------------------------------------------------------------- B200 arm: multi-GPU inference
SHARD_SHAPE, SHARD_OBJECTS, SHARD_BW = (272, 272, 272), 855, 10.0  # ~3.6 M foreground voxels, every one a seed


def bench_multi_gpu(dev, rank, world, dist):
    """Wall-clock jobs on the multi-GPU inference paths.

    sharded_volume: ONE 272^3 volume whose foreground voxels are all seeds; every rank compacts its z-slab, the
        point set is all-gathered, each rank climbs its slice of the seeds, (mode, count) are all-gathered,
        suppression is replicated, labels are assigned per slab.  Job = max over ranks of the device time of
        `sharded_mean_shift` (3 runs, last one reported, phases from CUDA events).  The N-rank centres and label
        checksums are asserted equal to the 1-rank pipeline run on rank 0.
    blockwise: BASELINE configs[4] -- a 16384^2 mosaic and a 1024^3 volume, scan blocks dealt round-robin; per
        block the T = 32 TTA predictions are generated on the device, aggregated, detected and the uint16 labels are
        copied to pinned host memory.  Job = ONE time.perf_counter around the whole loop per rank (generation,
        launches, host syncs, copies all inside), max over ranks.
    """
    from cellulus_b200 import kernels as K
    from cellulus_b200 import sharding, synthetic
    from cellulus_b200.detect import detect_embeddings
    from cellulus_b200.models import tta_aggregate
    from cellulus_b200.utils.mean_shift import segment_embeddings_device

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def allsum(x):
        t = torch.tensor([x], device=dev, dtype=torch.int64)
        if dist is not None:
            dist.all_reduce(t)
        return int(t.item())

    out = {}
    # ---- one volume, seeds sharded
    emb, _, _ = synthetic.blob_scene(SHARD_SHAPE, SHARD_OBJECTS, radius=10.0, seed=0)
    zs = sharding.shard_items(SHARD_SHAPE[0], rank, world)
    slab = torch.from_numpy(np.ascontiguousarray(emb[:, zs.start:zs.stop])).to(dev)
    pts, pix, n_local, _ = K.fg_compact(slab, 0.5)
    pts[2, :n_local] += float(zs.start)  # z of the slab inside the volume (column 2 = z)
    ops = sharding.cuda_ops("grid")
    phases, ms = {}, 0.0
    for _ in range(3):
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        phases = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        labels, centres = sharding.sharded_mean_shift(pts, n_local, SHARD_BW, ops, None, None, timings=phases)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = allmax(e0.elapsed_time(e1))
    n_fg = allsum(n_local)
    checksum = allsum(int((labels.to(torch.int64) * (torch.arange(n_local, device=dev) % 1009 + 1)).sum().item()))
    equal = None
    if rank == 0:  # the 1-rank pipeline on the whole volume: same centres, same labels
        whole = torch.from_numpy(emb).to(dev)
        ref_labels, info = segment_embeddings_device(whole, SHARD_BW, 0.5, 1.0, method="grid")
        ref = ref_labels[whole[3] < 0.5].to(torch.int64)
        bounds = [0]
        for r in range(world):
            z = sharding.shard_items(SHARD_SHAPE[0], r, world)
            bounds.append(int((whole[3, :z.stop] < 0.5).sum().item()))
        ref_sum = 0
        for r in range(world):
            seg = ref[bounds[r]:bounds[r + 1]]
            ref_sum += int((seg * (torch.arange(seg.numel(), device=dev) % 1009 + 1)).sum().item())
        same_centres = bool(torch.equal(info["centres"], centres))
        equal = bool(same_centres and ref_sum == checksum)
        assert equal, f"sharded mean-shift differs from the 1-rank pipeline (centres equal: {same_centres})"
        del whole, ref_labels, ref
    out["sharded_volume"] = {
        "workload": f"one {SHARD_SHAPE[0]}^3 volume, {n_fg} foreground voxels, all of them seeds, bw {SHARD_BW}",
        "n_gpus": world, "job_ms": ms, "fg_points_per_s": n_fg / ms * 1e3, "centres": int(centres.shape[1]),
        "equals_one_rank_pipeline": equal, "collectives": "2 x all_gather (points; modes + counts), 1 x all_reduce (sizes)",
        "phases_ms_rank0": {k: round(v, 3) for k, v in phases.items()}}
    del slab, pts, pix, labels, emb
    # ---- blockwise inference, BASELINE configs[4]
    for kind, total, block in (("mosaic_16384x16384", (16384, 16384), (1024, 1024)),
                               ("volume_1024x1024x1024", (1024, 1024, 1024), (128, 256, 256))):
        nd = len(block)
        blocks = sharding.scan_blocks(total, block)
        mine = [blocks[i] for i in sharding.shard_round_robin(len(blocks), rank, world)]
        host = torch.empty((1, *block), dtype=torch.uint16).pin_memory()
        # The T = 32 predictions of a block stand for the U-Net's test-time-augmentation passes (not the product).
        # They are produced INSIDE the timed job, as cheaply as torch allows: the noise-free scene of this rank is
        # built once, every block draws its own noise in place into one preallocated (T, D, *block) buffer.
        clean, fg_mask, _ = synthetic.block_scene(block, 10.0, 10_000 + rank, dev)
        sigma = torch.where(fg_mask, 0.02, 1.0)[None]
        del fg_mask
        stack = torch.empty((32, nd, *block), device=dev)
        gen = torch.Generator(device=dev)

        def produce(seed):
            gen.manual_seed(seed)
            stack.normal_(generator=gen)
            stack.mul_(sigma).add_(clean)
            return stack

        detect_embeddings(tta_aggregate(produce(1)), bandwidth=7.0, threshold=0.5 * nd, reduction_probability=0.1,
                          rng="philox")  # warm-up block
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * len(mine) + 1)]
        t0 = time.perf_counter()
        labelled = 0
        ev[0].record()
        for i, b in enumerate(mine):
            seed = int(np.ravel_multi_index(tuple(o // s for o, s in zip(b, block)),
                                            tuple(t // s + 1 for t, s in zip(total, block))))
            st = produce(seed)
            ev[3 * i + 1].record()
            emb_b = tta_aggregate(st)
            labels_b, _, _ = detect_embeddings(emb_b, bandwidth=7.0, threshold=0.5 * nd, reduction_probability=0.1,
                                               rng="philox")
            ev[3 * i + 2].record()
            host.copy_(labels_b, non_blocking=True)
            ev[3 * i + 3].record()
            torch.cuda.synchronize(dev)
            labelled += int(np.count_nonzero(host.numpy()))
            del emb_b, labels_b
        job_s = allmax(time.perf_counter() - t0)
        gen_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(len(mine)))
        path_ms = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(len(mine)))
        copy_ms = sum(ev[3 * i + 2].elapsed_time(ev[3 * i + 3]) for i in range(len(mine)))
        px = float(np.prod(block)) * len(blocks)
        out[kind] = {"workload": f"configs[4]: {'x'.join(map(str, total))} in {len(blocks)} scan blocks of "
                                 f"{'x'.join(map(str, block))}; per block: T=32 predictions produced on the device "
                                 "(stand-in for the U-Net passes), TTA aggregate + detect (bw 7, rp 0.1), labels to pinned host",
                     "n_gpus": world, "job_s": job_s, "Mpx_per_s": px / job_s / 1e6, "foreground_px": allsum(labelled),
                     "blocks_per_gpu_max": -(-len(blocks) // world),
                     "device_ms_rank0": {"producing_predictions": gen_ms, "tta_aggregate_and_detect": path_ms,
                                         "labels_to_host": copy_ms},
                     "path_only_Mpx_per_s": px / world / max(path_ms, 1e-9) / 1e3,
                     "timing": "job_s = one perf_counter around the whole job per rank, max over ranks"}
        del stack, clean, sigma
    return out


# ------------------------------------------------------------------------------- BASELINE configs[3] sweep
def run_sweep(args):
    """`--sweep`: mean-shift on synthetic embeddings of 1M - 64M foreground points, D in {2, 3}, bandwidth in
    {0.5, 1, 2} x object radius, seeds = every foreground point vs scikit-learn's grid-binned seeds
    (SURVEY 8d D2).  One JSON line {"sweep": [...]}; per row: device time threshold -> labels (CUDA events),
    foreground points labelled per second.  CPU legs (the reference's engine, scikit-learn, hill climb on one core)
    at 1 M points: binned seeding in full, all-point seeding on a 200-seed subset with the extrapolation stated."""
    from cellulus_b200 import kernels as K
    from cellulus_b200 import synthetic
    from cellulus_b200.utils.mean_shift import segment_embeddings_device

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    radius = 10.0
    rows = []
    for D in (2, 3):
        fg_frac = (np.pi * radius**2) / (2.6 * radius) ** 2 if D == 2 else (4.0 / 3.0 * np.pi * radius**3) / (2.6 * radius) ** 3
        for n_m in (1, 4, 16, 64):
            if n_m > args.sweep_max:
                continue
            side = int(round((n_m * 1e6 / fg_frac) ** (1.0 / D)))
            shape = (side,) * D
            base, fg, gen = synthetic.block_scene(shape, radius, 7 * D + n_m, dev)
            emb = torch.empty((D + 1, *shape), device=dev)
            emb[:D] = base + torch.where(fg, 0.5, 1.0)[None] * torch.randn(base.shape, generator=gen, device=dev)
            emb[D] = torch.where(fg, 0.0, 1.0) + 0.1 * torch.rand(shape, generator=gen, device=dev)
            n_fg = int(fg.sum().item())
            del base, fg
            for bw_factor in (0.5, 1.0, 2.0):
                bw = bw_factor * radius
                for seeding in ("all_foreground", "grid_binned"):
                    kw = dict(reduction_probability=1.0, bin_seeding=(seeding == "grid_binned"), method="grid",
                              label_dtype=torch.int32, distinct=True)
                    if seeding == "all_foreground" and n_m * bw_factor**D > 64:
                        # every point a seed with a wide window: > 1e12 distance tests -- tens of seconds here, weeks on
                        # the CPU reference; left out of the sweep
                        rows.append({"D": D, "fg_points": n_fg, "bandwidth": bw, "seeds": seeding, "skipped": "work > 64 M x r^D"})
                        continue
                    try:
                        segment_embeddings_device(emb, bw, 0.5, **kw)  # warm-up (allocator, first launches)
                        torch.cuda.synchronize(dev)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        labels, info = segment_embeddings_device(emb, bw, 0.5, **kw)
                        e1.record()
                        torch.cuda.synchronize(dev)
                        ms = e0.elapsed_time(e1)
                        row = {"D": D, "shape": list(shape), "fg_points": n_fg, "bandwidth": bw, "seeds": seeding,
                               "n_seeds": int(info["n_seeds"]), "centres": int(info["k"]), "ms": round(ms, 3),
                               "fg_points_per_s": round(n_fg / ms * 1e3), "Mpx_per_s": round(float(np.prod(shape)) / ms / 1e3, 1),
                               "distance_tests": K.grid_modes_distance_tests(), "climb_steps": K.grid_modes_climb_steps()}
                        del labels, info
                    except Exception as e:  # noqa: BLE001  (e.g. out of memory at the largest sizes)
                        row = {"D": D, "fg_points": n_fg, "bandwidth": bw, "seeds": seeding, "error": str(e)[:160]}
                        torch.cuda.empty_cache()
                    if n_m == 1 and bw_factor == 1.0 and not args.skip_cpu:
                        row["cpu_baseline"] = sweep_cpu_leg(emb, D, bw, seeding)
                    rows.append(row)
                    print(json.dumps(row), file=sys.stderr, flush=True)
            del emb
            torch.cuda.empty_cache()
    print(json.dumps({"metric": "mean-shift sweep (BASELINE configs[3])", "unit": "ms per point set (threshold -> labels, device)",
                      "sweep": rows}), flush=True)


def sweep_cpu_leg(emb, D, bw, seeding):
    from oracle import mean_shift as oms

    e = emb.cpu().numpy().astype(np.float64)
    X = oms.points_from_embedding(e[:D], e[D] < 0.5)
    if seeding == "grid_binned":
        fit_s, pred_s, n_seeds, k = oms.sklearn_cluster_seconds(X, bw, bin_seeding=True)
        return {"seconds": fit_s + pred_s, "fit_seconds": fit_s, "predict_seconds": pred_s, "n_seeds": n_seeds, "centres": k,
                "kind": "port", "cores": 1, "sample": f"full point set ({len(X)} points), scikit-learn MeanShift(bin_seeding=True)"}
    rng = np.random.default_rng(0)
    subset = X[rng.choice(len(X), 200, replace=False)]
    fit_s, pred_s, _, _ = oms.sklearn_cluster_seconds(X, bw, seeds=subset)
    per_seed = fit_s / 200  # includes the one-off KD-tree build: an upper bound per seed
    return {"seconds_extrapolated": per_seed * len(X) + pred_s, "per_seed_ms": per_seed * 1e3, "kind": "port", "cores": 1,
            "sample": f"200 of {len(X)} seeds climbed by scikit-learn on the full point set, x {len(X)} / 200 (the "
                      "reference seeds with every fit point; a full run would take hours)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-detect", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-multi", action="store_true", help="leave out the multi-GPU inference jobs (profiling runs)")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[3]: mean-shift sweep instead of the headline")
    ap.add_argument("--sweep-max", type=float, default=16, help="largest point set of the sweep, millions (64 = all)")
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
