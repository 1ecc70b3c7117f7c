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
  detect        secondary metric: mean-shift detection Mpx/s on BASELINE config #3
                (128 x 256 x 256 volume, 3-D embeddings), with its own cpu_baseline

`--impl reference` times the oracle port (torch CPU, all host threads) on rank 0.
Nothing here reads /root/reference.
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
# dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu capture (profiles/), or None
TRAFFIC_NCU = 208_558_848  # profiles/r01_loss_ncu_full_summary.txt: 204.45 MB read + 4.11 MB written
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
    line = {
        "impl": "reference", "metric": "OCELoss fwd+bwd px/s", "value": value, "unit": "px/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "px/s", "cores": cores, "kind": "port", "sample": sample,
                         "pairs_per_s": batch * P / t},
        "e2e": {"value": value, "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- B200 arm
def bench_detect(dev, windows):
    """Secondary metric: threshold -> labels on BASELINE config #3, inputs resident in HBM."""
    from cellulus_b200 import kernels as K
    from cellulus_b200 import synthetic
    from cellulus_b200.detect import detect_embeddings

    emb_np, _, ids = synthetic.blob_scene(DET_SHAPE, DET_OBJECTS, radius=DET_RADIUS, seed=0)
    emb = torch.from_numpy(emb_np).to(dev)
    n_vox = int(np.prod(DET_SHAPE))
    fg = int((ids > 0).sum())
    kw = dict(bandwidth=DET_BW, threshold=DET_THR, reduction_probability=DET_RP, rng="philox")
    for _ in range(2):
        labels, _, _, infos = detect_embeddings(emb, return_info=True, **kw)
    for _ in range(3):  # warm-up of the path that is timed (one C-ABI call per volume; its scratch arena grows here)
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
    # e2e: pinned host volume in, uint16 labels out
    host = torch.from_numpy(emb_np).pin_memory()
    out_host = torch.empty((1, *DET_SHAPE), dtype=torch.uint16).pin_memory()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        d = host.to(dev, non_blocking=True)
        labels, _, _ = detect_embeddings(d, **kw)
        out_host.copy_(labels, non_blocking=True)
        torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / 3
    info = infos[0]
    return {
        "metric": "detect Mpx/s (mean-shift)", "value": n_vox / ms / 1e3, "unit": "Mpx/s", "ms_per_volume": ms,
        "config": {"workload": "configs[2]: 128x256x256 volume, 3-D embeddings, 400 balls r=10, bw=7, threshold=0.5, "
                               "reduction_probability=0.1, seeds=all fit points",
                   "foreground_voxels": fg, "fit_points": int(info["n_fit"]), "centres": int(info["k"]),
                   "method": info["method"]},
        "e2e": {"value": n_vox / e2e_s / 1e6, "unit": "Mpx/s", "h2d_bytes_per_step": host.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * 2},
        "abi_calls_per_volume": int(calls),
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
        "nucleus": {"ms": t_nuc * 1e3, "Mpx/s": n / t_nuc / 1e6, "note": "numpy in/out incl. the host<->device copies (per-instance Otsu runs on the device)"},
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


def cpu_detect_baseline():
    """Oracle port of `mean_shift_segmentation` (scikit-learn MeanShift, hill climb on ONE core as shipped)
    on a bounded sub-volume of the same scene family."""
    from cellulus_b200 import synthetic
    from oracle import mean_shift as oms

    shape, objects = (24, 96, 96), 11  # same object density as config #3
    emb, _, _ = synthetic.blob_scene(shape, objects, radius=DET_RADIUS, seed=0)
    emb64 = emb.astype(np.float64)
    np.random.seed(0)
    t0 = time.perf_counter()
    labels = oms.mean_shift_segmentation(emb64[np.newaxis, :3].copy(), emb64[3], DET_BW, 0, DET_RP, DET_THR, None)
    dt = time.perf_counter() - t0
    n = int(np.prod(shape))
    return {"value": n / dt / 1e6, "unit": "Mpx/s", "cores": 1, "kind": "port",
            "sample": f"{shape[0]}x{shape[1]}x{shape[2]} sub-volume, {objects} balls, {int((labels > 0).sum())} fg voxels, "
                      f"{dt:.1f} s; sklearn hill climb is single-core (n_jobs=None), predict uses OpenMP"}


def run_b200(args):
    from cellulus_b200 import kernels as K
    from cellulus_b200.criterions import GraphedLossStep, oce_loss_fused

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

    sampler = ClockSampler(local) if rank == 0 else None
    windows = []

    # inputs resident in HBM: every rank owns its own batches (sharded by batch).  N_SETS distinct input
    # sets are visited round-robin so that no step finds its 180 MB of coordinate lists in the 126 MB L2.
    N_SETS = 3
    torch.manual_seed(rank)

    def make_steps(memory_format, dtype=torch.float32, coord_dtype=torch.int64):
        steps = []
        for i in range(N_SETS):
            off = torch.randn(B, D, *OUT, device=dev).to(dtype).contiguous(memory_format=memory_format)
            anc, ref = K.sample_pairs(B, (OUT[1], OUT[0]), KAPPA, N_ANCHORS, N_REFS, seed=1234 + 17 * rank + i,
                                      device=dev, dtype=coord_dtype)
            steps.append(GraphedLossStep(off, anc, ref, TEMP, REGW))
        return steps

    def timed(steps, n_steps, n_warm):
        for i in range(n_warm):
            steps[i % N_SETS].replay()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for i in range(n_steps):
            steps[i % N_SETS].replay()
        e1.record()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        windows.append((w0, time.time()))
        t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return t_ms.item() / n_steps

    warm = max(args.warmup, 3)
    # (1) headline: channels-last offsets (the layout this framework keeps the U-Net output in)
    steps_cl = make_steps(torch.channels_last)
    c0 = K.launch_counter["calls"]
    ms_per_step = timed(steps_cl, args.steps, warm)
    launches = (K.launch_counter["calls"] - c0 - warm) * 2  # zero-fill + fused kernel per replay
    value = world * N_PX / (ms_per_step * 1e-3)
    loss_value = steps_cl[0].loss.item()
    # (2) same op on the planar NCHW tensor the reference's model emits
    steps_pl = make_steps(torch.contiguous_format)
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

    # (3) end to end through the public API from pinned host buffers
    h_off = offsets.detach().cpu().contiguous().pin_memory()
    h_anc, h_ref = anchors.cpu().pin_memory(), refs.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        o = h_off.to(dev, non_blocking=True).requires_grad_(True)
        a = h_anc.to(dev, non_blocking=True)
        r = h_ref.to(dev, non_blocking=True)
        loss, _, _ = oce_loss_fused(o, a, r, TEMP, REGW)
        loss.backward()
        return loss.item()  # device -> host read of the step's result (train.py:180)

    for _ in range(2):
        e2e_step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    w0 = time.time()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev)
    windows.append((w0, time.time()))
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * N_PX / e2e_s.item()

    # (3b) the same step when the pair lists are drawn ON THE DEVICE (cb200_sample_pairs replaces the
    # DataLoader-side sampler of zarr_dataset.py:198-242): only the offsets cross PCIe
    def e2e_sampled_step(i):
        o = h_off.to(dev, non_blocking=True).requires_grad_(True)
        a, r = K.sample_pairs(B, (OUT[1], OUT[0]), KAPPA, N_ANCHORS, N_REFS, seed=99, sequence=i, device=dev,
                                  dtype=torch.int16)  # the training loop's list format
        loss, _, _ = oce_loss_fused(o, a, r, TEMP, REGW)
        loss.backward()
        return loss.item()

    for i in range(2):
        e2e_sampled_step(i)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_sampled_step(i)
    torch.cuda.synchronize(dev)
    e2e_sampled_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev)
    if dist is not None:
        dist.all_reduce(e2e_sampled_s, op=dist.ReduceOp.MAX)

    detect = None
    cpu_base = None
    if rank == 0:
        if not args.skip_detect:
            detect = bench_detect(dev, windows)
            detect["tta_aggregate"] = bench_tta(dev, windows)
            detect["post_processing"] = bench_post(dev, with_cpu=(world == 1 and not args.skip_cpu))
        if world == 1 and not args.skip_cpu:
            t_cpu = cpu_loss_step_time(B, 3, 1)
            cpu_base = {"value": N_PX / t_cpu, "unit": "px/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"full configs[1] batch ({B * P} pairs), 3 steps after 1 warm-up, {t_cpu * 1e3:.0f} ms/step",
                        "pairs_per_s": B * P / t_cpu}
            if detect is not None:
                detect["cpu_baseline"] = cpu_detect_baseline()
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
                       "step": "CUDA-graph replay of zero-fill + fused gather/loss/backward kernel",
                       "sharding": "by batch, one batch per rank, no data-path collective"},
            "pairs_per_s": world * B * P / (ms_per_step * 1e-3),
            "loss_value_check": loss_value,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_NCU, "peak_source": peak_src,
                         "kernel": "oce_loss_fused_kernel<2,int64,f32,bwd,channels_last> (+ zero_fill_kernel)",
                         "kernel_ms": ms_per_step, "algorithmic_bytes_per_launch": ALGO_BYTES,
                         "note": "duration is the whole step: the 15.7 MB gradient zero-fill is included"},
            "planar": {"value": world * N_PX / (ms_planar * 1e-3), "unit": "px/s", "ms_per_step": ms_planar,
                       "offsets_layout": "planar NCHW (what the reference's model emits)",
                       "roofline": {"bound": "hbm", "achieved": achieved_planar, "peak": peak, "unit": "GB/s",
                                    "frac": achieved_planar / peak}},
            "bf16_offsets": {"value": world * N_PX / (ms_bf16 * 1e-3), "unit": "px/s", "ms_per_step": ms_bf16,
                             "offsets_layout": "channels_last, bf16 storage, fp32 arithmetic, fp32 gradient",
                             "roofline": {"bound": "hbm", "achieved": algo_bf16 / (ms_bf16 * 1e-3) / 1e9,
                                          "peak": peak, "unit": "GB/s",
                                          "frac": algo_bf16 / (ms_bf16 * 1e-3) / 1e9 / peak,
                                          "algorithmic_bytes_per_launch": algo_bf16}},
            "i16_pairs": {"value": world * N_PX / (ms_i16 * 1e-3), "unit": "px/s", "ms_per_step": ms_i16,
                          "note": "channels_last fp32 offsets, int16 coordinate lists as drawn by the device pair "
                                  "sampler (the training loop's format; the reference's lists are int64)",
                          "roofline": {"bound": "hbm", "achieved": algo_i16 / (ms_i16 * 1e-3) / 1e9, "peak": peak,
                                       "unit": "GB/s", "frac": algo_i16 / (ms_i16 * 1e-3) / 1e9 / peak,
                                       "algorithmic_bytes_per_launch": algo_i16}},
            "fused_sampling": {"value": world * N_PX / (ms_sampled * 1e-3), "unit": "px/s", "ms_per_step": ms_sampled,
                               "pairs_per_s": world * B * P / (ms_sampled * 1e-3),
                               "planar_ms_per_step": ms_sampled_planar,
                               "note": "cb200_oce_loss_sampled: pairs drawn inside the kernel (device pair stream), no "
                                       "coordinate list in HBM; replaces sampler + list kernel of the training step",
                               "roofline": {"bound": "hbm", "achieved": algo_sampled / (ms_sampled * 1e-3) / 1e9,
                                            "peak": peak, "unit": "GB/s",
                                            "frac": algo_sampled / (ms_sampled * 1e-3) / 1e9 / peak,
                                            "algorithmic_bytes_per_launch": algo_sampled,
                                            "note": "31.5 MB: at this size the kernel is issue / L2-latency bound "
                                                    "(SURVEY 8d), the fraction is reported for completeness"}},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": "px/s",
                    "h2d_bytes_per_step": int(h_off.numel() * 4 + h_anc.numel() * 8 + h_ref.numel() * 8),
                    "d2h_bytes_per_step": 4, "steps": e2e_steps, "ms_per_step": e2e_s.item() * 1e3,
                    "with_device_pair_sampler": {
                        "value": world * N_PX / e2e_sampled_s.item(), "unit": "px/s",
                        "ms_per_step": e2e_sampled_s.item() * 1e3, "h2d_bytes_per_step": int(h_off.numel() * 4),
                        "note": "int16 pair lists drawn on the device (same distribution as the reference "
                                "sampler), sampling kernel inside the timed step"}},
            "gpu_launches": int(launches * world),
            "clocks": clocks,
            "detect": detect,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-detect", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
