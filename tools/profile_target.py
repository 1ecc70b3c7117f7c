"""Small driver for ncu captures: a few fused-loss steps (config #2) and one detect volume (config #3)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cellulus_b200 import kernels as K  # noqa: E402
from cellulus_b200 import synthetic  # noqa: E402
from cellulus_b200.detect import detect_embeddings  # noqa: E402
from cellulus_b200.models import tta_aggregate  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if os.environ.get("CB200_PROFILE_SMALL"):  # compute-sanitizer runs: same code paths, sizes it finishes in seconds
    bench.B, bench.OUT = 2, (124, 124)
    bench.N_ANCHORS, bench.N_REFS = int(0.1 * 104 * 104), 31
    bench.DET_SHAPE, bench.DET_OBJECTS = (32, 64, 64), 25
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
if what in ("all", "loss"):
    offsets = torch.randn(bench.B, bench.D, *bench.OUT, device=dev)
    anchors, refs = K.sample_pairs(bench.B, (bench.OUT[1], bench.OUT[0]), bench.KAPPA, bench.N_ANCHORS, bench.N_REFS,
                                   seed=1, device=dev)
    off_cl = offsets.contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        K.oce_loss_fwd_bwd(offsets, anchors, refs, bench.TEMP, bench.REGW)
    for _ in range(3):
        K.oce_loss_fwd_bwd(off_cl, anchors, refs, bench.TEMP, bench.REGW)
    torch.cuda.synchronize()
if what in ("all", "sampled"):
    offsets = torch.randn(bench.B, bench.D, *bench.OUT, device=dev)
    off_cl = offsets.contiguous(memory_format=torch.channels_last)
    for o in (off_cl, offsets):
        for i in range(3):
            K.oce_loss_sampled(o, bench.KAPPA, bench.N_ANCHORS, bench.N_REFS, 5, i, bench.TEMP, bench.REGW,
                               extent_xyz=(bench.OUT[1], bench.OUT[0]))
    torch.cuda.synchronize()
if what in ("all", "tta"):
    stack = torch.randn(32, 2, 496, 496, device=dev)
    for _ in range(3):
        tta_aggregate(stack)
    torch.cuda.synchronize()
if what in ("all", "detect"):
    emb, _, _ = synthetic.blob_scene(bench.DET_SHAPE, bench.DET_OBJECTS, radius=bench.DET_RADIUS, seed=0)
    d = torch.from_numpy(emb).to(dev)
    for _ in range(2):
        detect_embeddings(d, bandwidth=bench.DET_BW, threshold=bench.DET_THR, reduction_probability=bench.DET_RP,
                          rng="philox")
    torch.cuda.synchronize()

if what in ("all", "brute", "greedy"):
    from cellulus_b200.utils.mean_shift import segment_embeddings_device
    import time
    emb, _, _ = synthetic.blob_scene((64, 256, 256), 200, radius=bench.DET_RADIUS, seed=1)
    d = torch.from_numpy(emb).to(dev)
    if what in ("all", "brute"):
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            labels, info = segment_embeddings_device(d, bench.DET_BW, bench.DET_THR, 0.1, rng="philox", method="brute")
            torch.cuda.synchronize()
            iters = info["iters"][: info["n_seeds"]].double()
            pair_tests = float((iters + 1).sum().item()) * info["n_fit"]
            print(f"brute: n_fit {info['n_fit']} seeds {info['n_seeds']} k {info['k']} total {time.time() - t0:.4f} s "
                  f"pair tests {pair_tests:.3e}")
    if what in ("all", "greedy"):
        mask = (d[3] < 0.5).to(torch.uint8)
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            inst, n_obj, n_iter = K.greedy_cluster(d, mask, bench.DET_BW, 100)
            torch.cuda.synchronize()
            print(f"greedy: fg {int(mask.sum())} objects {n_obj} seeds tried {n_iter} total {time.time() - t0:.4f} s")

if what in ("all", "post"):
    from cellulus_b200.evaluate import compute_pairwise_IoU
    from cellulus_b200.segment import nucleus

    seg_np, raw_np = bench._label_scene((2048, 2048), 1200, 26, seed=3)
    seg = torch.from_numpy(seg_np).to(dev)
    for _ in range(2):
        K.grow_shrink_(seg.clone(), 3, 6)
        K.size_filter_(seg.clone(), 25)
        nucleus(seg_np, raw_np)
        compute_pairwise_IoU(seg_np.astype(np.uint16), np.roll(seg_np, (3, -2), axis=(0, 1)).astype(np.uint16))
    torch.cuda.synchronize()
