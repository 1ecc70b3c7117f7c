"""How much of the fused-loss time is the scattered reference gather?  Same pair count, shrinking window."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cellulus_b200 import kernels as K
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
offsets = torch.randn(bench.B, bench.D, *bench.OUT, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=8, inner=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner): fn()
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / inner)
    return float(np.median(ts)) * 1e3
for kappa in [10.0, 5.0, 2.0]:
    anchors, refs = K.sample_pairs(bench.B, (bench.OUT[1], bench.OUT[0]), kappa, bench.N_ANCHORS, bench.N_REFS, seed=1, device=dev)
    # keep anchors in the same range as the kappa=10 case irrespective of kappa
    for name, r in [("window", refs), ("refs==anchors", anchors)]:
        res = []
        for layout in ["planar", "cl"]:
            off = offsets if layout == "planar" else offsets.contiguous(memory_format=torch.channels_last)
            for bwd in [True, False]:
                res.append(f"{layout}/{'bwd' if bwd else 'fwd'} {timeit(lambda: K.oce_loss_fwd_bwd(off, anchors, r, bench.TEMP, bench.REGW, want_grad=bwd)):6.1f}")
        print(f"kappa {kappa:4.1f} {name:14s}", " | ".join(res), flush=True)
# sorted anchors: neighbouring runs share lines too
anchors, refs = K.sample_pairs(bench.B, (bench.OUT[1], bench.OUT[0]), 10.0, bench.N_ANCHORS, bench.N_REFS, seed=1, device=dev)
