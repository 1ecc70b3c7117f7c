"""Time one library variant (CELLULUS_B200_LIB) of the fused loss on config #2; one line per layout."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cellulus_b200 import kernels as K
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
offsets = torch.randn(bench.B, bench.D, *bench.OUT, device=dev)
anchors, refs = K.sample_pairs(bench.B, (bench.OUT[1], bench.OUT[0]), bench.KAPPA, bench.N_ANCHORS, bench.N_REFS, seed=1, device=dev)
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
tag = os.path.basename(os.environ.get("CELLULUS_B200_LIB", "default"))
res = []
for layout in ["planar", "cl"]:
    off = offsets if layout == "planar" else offsets.contiguous(memory_format=torch.channels_last)
    for bwd in [True, False]:
        res.append(f"{layout}/{'bwd' if bwd else 'fwd'} {timeit(lambda: K.oce_loss_fwd_bwd(off, anchors, refs, bench.TEMP, bench.REGW, want_grad=bwd)):6.1f}")
g = torch.empty_like(offsets)
res.append(f"memset {timeit(lambda: g.zero_()):5.1f}")
print(tag, " | ".join(res), flush=True)
