"""Time the sampled-loss kernel (library variant from CELLULUS_B200_LIB) at configs[1], both layouts."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cellulus_b200.criterions import GraphedLossStep
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
res = []
for name, fmt in (("cl", torch.channels_last), ("planar", torch.contiguous_format)):
    steps = [GraphedLossStep(torch.randn(bench.B, bench.D, *bench.OUT, device=dev).contiguous(memory_format=fmt), None, None,
                             bench.TEMP, bench.REGW, sampled=dict(kappa=bench.KAPPA, num_anchors=bench.N_ANCHORS,
                                                                  num_references=bench.N_REFS, seed=5 + i,
                                                                  extent_xyz=(bench.OUT[1], bench.OUT[0]))) for i in range(3)]
    for i in range(10): steps[i % 3].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): steps[i % 3].replay()
    e1.record(); torch.cuda.synchronize()
    res.append(f"{name} {e0.elapsed_time(e1) / 200 * 1e3:.1f} us (loss {steps[0].loss.item():.1f})")
print(os.path.basename(os.environ.get("CELLULUS_B200_LIB", "default")), os.environ.get("CB200_SAMPLED_KERNEL", "tiled"), " | ".join(res), flush=True)
