"""Time the explicit-list loss step at configs[1] the way bench.py does (CUDA-graph replay, 3 input sets of 211 MB
visited round-robin) for one library variant / one set of CB200_LOSS_* switches; one line per run.

    python tools/loss_step_time.py [cl cl_multi planar planar_multi cl_i16]
    CB200_LOSS_STAGED=0 python tools/loss_step_time.py planar            # planar offsets gathered in place
    CELLULUS_B200_LIB=variants/x.so python tools/loss_step_time.py       # a library built with other -D switches

`*_multi`: the three steps captured in ONE graph (what bench.py replays); otherwise one graph per step.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cellulus_b200 import kernels as K  # noqa: E402
from cellulus_b200.criterions import GraphedLossStep  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
N_SETS, STEPS = 3, int(os.environ.get("STEPS", "300"))
what = sys.argv[1:] or ["cl", "cl_multi", "planar", "planar_multi", "cl_i16"]
res = []
for name in what:
    fmt = torch.contiguous_format if name.startswith("planar") else torch.channels_last
    cdt = torch.int16 if name.endswith("i16") else torch.int64
    torch.manual_seed(0)
    steps, inputs = [], []
    for i in range(N_SETS):
        off = torch.randn(bench.B, bench.D, *bench.OUT, device=dev).contiguous(memory_format=fmt)
        anc, ref = K.sample_pairs(bench.B, (bench.OUT[1], bench.OUT[0]), bench.KAPPA, bench.N_ANCHORS, bench.N_REFS,
                                  seed=1234 + i, device=dev, dtype=cdt)
        inputs.append((off, anc, ref))
        steps.append(GraphedLossStep(off, anc, ref, bench.TEMP, bench.REGW))
    if "multi" in name:  # the N_SETS steps captured in ONE graph: kernel -> kernel edges instead of graph -> graph
        class Multi:
            def __init__(self):
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.outs = [K.oce_loss_fwd_bwd(o, a, r, bench.TEMP, bench.REGW) for o, a, r in inputs]
                self.loss, self.grad = self.outs[0][0][0], self.outs[0][1]
            def replay(self):
                self.graph.replay()
        m = Multi()
        class Third:  # one replay per N_SETS "steps"
            def __init__(self, i): self.i = i; self.loss, self.grad = m.loss, m.grad
            def replay(self):
                if self.i == 0: m.replay()
        steps = [Third(i) for i in range(N_SETS)]
    best = 1e9
    for _ in range(3):
        for i in range(20):
            steps[i % N_SETS].replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(STEPS):
            steps[i % N_SETS].replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / STEPS * 1e3)
    g = steps[0].grad.double()
    res.append(f"{name} {best:.2f} us (loss {steps[0].loss.item():.3f} |g| {g.abs().sum().item():.6f} "
               f"gsum {g.sum().item():.6f})")
    del steps
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("CB200_") or k == "CELLULUS_B200_LIB")
print(f"[{tag or 'default'}] " + " | ".join(res), flush=True)
