"""How long does clearing the 15.7 MB gradient take?  Times (CUDA-graph replays, events) of
torch's fill kernel, cudaMemsetAsync and the library's zero-fill + an empty fused launch (P = 0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cellulus_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
bufs = [torch.empty(bench.B, bench.D, *bench.OUT, device=dev) for _ in range(12)]  # 12 x 15.7 MB > L2
off = torch.randn(bench.B, bench.D, *bench.OUT, device=dev).contiguous(memory_format=torch.channels_last)
empty = torch.empty((bench.B, 0, 2), dtype=torch.int64, device=dev)


def graph_time(fn, reps=20):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(len(bufs)):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * len(bufs)) * 1e3


print("torch .zero_()            %.2f us" % graph_time(lambda i: bufs[i].zero_()))
print("torch .fill_(0) via copy  %.2f us" % graph_time(lambda i: bufs[i].copy_(bufs[(i + 1) % len(bufs)])), "(15.7 MB copy, for scale)")
keep = []
print("lib zero-fill + empty fused launch (P=0)  %.2f us" % graph_time(
    lambda i: keep.append(K.oce_loss_fwd_bwd(off, empty, empty, bench.TEMP, bench.REGW))))
