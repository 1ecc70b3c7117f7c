"""Hill climb of the 272^3 sharded-volume scene on ONE GPU (3.23 M seeds = every foreground voxel, bw 10): every seed to
convergence against the distinct-trajectory form with 1..N merge rounds (device time from CUDA graphs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cellulus_b200 import kernels as K, synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
emb_np, _, _ = synthetic.blob_scene(bench.SHARD_SHAPE, bench.SHARD_OBJECTS, radius=10.0, seed=0)
emb = torch.from_numpy(emb_np).to(dev)
pts, _, n, _ = K.fg_compact(emb, 0.5)
lo, hi = K.bounding_box(pts, n)
grid = K.plan_grid(lo, hi, bench.SHARD_BW)
sorted_pts, cell_start, _ = K.grid_build(pts, n, grid)


def timed(fn):
    ts = []
    for _ in range(3):
        seeds = pts.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(seeds)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[1:])), K.grid_modes_distance_tests(), K.grid_modes_climb_steps()


print(n, "seeds")
print("every seed to convergence: %.2f ms, %d tests, %d evaluations" % timed(
    lambda s: K.ms_grid_modes(sorted_pts, n, grid, cell_start, s, n, bench.SHARD_BW)), flush=True)
for rounds in (1, 2, 4, 8, 16, 30):
    print("merge rounds %2d: %.2f ms, %d tests, %d evaluations" % ((rounds,) + timed(
        lambda s: K.ms_grid_modes_distinct(sorted_pts, n, grid, cell_start, s, n, bench.SHARD_BW, merge_rounds=rounds))), flush=True)
