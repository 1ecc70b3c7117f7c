import sys, os, json, torch, numpy as np
sys.path.insert(0, ".")
import bench
from cellulus_b200 import kernels as K, synthetic
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
emb_np, _, ids = synthetic.blob_scene(bench.DET_SHAPE, bench.DET_OBJECTS, radius=bench.DET_RADIUS, seed=0)
emb = torch.from_numpy(emb_np).to(dev)
pts, _, n, _ = K.fg_compact(emb, bench.DET_THR)
fit_pts, n_fit = K.select_points(pts, n, K.bernoulli_flags(n, bench.DET_RP, 0, dev))
lo, hi = K.bounding_box(fit_pts, n_fit)
grid = K.plan_grid(lo, hi, bench.DET_BW)
sorted_pts, cell_start, _ = K.grid_build(fit_pts, n_fit, grid)
ts = []
for _ in range(7):
    seeds = fit_pts.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c, it = K.ms_grid_modes(sorted_pts, n_fit, grid, cell_start, seeds, n_fit, bench.DET_BW); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.path.basename(os.environ.get("CELLULUS_B200_LIB", "default")), "ms_grid_modes ms", round(float(np.median(ts[2:])), 4), int(c.sum()), int(it.sum()))
