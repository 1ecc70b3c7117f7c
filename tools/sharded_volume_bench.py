"""ONE large volume, mean-shift seeds sharded over the ranks (last row): every rank compacts its
slab, all-gather of the fit points, each rank climbs its slice of the seeds, all-gather of (mode, count),
replicated centre suppression, labels per slab.  Device-timed, max over ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_volume_bench.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellulus_b200 import kernels as K  # noqa: E402
from cellulus_b200 import sharding, synthetic  # noqa: E402

world, rank, local = (int(os.environ.get(k, d)) for k, d in [("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)

shape, n_obj, radius, bw = (272, 272, 272), 855, 10.0, 10.0  # ~3.6 M foreground voxels, every one a seed
emb, _, _ = synthetic.blob_scene(shape, n_obj, radius=radius, seed=0)
# slab = contiguous z-range of the volume (rank order = raster order)
zs = sharding.shard_items(shape[0], rank, world)
slab = torch.from_numpy(np.ascontiguousarray(emb[:, zs.start:zs.stop])).to(dev)
pts, pix, n_local, _ = K.fg_compact(slab, 0.5)
pts[2, :n_local] += float(zs.start)  # z coordinate of the slab inside the volume (channel 2 = z)
ops = sharding.cuda_ops("grid")
for it in range(3):
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    labels, centres = sharding.sharded_mean_shift(pts, n_local, bw, ops, None, None)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    n = torch.tensor([n_local], device=dev, dtype=torch.int64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(n)
if rank == 0:
    print(json.dumps({"workload": f"one {shape[0]}^3 volume, {int(n.item())} foreground voxels, all of them seeds, bw {bw}; "
                                  "seed-sharded mean-shift with 2 all-gathers", "n_gpus": world, "ms": t.item(),
                      "fg_points_per_s": n.item() / t.item() * 1e3, "centres": int(centres.shape[1])}), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
