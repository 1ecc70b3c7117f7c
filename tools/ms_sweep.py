"""BASELINE configs[3]: mean-shift sweep on synthetic embeddings (disc / ball scenes sized to N foreground points).

    python tools/ms_sweep.py [max_points_millions [min_points_millions]]

For each (D, N, bandwidth, seeding) prints one JSON line: time of threshold -> labels on the device, foreground
points labelled per second, the kernels' pair-test counts where known.
seeding = "all" (seeds = every foreground point, reduction_probability 1.0) or "rp0.1".
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellulus_b200 import synthetic  # noqa: E402
from cellulus_b200.detect import detect_embeddings  # noqa: E402

max_m = float(sys.argv[1]) if len(sys.argv) > 1 else 16
min_m = float(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
radius = 10.0
for D in (2, 3):
    for n_m in (1, 4, 16, 64):
        if n_m > max_m or n_m < min_m:
            continue
        shape, n_obj = synthetic.scene_for_points(int(n_m * 1e6), D, radius=radius)
        t0 = time.time()
        emb, _, ids = synthetic.blob_scene(shape, n_obj, radius=radius, seed=0)
        gen_s = time.time() - t0
        d = torch.from_numpy(emb).to(dev)
        n_fg = int((ids > 0).sum())
        del emb, ids
        warmed = False
        for bw_factor in (0.5, 1.0, 2.0):
            bw = bw_factor * radius
            for seeding, rp in (("rp0.1", 0.1), ("all", 1.0)):
                if seeding == "all" and n_m > 4:
                    continue  # all-foreground seeding beyond 4 M points: hours on the CPU reference, minutes here
                try:
                    if not warmed:  # allocator / first-launch effects are not part of the measurement
                        detect_embeddings(d, bandwidth=bw, threshold=0.5, reduction_probability=0.05, rng="philox")
                        warmed = True
                    torch.cuda.synchronize()
                    t0 = time.time()
                    labels, _, _, infos = detect_embeddings(d, bandwidth=bw, threshold=0.5, reduction_probability=rp,
                                                            rng="philox", return_info=True)
                    torch.cuda.synchronize()
                    dt = time.time() - t0
                    info = infos[0]
                    print(json.dumps({"D": D, "shape": list(shape), "fg_points": n_fg, "bandwidth": bw, "seeding": seeding,
                                      "seeds": int(info["n_seeds"]), "centres": int(info["k"]), "method": info["method"],
                                      "seconds": round(dt, 4), "fg_points_per_s": round(n_fg / dt),
                                      "Mpx_per_s": round(float(np.prod(shape)) / dt / 1e6, 1),
                                      "max_iters": int(info["iters"].max().item()), "scene_gen_s": round(gen_s, 1)}), flush=True)
                except Exception as e:  # noqa: BLE001
                    print(json.dumps({"D": D, "fg_points": n_fg, "bandwidth": bw, "seeding": seeding,
                                      "error": repr(e)[:200]}), flush=True)
        del d
        torch.cuda.empty_cache()
