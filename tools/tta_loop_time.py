"""Infer-mode forward (test-time-augmentation loop of models/unet.py:73-100, T = 32 passes) of the stand-in U-Net on
one scan block: eager loop against the CUDA-graph replay."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellulus_b200.models import get_model  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for nd, crop, fmaps, down in ((2, (252, 252), 24, [(2, 2)]), (2, (512, 512), 24, [(2, 2)]), (3, (40, 100, 100), 12, [(1, 2, 2)])):
    torch.manual_seed(0)
    res = []
    for graph in (False, True):
        model = get_model(1, nd, fmaps, 3, fmaps, down, nd).to(dev).eval()
        model.set_infer(p_salt_pepper=0.01, num_infer_iterations=16, device=dev, cuda_graph=graph)
        raw = torch.rand(1, 1, *crop, device=dev)
        with torch.no_grad():
            for _ in range(2):
                out = model(raw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                out = model(raw)
            torch.cuda.synchronize()
        res.append((time.perf_counter() - t0) / 5 * 1e3)
    print(f"{nd}-D crop {crop} -> {tuple(out.shape[2:])}, fmaps {fmaps}: eager {res[0]:.2f} ms, CUDA graph {res[1]:.2f} ms "
          f"per block of 32 passes ({res[0] / res[1]:.2f}x)", flush=True)
