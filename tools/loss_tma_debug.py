"""A/B the TMA-pipelined loss kernel against the direct kernel on several sizes; report mismatches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellulus_b200 import _cabi, kernels as K
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
lib = _cabi.load()
for (B, S, na, nr) in [(1, (64, 64), 50, 31), (2, (128, 128), 900, 31), (8, (496, 496), 22657, 31), (3, (200, 180), 5000, 7)]:
    off = torch.randn(B, 2, *S, device=dev)
    a, r = K.sample_pairs(B, (S[1], S[0]), 10.0, na, nr, seed=3, device=dev)
    for fmt in ["planar", "cl"]:
        o = off if fmt == "planar" else off.contiguous(memory_format=torch.channels_last)
        lib.cb200_oce_loss_set_variant(0)
        out0, g0 = K.oce_loss_fwd_bwd(o, a, r, 10.0, 1e-5)
        lib.cb200_oce_loss_set_variant(1)
        worst = 0.0
        for rep in range(20):
            out1, g1 = K.oce_loss_fwd_bwd(o, a, r, 10.0, 1e-5)
            d = (g1 - g0).abs().max().item() / g0.abs().max().item()
            worst = max(worst, d)
        nbad = ((g1 - g0).abs() > 1e-4 * g0.abs().max()).sum().item()
        print(B, S, na * nr, fmt, "loss", out0[0].item(), out1[0].item(), "worst grad relerr over 20 reps", worst, "bad px", nbad)
