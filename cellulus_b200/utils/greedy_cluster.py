"""`Cluster2d` / `Cluster3d` of `cellulus/utils/greedy_cluster.py` on the B200 cooperative kernel
(`cb200_greedy_cluster`): same constructor and `cluster(prediction, bandwidth, min_object_size, seed_thresh,
min_unclustered_sum)` signature, numpy prediction in, int16 instance map out."""

from __future__ import annotations

import numpy as np
import torch

from cellulus_b200 import kernels as K


class _Cluster:
    ndim = 2

    def __init__(self, *extent_and_mask, fg_mask=None, device="cuda"):
        # Cluster2d(width, height, fg_mask, device) / Cluster3d(width, height, depth, fg_mask, device)
        args = list(extent_and_mask)
        if fg_mask is None:
            fg_mask = args.pop(self.ndim)
            if len(args) > self.ndim:
                device = args.pop(self.ndim)
        self.device = torch.device(device)
        self.fg_mask = torch.from_numpy(np.ascontiguousarray(fg_mask).astype(np.uint8)).to(self.device)

    def cluster(self, prediction, bandwidth, min_object_size, seed_thresh=0.9, min_unclustered_sum=0):
        pred = torch.from_numpy(np.ascontiguousarray(prediction)).to(self.device)
        with torch.cuda.device(self.device):
            instance_map, _, _ = K.greedy_cluster(pred, self.fg_mask, bandwidth, min_object_size, seed_thresh,
                                                  min_unclustered_sum)
        return instance_map.cpu()


class Cluster2d(_Cluster):
    ndim = 2

    def __init__(self, width, height, fg_mask, device):
        super().__init__(fg_mask=fg_mask, device=device)


class Cluster3d(_Cluster):
    ndim = 3

    def __init__(self, width, height, depth, fg_mask, device):
        super().__init__(fg_mask=fg_mask, device=device)
