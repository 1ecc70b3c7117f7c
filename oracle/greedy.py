"""Oracle: greedy seed-and-grow clustering (torch CPU port).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows `cellulus/utils/greedy_cluster.py:46-120` (2-D,
fp32) and `:176-253` (3-D, dtype of the prediction).  Pinned against the reference's own classes by
`tests/golden/make_golden.py` (`greedy.npz`).
"""

from __future__ import annotations

import numpy as np
import torch


def greedy_cluster(prediction: np.ndarray, fg_mask: np.ndarray, bandwidth, min_object_size, seed_thresh=0.9,
                   min_unclustered_sum=0):
    D = prediction.shape[0] - 1
    spatial = prediction.shape[1:]
    pred = torch.from_numpy(prediction)
    if D == 2:
        pred = pred.float()  # :84
    grids = torch.meshgrid(*[torch.linspace(0, s - 1, s) for s in spatial], indexing="ij")
    coords = torch.stack(list(reversed(grids)), 0)  # x, y[, z]
    fg = torch.from_numpy(fg_mask[np.newaxis])
    embeddings = pred[0:D] + coords
    seed_map = pred[D:D + 1]
    seed_map = (seed_map - seed_map.max()) / (seed_map.min() - seed_map.max())
    instance_map = torch.zeros(*spatial).short()
    count = 1
    emb_m = embeddings[fg.expand_as(embeddings)].view(D, -1)
    seed_m = seed_map[fg].view(1, -1)
    n = int(fg.sum())
    unclustered = torch.ones(n).short()
    inst_m = torch.zeros(n).short()
    tried = 0
    while unclustered.sum() > min_unclustered_sum:
        seed = (seed_m * unclustered.float()).argmax().item()
        seed_score = (seed_m * unclustered.float()).max().item()
        if seed_score < seed_thresh:
            break
        tried += 1
        center = emb_m[:, seed:seed + 1]
        unclustered[seed] = 0
        dist = torch.exp(-1 * torch.sum(torch.pow(emb_m - center, 2) / (2 * (bandwidth**2)), 0))
        proposal = (dist > 0.5).squeeze()
        if proposal.sum() > min_object_size:
            if unclustered[proposal].sum().float() / proposal.sum().float() > 0.5:
                inst_m[proposal.squeeze()] = count
                count += 1
        unclustered[proposal] = 0
    instance_map[fg.squeeze()] = inst_m
    return instance_map.numpy(), count - 1, tried
