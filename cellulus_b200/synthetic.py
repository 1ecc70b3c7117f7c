"""Deterministic synthetic inputs for the embedding-space hot path.

These generators produce the workloads named in SURVEY.md §8(d): loss inputs
(L1-L3), TTA stacks (T1) and disc/ball embedding scenes (D1/D2).  They are
pure numpy so the same bytes are produced on the authoring box and on the
B200 box; neither the oracle nor the CUDA path is involved.
"""

from __future__ import annotations

import numpy as np


def loss_offsets(batch: int, num_dims: int, out_shape, seed: int = 0) -> np.ndarray:
    """`offsets ~ N(0,1)` of shape (B, D, *out_shape), fp32."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((batch, num_dims, *out_shape), dtype=np.float32)


def tta_stack(num_passes: int, num_dims: int, out_shape, seed: int = 0) -> np.ndarray:
    """T noisy predictions (T, D, *S) fp32: a smooth field plus per-pass noise."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((1, num_dims, *out_shape), dtype=np.float32) * 4.0
    noise = rng.standard_normal((num_passes, num_dims, *out_shape), dtype=np.float32)
    return (base + 0.25 * noise).astype(np.float32)


def blob_scene(
    shape,
    num_objects: int,
    radius: float = 10.0,
    offset_sigma: float = 0.5,
    seed: int = 0,
    dtype=np.float32,
):
    """Disc (2D) / ball (3D) scene with object-centric embeddings.

    Returns `(embeddings, centres, instance_ids)`:

    * `embeddings` (D+1, *shape): channel k (k=0 is **x**, the last axis; the
      reference's channel order, `utils/mean_shift.py:16-32`) holds
      `centre_k - coordinate_k + N(0, sigma^2)` inside an object and small noise
      outside; the last channel is the "std" channel: U(0,0.1) inside objects,
      1+U(0,0.1) outside.
    * `centres` (K, D) in (x, y[, z]) order.
    * `instance_ids` int32 (*shape): 0 background, 1..K the painting order.
    """
    shape = tuple(int(s) for s in shape)
    D = len(shape)
    rng = np.random.default_rng(seed)
    # centres in array-axis order (z, y, x); keep objects away from the border
    lo = radius
    centres_axis = np.stack(
        [rng.uniform(lo, s - 1 - lo, size=num_objects) for s in shape], axis=1
    )
    ids = np.zeros(shape, dtype=np.int32)
    r = int(np.ceil(radius))
    for k, c in enumerate(centres_axis):
        sl = tuple(
            slice(max(0, int(np.floor(ci)) - r), min(s, int(np.floor(ci)) + r + 2))
            for ci, s in zip(c, shape)
        )
        grids = np.meshgrid(
            *[np.arange(s.start, s.stop) for s in sl], indexing="ij", sparse=True
        )
        d2 = sum((g - ci) ** 2 for g, ci in zip(grids, c))
        sub = ids[sl]
        sub[(d2 <= radius * radius) & (sub == 0)] = k + 1
    emb = np.empty((D + 1, *shape), dtype=dtype)
    fg = ids > 0
    coords_axis = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij", sparse=True)
    centre_lut = np.concatenate([np.zeros((1, D)), centres_axis], axis=0)
    for ch in range(D):
        axis = D - 1 - ch  # channel 0 = x = last axis
        target = centre_lut[:, axis][ids]
        off = target - coords_axis[axis]
        off = np.where(fg, off, 0.0)
        off = off + rng.normal(0.0, offset_sigma, size=shape)
        emb[ch] = off.astype(dtype)
    u = rng.uniform(0.0, 0.1, size=shape)
    emb[D] = np.where(fg, u, 1.0 + u).astype(dtype)
    centres_xyz = centres_axis[:, ::-1].copy()
    return emb, centres_xyz, ids


def scene_for_points(num_points: int, num_dims: int, radius: float = 10.0, seed: int = 0,
                     fg_fraction: float = 0.2):
    """Pick a canvas / object count so a blob scene has ~`num_points` foreground
    pixels (SURVEY §8d D2 sweep).  Returns `(shape, num_objects)`."""
    vol = np.pi * radius**2 if num_dims == 2 else 4.0 / 3.0 * np.pi * radius**3
    num_objects = max(1, int(round(num_points / vol)))
    total = num_points / fg_fraction
    side = int(np.ceil(total ** (1.0 / num_dims)))
    side = max(side, int(4 * radius))
    return (side,) * num_dims, num_objects


def block_scene(shape, radius, seed, dev):
    """Noise-free embeddings of a jittered-lattice blob scene, built with torch ON THE DEVICE `dev`:
    `(base (D, *shape) float32, foreground (*shape) bool, generator)`; channel 0 = x (last axis)."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    D = len(shape)
    spacing = 2.6 * radius
    cells = [int(np.ceil(s / spacing)) + 2 for s in shape]
    centres = (torch.stack(torch.meshgrid(*[torch.arange(c, device=dev) for c in cells], indexing="ij"), -1).float()
               - 0.5) * spacing + (torch.rand((*cells, D), generator=g, device=dev) - 0.5) * (spacing - 2 * radius)
    coords = torch.stack(torch.meshgrid(*[torch.arange(s, device=dev) for s in shape], indexing="ij"), -1).float()
    ci = torch.floor(coords / spacing + 1.0).long().clamp_(min=0)
    best_d = torch.full(shape, 1e9, device=dev)
    best_c = torch.zeros((*shape, D), device=dev)
    for off in np.ndindex(*(3,) * D):
        idx = [(ci[..., k] + off[k] - 1).clamp_(0, cells[k] - 1) for k in range(D)]
        c = centres[tuple(idx)]
        d = ((coords - c) ** 2).sum(-1)
        closer = d < best_d
        best_d = torch.where(closer, d, best_d)
        best_c = torch.where(closer[..., None], c, best_c)
    fg = best_d <= radius * radius
    base = torch.where(fg[..., None], best_c - coords, torch.zeros_like(coords))
    base = base.flip(-1).movedim(-1, 0).contiguous()  # channel 0 = x (last axis)
    return base, fg, g


def block_stack(shape, radius, T, seed, dev):
    """T noisy predictions (T, D, *shape) of a jittered-lattice blob scene, built with torch ON THE DEVICE `dev`
    (stands for the T test-time-augmentation passes of the U-Net over one scan block, BASELINE configs[4]):
    tight noise inside the objects (per-pixel std 0.02 per channel), unit noise outside."""
    import torch

    base, fg, g = block_scene(shape, radius, seed, dev)
    D = len(shape)
    sigma = torch.where(fg, 0.02, 1.0)[None, None]
    return base[None] + sigma * torch.randn((T, D, *shape), generator=g, device=dev)
