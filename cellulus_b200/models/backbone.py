"""Plain-PyTorch valid-convolution U-Net: the stand-in for `funlib.learn.torch.models.UNet`
(an un-vendored git dependency of the reference, `pyproject.toml:30`, constructed at `models/unet.py:24-51`).

NOT the product ("the U-Net backbone stays in PyTorch", north_star) -- it exists so that `train()` /
`infer()` run end to end and so that checkpoints keep the reference's state-dict layout:
`l_conv.{level}.conv_pass.{0,2,4,6}`, `l_down.{level}.down`, `r_up.0.{level}.up`,
`r_conv.0.{level}.conv_pass.{0,2,4,6}`.  Geometry: every conv pass is [3,1,1,3] valid convolutions (-4 px
per axis at its resolution), max-pool downsampling, nearest-neighbour upsampling, the upsampled map is cropped
so the remaining valid convolutions stay aligned with the stride ("crop to factor"), the skip connection is
centre-cropped and concatenated in front.
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn


def _conv(nd):
    return nn.Conv2d if nd == 2 else nn.Conv3d


class ConvPass(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_sizes: Sequence[Tuple[int, ...]]):
        super().__init__()
        nd = len(kernel_sizes[0])
        layers = []
        for k in kernel_sizes:
            layers += [_conv(nd)(in_channels, out_channels, k, padding=0), nn.ReLU()]
            in_channels = out_channels
        self.conv_pass = nn.Sequential(*layers)

    def forward(self, x):
        return self.conv_pass(x)


class Downsample(nn.Module):
    def __init__(self, factor: Tuple[int, ...]):
        super().__init__()
        self.factor = tuple(factor)
        pool = nn.MaxPool2d if len(factor) == 2 else nn.MaxPool3d
        self.down = pool(self.factor, stride=self.factor)

    def forward(self, x):
        for s, f in zip(x.shape[2:], self.factor):
            if s % f:
                raise RuntimeError(f"Can not downsample shape {tuple(x.shape)} with factor {self.factor}")
        return self.down(x)


def _centre_crop(x, spatial):
    offs = [(a - b) // 2 for a, b in zip(x.shape[2:], spatial)]
    return x[(slice(None), slice(None)) + tuple(slice(o, o + s) for o, s in zip(offs, spatial))]


class Upsample(nn.Module):
    def __init__(self, factor, crop_factor, next_conv_kernel_sizes):
        super().__init__()
        self.crop_factor = tuple(crop_factor)
        self.conv_crop = tuple(sum(k[d] - 1 for k in next_conv_kernel_sizes) for d in range(len(factor)))
        self.up = nn.Upsample(scale_factor=tuple(float(f) for f in factor), mode="nearest")

    def forward(self, skip, x):
        g = self.up(x)
        target = tuple(int(math.floor((s - c) / f)) * f + c
                       for s, c, f in zip(g.shape[2:], self.conv_crop, self.crop_factor))
        if target != tuple(g.shape[2:]):
            g = _centre_crop(g, target)
        return torch.cat([_centre_crop(skip, g.shape[2:]), g], dim=1)


class UNet(nn.Module):
    def __init__(self, in_channels, num_fmaps, fmap_inc_factor, downsample_factors: List[Tuple[int, ...]],
                 num_fmaps_out, kernel_size_down, kernel_size_up, **_):
        super().__init__()
        levels = len(downsample_factors) + 1
        fm = [num_fmaps * fmap_inc_factor**level for level in range(levels)]
        self.l_conv = nn.ModuleList(
            ConvPass(in_channels if level == 0 else fm[level - 1], fm[level], kernel_size_down[level])
            for level in range(levels))
        self.l_down = nn.ModuleList(Downsample(f) for f in downsample_factors)
        crop_factors, prod = [], None
        for f in downsample_factors:
            prod = list(f) if prod is None else [a * b for a, b in zip(f, prod)]
            crop_factors.append(prod)
        self.r_up = nn.ModuleList([nn.ModuleList(
            Upsample(downsample_factors[level], crop_factors[level], kernel_size_up[level])
            for level in range(levels - 1))])
        self.r_conv = nn.ModuleList([nn.ModuleList(
            ConvPass(fm[level] + fm[level + 1], num_fmaps_out if level == 0 else fm[level], kernel_size_up[level])
            for level in range(levels - 1))])
        self.levels = levels

    def forward(self, x):
        skips = []
        for level in range(self.levels):
            x = self.l_conv[level](x)
            if level < self.levels - 1:
                skips.append(x)
                x = self.l_down[level](x)
        for level in reversed(range(self.levels - 1)):
            x = self.r_conv[0][level](self.r_up[0][level](skips[level], x))
        return x
