"""Oracle: anchor / reference pixel-pair sampler (numpy, global RNG).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows
`cellulus/datasets/zarr_dataset.py:94-98,177-251`, consuming `np.random` in
the same call order so a seeded run reproduces the reference bit for bit.
"""

from __future__ import annotations

import numpy as np


def output_shape_of(crop_size):
    """`zarr_dataset.py:94`: the hard-coded `crop - 16`."""
    return tuple(int(c - 16) for c in crop_size)


def unbiased_shape_of(output_shape, kappa):
    """`zarr_dataset.py:96-98`."""
    return tuple(int(o - (2 * kappa)) for o in output_shape)


def num_anchors(density, unbiased_shape):
    """`zarr_dataset.py:244-245` (first two dims only, even in 3-D)."""
    return int(density * unbiased_shape[0] * unbiased_shape[1])


def num_references(density, kappa):
    """`zarr_dataset.py:247-248` (kappa^2 * pi even in 3-D)."""
    return int(density * kappa**2 * np.pi)


def sample_offsets_within_radius(radius, number_offsets, num_spatial_dims):
    """`zarr_dataset.py:177-196`: rejection sample integer offsets in the open
    disc/ball `sum o^2 < radius^2` minus the origin; redraw everything if short."""
    if num_spatial_dims == 2:
        ox = np.random.randint(-radius, radius + 1, size=2 * number_offsets)
        oy = np.random.randint(-radius, radius + 1, size=2 * number_offsets)
        offsets = np.stack((ox, oy), axis=1)
    elif num_spatial_dims == 3:
        ox = np.random.randint(-radius, radius + 1, size=3 * number_offsets)
        oy = np.random.randint(-radius, radius + 1, size=3 * number_offsets)
        oz = np.random.randint(-radius, radius + 1, size=3 * number_offsets)
        offsets = np.stack((ox, oy, oz), axis=1)
    else:
        raise ValueError("num_spatial_dims must be 2 or 3")
    in_circle = (offsets**2).sum(axis=1) < radius**2
    offsets = offsets[in_circle]
    not_zero = np.absolute(offsets).sum(axis=1) > 0
    offsets = offsets[not_zero]
    if len(offsets) < number_offsets:
        return sample_offsets_within_radius(radius, number_offsets, num_spatial_dims)
    return offsets[:number_offsets]


def sample_coordinates(output_shape, kappa, density, num_spatial_dims):
    """`zarr_dataset.py:198-242`: anchors uniform in [kappa, out-kappa]
    (inclusive), each repeated `num_references` times consecutively, references
    = anchor + offset.  Columns are (x, y[, z]); x is drawn from
    `output_shape[0]` (quirk Q4)."""
    unbiased = unbiased_shape_of(output_shape, kappa)
    n_anchor = num_anchors(density, unbiased)
    n_ref = num_references(density, kappa)
    cols = []
    for d in range(num_spatial_dims):
        cols.append(np.random.randint(kappa, output_shape[d] - kappa + 1, size=n_anchor))
    anchors = np.stack(cols, axis=1)
    anchor_samples = np.repeat(anchors, n_ref, axis=0)
    offsets = sample_offsets_within_radius(kappa, len(anchor_samples), num_spatial_dims)
    reference_samples = anchor_samples + offsets
    return anchor_samples, reference_samples
