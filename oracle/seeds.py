"""Oracle: seed finder of the `use_seeds=True` branch (numpy + scipy).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows `cellulus/detect.py:128-132`:
`seeds = flip(peak_local_max(-gaussian_filter(norm(centred[:-1], axis=0), sigma=2)), 1)`.

`scipy.ndimage.gaussian_filter` is installed and is used directly; `gaussian_filter_restated` spells out
its arithmetic (what the CUDA kernel reproduces bit for bit) and is checked against it.
`skimage.feature.peak_local_max` is NOT installed and its source is not under /root/reference:
restated from the published algorithm -- **parity unpinned**.
"""

from __future__ import annotations

import numpy as np
from scipy import ndimage


def gaussian_weights(sigma: float, truncate: float = 4.0):
    """scipy `_gaussian_kernel1d` (order 0): returns `(weights[radius..0..radius] , radius)`."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x**2)
    return phi / phi.sum(), radius


def gaussian_filter_restated(image: np.ndarray, sigma: float) -> np.ndarray:
    """`ndimage.gaussian_filter(image, sigma)` (mode="reflect") with scipy's exact operation order:
    per axis, tmp = x[0]*w[0]; for j = radius..1: tmp += (x[-j] + x[+j]) * w[j]."""
    w, radius = gaussian_weights(sigma)
    out = np.asarray(image, dtype=np.float64)
    for axis in range(out.ndim):
        n = out.shape[axis]
        src = np.moveaxis(out, axis, 0)
        idx = np.arange(n)

        def refl(i):
            i = np.mod(i, 2 * n)
            return np.where(i < n, i, 2 * n - 1 - i)

        tmp = src[idx] * w[radius]
        for j in range(radius, 0, -1):
            tmp = tmp + (src[refl(idx - j)] + src[refl(idx + j)]) * w[radius + j]
        out = np.moveaxis(tmp, 0, axis)
    return out


def peak_local_max(image: np.ndarray) -> np.ndarray:
    """scikit-image `peak_local_max(image)` with default arguments (min_distance=1, exclude_border=True):
    pixels equal to the 3^D maximum filter (mode="nearest"), strictly above the global minimum, not on the
    1-px border; returned as (row[, ...]) coordinates sorted by descending intensity (stable)."""
    image_max = ndimage.maximum_filter(image, size=3, mode="nearest")
    mask = image == image_max
    if np.all(mask):
        mask[:] = False
    mask &= image > image.min()
    for axis in range(image.ndim):
        sl = [slice(None)] * image.ndim
        sl[axis] = slice(0, 1)
        mask[tuple(sl)] = False
        sl[axis] = slice(-1, None)
        mask[tuple(sl)] = False
    coords = np.nonzero(mask)
    order = np.argsort(-image[coords], kind="stable")
    return np.transpose(coords)[order]


def find_seeds(embeddings_centered: np.ndarray, sigma: float = 2.0) -> np.ndarray:
    """`detect.py:129-132`: (n_seeds, D) int64 in (x, y[, z]) order."""
    magnitude = np.linalg.norm(embeddings_centered[:-1], axis=0)
    smooth = ndimage.gaussian_filter(magnitude, sigma=sigma)
    return np.flip(peak_local_max(-smooth), 1)


def use_seeds_detections(embeddings_centered: np.ndarray, threshold: float, bandwidth: float, num_bandwidths: int,
                         reduction_probability: float) -> np.ndarray:
    """The `use_seeds=True` bandwidth loop of `detect.py:121-144,160`, side effects included.

    `embeddings_centered_mean` starts as a VIEW of `embeddings_centered` (`:121-123`), so the first
    `mean_shift_segmentation` call adds the coordinate grids to `embeddings_centered` itself
    (`utils/mean_shift.py:15-32`); from the second bandwidth on the seeds are found on those shifted channels
    (`:129-132`) and the clustering runs on a COPY of them (`:142-144`) to which the coordinates are added once
    more.  Returns the `(num_bandwidths, *S)` int32 detections.  Mutates `embeddings_centered` like the reference.
    """
    from oracle import mean_shift as oms

    nd = embeddings_centered.shape[0] - 1
    mean = embeddings_centered[np.newaxis, :nd]  # a view
    std = embeddings_centered[-1]
    out = []
    for k in range(num_bandwidths):
        seeds = find_seeds(embeddings_centered)
        out.append(oms.mean_shift_segmentation(mean, std, bandwidth / (2**k), 0, reduction_probability, threshold, seeds))
        mean = embeddings_centered[np.newaxis, :nd, ...].copy()
    return np.stack(out)
