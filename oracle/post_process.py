"""Oracle: `segment()` post-processing (numpy + scipy).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows `cellulus/segment.py:41-101`.  The distance
transform and the hole filling are scipy's own (`scipy.ndimage.distance_transform_edt`,
`binary_fill_holes`, the functions the reference calls; scipy is installed, so these two are the real
thing, version recorded by `tests/golden/make_golden.py`).  `threshold_otsu` is scikit-image's
(not installed): restated in `oracle/otsu.py`, **parity unpinned** for that piece.
"""

from __future__ import annotations

import numpy as np
from scipy.ndimage import binary_fill_holes
from scipy.ndimage import distance_transform_edt as dtedt

from . import otsu as _otsu


def grow_shrink(segmentation: np.ndarray, grow_distance, shrink_distance) -> np.ndarray:
    """`segment.py:46-50`, in place like the reference; returns the same array."""
    distance_foreground = dtedt(segmentation == 0)
    expanded_mask = distance_foreground < grow_distance
    distance_background = dtedt(expanded_mask)
    segmentation[distance_background < shrink_distance] = 0
    return segmentation


def edt_within(mask: np.ndarray, radius) -> np.ndarray:
    return dtedt(mask) < radius


def nucleus(segmentation: np.ndarray, raw_image: np.ndarray) -> np.ndarray:
    """`segment.py:52-101`: per instance, Otsu on the raw intensities under the instance, holes filled inside
    the instance's bounding box; instances written in ascending id order (later ids overwrite)."""
    out = np.zeros_like(segmentation)
    ids = np.unique(segmentation)
    for id_ in ids[ids != 0]:
        m = segmentation == id_
        idx = np.where(m)
        box = tuple(slice(int(i.min()), int(i.max()) + 1) for i in idx)
        mask = m & (raw_image > _otsu.threshold_otsu(raw_image[m]))
        mask[box] = binary_fill_holes(mask[box])
        out[mask] = id_
    return out
