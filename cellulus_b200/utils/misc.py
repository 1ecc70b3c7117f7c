"""`size_filter` of `cellulus/utils/misc.py:11-25` on the B200 connected-component kernels."""

from __future__ import annotations

import numpy as np
import torch

from cellulus_b200 import kernels as K


def size_filter_device(segmentation: torch.Tensor, min_size: int) -> torch.Tensor:
    """Device form: `segmentation` int32 CUDA tensor, modified in place (small
    components zeroed) exactly like the reference; returns the relabelled image."""
    if min_size == 0:
        return segmentation
    labels, _ = K.size_filter_(segmentation, int(min_size))
    return labels


def size_filter(segmentation, min_size, filter_non_connected=True, device="cuda"):
    """Drop-in for `utils/misc.py:11-25` (numpy in, numpy out).

    `min_size == 0` returns the input untouched; otherwise components (full
    connectivity, equal-valued regions) smaller than `min_size` are zeroed IN
    PLACE in `segmentation` and the relabelled image is returned.
    """
    if min_size == 0:
        return segmentation
    if not filter_non_connected:
        raise NotImplementedError(
            "filter_non_connected=False is never used by the reference (segment.py:104-108) and is not built")
    seg_np = np.asarray(segmentation)
    seg = torch.from_numpy(np.ascontiguousarray(seg_np).astype(np.int32)).to(device)
    labels = size_filter_device(seg, int(min_size))
    seg_np[...] = seg.cpu().numpy().astype(seg_np.dtype)  # the reference mutates its argument (:23)
    return labels.cpu().numpy().astype(np.int64)  # skimage.measure.label returns an integer label image
