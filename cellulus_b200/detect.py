"""Embedding -> instance detection on the B200 kernels (`cellulus/detect.py:82-161`).

`detect_embeddings` is the per-sample body of the reference's `detect()` for
`clustering="meanshift"`: foreground threshold (Otsu on the std channel unless
configured), flat-kernel mean-shift per bandwidth, nearest-mode labels.  The
zarr I/O around it stays with the caller.
"""

from __future__ import annotations

import numpy as np
import torch

from cellulus_b200 import kernels as K
from cellulus_b200.utils import mean_shift as MS


def otsu_from_histogram(counts: np.ndarray, edges: np.ndarray):
    """The O(nbins) tail of scikit-image's `threshold_otsu` (host, 256 numbers):
    float32 counts, bin centres, argmax of w1*w2*(mu1-mu2)^2, returns a bin centre."""
    bin_centers = (edges[:-1] + edges[1:]) / 2.0
    counts = counts.astype("float32", copy=False)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(counts * bin_centers) / weight1
        mean2 = (np.cumsum((counts * bin_centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return bin_centers[np.argmax(variance12)]


def threshold_otsu(std: torch.Tensor, nbins: int = 256):
    """`threshold_otsu(embeddings_std)` of `detect.py:88-89` for a CUDA tensor:
    min/max and the numpy-exact 256-bin histogram run on the device (the only
    O(N) part); the 256-number tail runs on the host.  Returns a python float."""
    mm = K.minmax(std).cpu().numpy()
    lo, hi = float(mm[0]), float(mm[1])
    if lo == hi:  # constant image: skimage returns that value
        return lo
    edges = np.linspace(lo, hi, nbins + 1, endpoint=True, dtype=np.float64)  # np.histogram's bin edges
    counts = K.histogram(std, torch.from_numpy(edges).to(std.device)).cpu().numpy()
    return float(otsu_from_histogram(counts, edges))


def detect_embeddings(embeddings, bandwidth, threshold=None, num_bandwidths=1, reduction_probability=0.1,
                      seeds=None, rng="numpy", method="auto", label_dtype=torch.uint16, return_info=False):
    """Per-sample body of `detect.py:82-161` (`clustering="meanshift"`, `use_seeds=False`).

    embeddings : (D+1, *S) CUDA tensor, fp32 or fp64 (channel D = std), or a numpy array (uploaded)
    returns    : (num_bandwidths, *S) label tensor (uint16 like `detect.py:30`), the threshold used,
                 the (*S) uint8 foreground mask [and per-bandwidth info dicts]
    """
    if isinstance(embeddings, np.ndarray):
        embeddings = torch.from_numpy(np.ascontiguousarray(embeddings)).cuda()
    D = embeddings.shape[0] - 1
    if threshold is None:
        threshold = threshold_otsu(embeddings[D])
    out, infos, mask = [], [], None
    for k in range(num_bandwidths):
        labels, info = MS.segment_embeddings_device(
            embeddings, bandwidth / (2**k), threshold, reduction_probability, seeds=seeds, rng=rng, method=method,
            label_dtype=label_dtype, want_mask=(k == 0))
        if k == 0:
            mask = info.pop("mask")
        out.append(labels)
        infos.append(info)
    result = (torch.stack(out, 0), threshold, mask)
    return result + (infos,) if return_info else result
