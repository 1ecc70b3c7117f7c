"""Oracle: foreground threshold of the detect preamble (numpy).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows
`cellulus/detect.py:88-94`.  `threshold_otsu` is scikit-image's (an unpinned
dependency, `pyproject.toml:26`; NOT installed here and its source is not
under /root/reference): it is restated from the published algorithm --
**parity unpinned** -- on top of `np.histogram`, which *is* the function
scikit-image calls for float images and is present.
"""

from __future__ import annotations

import numpy as np


def threshold_otsu(image: np.ndarray, nbins: int = 256):
    """scikit-image `filters.threshold_otsu` for a float image:

    * constant image -> that value;
    * `counts, edges = np.histogram(image.ravel(), nbins)` over [min, max],
      bin centres = midpoints, counts cast to float32
      (`_validate_image_histogram`);
    * maximise `w1[:-1] * w2[1:] * (mu1[:-1] - mu2[1:])**2`; return the bin
      *centre* at the argmax.
    """
    flat = image.reshape(-1)
    first = flat[0]
    if np.all(image == first):
        return first
    if np.issubdtype(flat.dtype, np.integer):
        # `exposure.histogram` takes its bincount path for integer images: one bin per value in
        # [min, max], the bin centres are those integers
        lo, hi = int(flat.min()), int(flat.max())
        counts = np.bincount((flat.astype(np.int64) - lo).ravel(), minlength=hi - lo + 1)
        return otsu_from_centers(counts, np.arange(lo, hi + 1))
    counts, edges = np.histogram(flat, nbins)
    return otsu_from_histogram(counts, edges)


def otsu_from_histogram(counts: np.ndarray, edges: np.ndarray):
    """The O(nbins) tail of `threshold_otsu`, split out because the CUDA path
    computes the histogram on the device and finishes with these few lines."""
    return otsu_from_centers(counts, (edges[:-1] + edges[1:]) / 2.0)


def otsu_from_centers(counts: np.ndarray, bin_centers: np.ndarray):
    counts = counts.astype("float32", copy=False)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(counts * bin_centers) / weight1
        mean2 = (np.cumsum((counts * bin_centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    idx = np.argmax(variance12)
    return bin_centers[idx]


def foreground_mask(embeddings_std: np.ndarray, threshold=None):
    """`detect.py:88-94`: Otsu unless configured; `mask = std < threshold`."""
    if threshold is None:
        threshold = threshold_otsu(embeddings_std)
    return embeddings_std < threshold, threshold


def centre_embeddings(embeddings: np.ndarray, binary_mask: np.ndarray) -> np.ndarray:
    """`detect.py:97-119`: subtract, per offset channel, the mean over the
    NON-ZERO entries of `mask * channel` (the std channel is left alone)."""
    centred = embeddings.copy()
    D = embeddings.shape[0] - 1
    for ch in range(D):
        masked = binary_mask * embeddings[ch]
        centred[ch] -= masked[masked != 0].mean()
    return centred
