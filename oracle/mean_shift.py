"""Oracle: embedding -> instance detection by flat-kernel mean-shift.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Two layers:

* `mean_shift_segmentation` / `AnchorMeanshift`: a port of the reference's
  wrapper `cellulus/utils/mean_shift.py:6-121`, which delegates the arithmetic
  to `sklearn.cluster.MeanShift` exactly like the reference does (scikit-learn
  is an unpinned dependency of the reference, `pyproject.toml:29`; version
  1.9.0 is installed in this image, here and on the GPU box).  This is the
  function timed as the CPU baseline.
* `mean_shift_modes` / `nms_centres` / `predict_labels`: a numpy restatement
  of the published algorithm inside scikit-learn 1.9.0
  (`sklearn/cluster/_mean_shift.py:108-128` hill climb, `:511-547` dedupe +
  greedy NMS, `:563-579` predict), exposing the per-seed `(mode, count,
  iterations)` triples that sklearn does not return, so that the CUDA kernels
  can be checked stage by stage.  `tests/test_oracle.py` checks this
  restatement against sklearn itself and against the committed goldens.
"""

from __future__ import annotations

import numpy as np
import torch


# --------------------------------------------------------------------------
# layer 1: the reference wrapper (port), arithmetic delegated to scikit-learn
# --------------------------------------------------------------------------
def add_coordinates_(embedding_mean: torch.Tensor) -> None:
    """`utils/mean_shift.py:15-32`: in-place add of the pixel-coordinate grids;
    channel 0 += x (last axis), 1 += y, [2 += z (first spatial axis)]."""
    if embedding_mean.ndim == 4:
        embedding_mean[:, 1] += torch.arange(embedding_mean.shape[2])[None, :, None]
        embedding_mean[:, 0] += torch.arange(embedding_mean.shape[3])[None, None, :]
    elif embedding_mean.ndim == 5:
        embedding_mean[:, 2] += torch.arange(embedding_mean.shape[2])[None, :, None, None]
        embedding_mean[:, 1] += torch.arange(embedding_mean.shape[3])[None, None, :, None]
        embedding_mean[:, 0] += torch.arange(embedding_mean.shape[4])[None, None, None, :]


def mean_shift_segmentation(
    embedding_mean, embedding_std, bandwidth, min_size, reduction_probability, threshold, seeds
):
    """`utils/mean_shift.py:6-45`.  Mutates `embedding_mean` (a numpy array)
    in place, draws the fit subset from the global numpy RNG, returns int32
    labels with 0 = background."""
    embedding_mean = torch.from_numpy(embedding_mean)
    add_coordinates_(embedding_mean)
    mask = embedding_std < threshold
    mask = mask[None]
    ams = AnchorMeanshift(bandwidth, reduction_probability, cluster_all=False, seeds=seeds)
    return (ams(embedding_mean, mask=mask) + 1)[0]


class AnchorMeanshift:
    """`utils/mean_shift.py:60-121`."""

    def __init__(self, bandwidth, reduction_probability, cluster_all, seeds):
        from sklearn.cluster import MeanShift

        self.mean_shift = MeanShift(bandwidth=bandwidth, cluster_all=cluster_all, seeds=seeds)
        self.reduction_probability = reduction_probability

    def compute_mean_shift(self, X):  # :67-76
        if self.reduction_probability < 1.0:
            X_reduced = X[np.random.rand(len(X)) < self.reduction_probability]
            self.mean_shift.fit(X_reduced)
        else:
            self.mean_shift.fit(X)
        return self.mean_shift.predict(X)

    def compute_masked_ms(self, embedding, mask):  # :78-110 (mask is never None on the path)
        c = embedding.shape[0]
        if mask.sum() == 0:
            return -1 * np.ones(mask.shape, dtype=np.int32)
        if embedding.ndim == 3:
            X = embedding.permute(1, 2, 0)[mask].view(-1, c)
        else:
            X = embedding.permute(1, 2, 3, 0)[mask].view(-1, c)
        X = X.contiguous().numpy()
        labels = self.compute_mean_shift(X)
        out = -1 * np.ones(mask.shape, dtype=np.int32)
        out[mask] = labels
        return out

    def __call__(self, embedding, mask):  # :112-121
        return np.stack(
            [self.compute_masked_ms(embedding[j], mask[j]) for j in range(len(embedding))]
        )


# --------------------------------------------------------------------------
# layer 2: the published algorithm restated (numpy, float64, brute force)
# --------------------------------------------------------------------------
def points_from_embedding(embedding_mean: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """`utils/mean_shift.py:16-32,85,94`: foreground points (N, D) float64 in
    raster order, column k = channel k + coordinate k (x first)."""
    emb = np.array(embedding_mean, dtype=np.float64)
    D = emb.shape[0]
    spatial = emb.shape[1:]
    for ch in range(D):
        axis = D - 1 - ch
        shape = [1] * D
        shape[axis] = spatial[axis]
        emb[ch] += np.arange(spatial[axis]).reshape(shape)
    return np.ascontiguousarray(np.moveaxis(emb, 0, -1)[mask])


def _sqdist(means: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Reduced distance exactly as the KD-tree evaluates it: sequential
    `d += t*t` over the dims in float64 (no fused multiply-add)."""
    d = np.zeros((means.shape[0], X.shape[0]), dtype=np.float64)
    for k in range(X.shape[1]):
        t = means[:, k : k + 1] - X[None, :, k]
        d += t * t
    return d


def mean_shift_modes(X, seeds, bandwidth, max_iter=300, chunk=256):
    """`_mean_shift.py:108-128` for every seed.

    loop: neighbours = {x : sum (m-x)^2 <= bw^2} (inclusive); empty -> stop;
    m <- mean(neighbours); stop if ||m - m_old|| <= 1e-3*bw or
    completed == max_iter; completed += 1.
    Returns `(modes (S,D) f64, counts (S,) int, iterations (S,) int)`; the
    count is that of the LAST neighbourhood evaluated (around the old mean).
    """
    X = np.asarray(X, dtype=np.float64)
    means = np.array(seeds, dtype=np.float64, copy=True)
    S = len(means)
    counts = np.zeros(S, dtype=np.int64)
    iters = np.zeros(S, dtype=np.int64)
    active = np.ones(S, dtype=bool)
    r2 = bandwidth * bandwidth
    stop = 1e-3 * bandwidth
    while active.any():
        idx = np.nonzero(active)[0]
        for s0 in range(0, len(idx), chunk):
            ii = idx[s0 : s0 + chunk]
            within = _sqdist(means[ii], X) <= r2
            n = within.sum(axis=1)
            for j, i in enumerate(ii):
                counts[i] = n[j]
                if n[j] == 0:
                    active[i] = False
                    continue
                old = means[i].copy()
                means[i] = np.mean(X[within[j]], axis=0)
                if np.linalg.norm(means[i] - old) <= stop or iters[i] == max_iter:
                    active[i] = False
                else:
                    iters[i] += 1
    return means, counts, iters


def nms_centres(modes, counts, bandwidth):
    """`_mean_shift.py:511-547`: drop empty seeds, dedupe exact tuples, sort by
    (count, coords) descending, greedy suppression of every centre within
    `bandwidth` (inclusive) of a surviving earlier one."""
    d = {}
    for m, c in zip(modes, counts):
        if c:
            d[tuple(m)] = int(c)
    if not d:
        raise ValueError("No point was within bandwidth of any seed.")
    ordered = sorted(d.items(), key=lambda t: (t[1], t[0]), reverse=True)
    centres = np.array([t[0] for t in ordered])
    unique = np.ones(len(centres), dtype=bool)
    r2 = bandwidth * bandwidth
    for i in range(len(centres)):
        if unique[i]:
            near = _sqdist(centres[i : i + 1], centres)[0] <= r2
            unique[near] = False
            unique[i] = True
    return centres[unique]


def predict_labels(X, centres, chunk=4096):
    """`_mean_shift.py:563-579`: nearest centre (Euclidean), ties -> lowest index."""
    X = np.asarray(X, dtype=np.float64)
    out = np.empty(len(X), dtype=np.int64)
    for s0 in range(0, len(X), chunk):
        out[s0 : s0 + chunk] = np.argmin(_sqdist(X[s0 : s0 + chunk], centres), axis=1)
    return out


def get_bin_seeds(X, bin_size):
    """`_mean_shift.py:254-297` (min_bin_freq=1): unique `round(x / bin)` bins in
    first-seen order, cast to float32, times `bin_size`."""
    if bin_size == 0:
        return X
    binned = np.round(np.asarray(X) / bin_size)
    _, first = np.unique(binned, axis=0, return_index=True)
    bins = binned[np.sort(first)].astype(np.float32)
    if len(bins) == len(X):
        return X
    return bins * bin_size


def segment_points(X, fit_mask, bandwidth, seeds=None):
    """fit on `X[fit_mask]` (`utils/mean_shift.py:67-72`), predict on all of
    `X` (`:74`).  Returns `(labels (N,), centres (K,D))`."""
    Xr = X if fit_mask is None else X[fit_mask]
    s = Xr if seeds is None else np.asarray(seeds, dtype=np.float64)
    modes, counts, _ = mean_shift_modes(Xr, s, bandwidth)
    centres = nms_centres(modes, counts, bandwidth)
    return predict_labels(X, centres), centres


def sklearn_cluster_seconds(X, bandwidth, seeds=None, bin_seeding=False):
    """Wall time of the reference's engine on a point set: `MeanShift(bandwidth, seeds=, bin_seeding=,
    cluster_all=False).fit(X)` followed by `.predict(X)` (what `AnchorMeanshift.compute_mean_shift` runs,
    `cellulus/utils/mean_shift.py:60-76`; `n_jobs=None`: the hill climb is single-core).  For the CPU legs of the
    benchmark sweep.  Returns `(fit_seconds, predict_seconds, n_seeds_climbed, n_centres)`."""
    import time

    from sklearn.cluster import MeanShift

    ms = MeanShift(bandwidth=bandwidth, seeds=seeds, bin_seeding=bin_seeding, cluster_all=False)
    t0 = time.perf_counter()
    ms.fit(X)
    t1 = time.perf_counter()
    ms.predict(X)
    t2 = time.perf_counter()
    n_seeds = len(seeds) if seeds is not None else (len(get_bin_seeds(X, bandwidth)) if bin_seeding else len(X))
    return t1 - t0, t2 - t1, n_seeds, len(ms.cluster_centers_)
