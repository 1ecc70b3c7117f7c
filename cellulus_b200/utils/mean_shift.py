"""Mean-shift segmentation of per-pixel embeddings on the B200 kernels.

`mean_shift_segmentation` keeps the reference's signature and side effects
(`cellulus/utils/mean_shift.py:6-45`): numpy in / numpy out, coordinates added
to `embedding_mean` IN PLACE, fit subset drawn from the global numpy RNG.
`segment_embeddings_device` is the device-resident core used by it, by
`cellulus_b200.detect` and by the benchmark.
"""

from __future__ import annotations

import numpy as np
import torch

from cellulus_b200 import kernels as K

# brute force is exact and simple; beyond this many (seed x point) pair tests per iteration the
# grid-hash kernel takes over ("auto")
_BRUTE_PAIR_LIMIT = 64_000_000


def _seeds_to_soa(seeds, D, device):
    s = torch.as_tensor(np.asarray(seeds), dtype=torch.float64)
    if s.ndim != 2 or s.shape[1] != D:
        raise ValueError(f"seeds must be (n_seeds, {D})")
    n = s.shape[0]
    cap = max(2, (n + 1) & ~1)
    soa = torch.zeros((D, cap), dtype=torch.float64, device=device)
    soa[:, :n] = s.t().to(device)
    return soa, n


def cluster_points_device(points, n, fit_points, n_fit, bandwidth, seeds=None, method="auto", max_iter=300,
                          bin_seeding=False, distinct=False):
    """sklearn `MeanShift(bandwidth, seeds).fit(fit_points)` centre finding on the device.

    points / fit_points: SoA (D, cap) float64.  Returns `(centres SoA (D, >=K), K, info)`.
    `distinct`: climb distinct trajectories only (`cb200_ms_grid_modes_distinct`, grid method): same centres; the
    per-seed `counts` / `iters` in `info` then mark the merged copies (count 0, negative iterations).
    """
    D = points.shape[0]
    dev = points.device
    if n_fit == 0:
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by MeanShift.")
    if seeds is None and bin_seeding:
        # sklearn MeanShift(bin_seeding=True): seeds = get_bin_seeds(X, bandwidth) (sklearn:493-494)
        seeds_soa, n_seeds = K.bin_seeds(fit_points, n_fit, bandwidth)
    elif seeds is None:
        seeds_soa, n_seeds = fit_points.clone(), n_fit  # every fit point is a seed (sklearn:491-496)
    else:
        seeds_soa, n_seeds = _seeds_to_soa(seeds, D, dev)
    lo, hi = K.bounding_box(fit_points, n_fit)
    grid = K.plan_grid(lo, hi, bandwidth)
    if method == "auto":
        method = "brute" if n_seeds * n_fit <= _BRUTE_PAIR_LIMIT else "grid"
    if method == "grid":
        sorted_pts, cell_start, _ = K.grid_build(fit_points, n_fit, grid)
        # the merge pays from a few ten thousand seeds on (two more launches, a hash pass and a compaction)
        climb = K.ms_grid_modes_distinct if (distinct and n_seeds >= 20_000) else K.ms_grid_modes
        counts, iters = climb(sorted_pts, n_fit, grid, cell_start, seeds_soa, n_seeds, bandwidth, max_iter)
    elif method == "brute":
        counts, iters = K.ms_brute_modes(fit_points, n_fit, seeds_soa, n_seeds, bandwidth, max_iter)
    else:
        raise ValueError("method must be 'auto', 'grid' or 'brute'")
    centres, k = K.nms_centres(seeds_soa, counts, n_seeds, bandwidth, grid)
    if k == 0:
        raise ValueError(
            "No point was within bandwidth=%f of any seed. Try a different seeding strategy "
            "                             or increase the bandwidth." % bandwidth)
    info = {"method": method, "n_seeds": n_seeds, "n_fit": n_fit, "modes": seeds_soa, "counts": counts,
            "iters": iters, "grid_cells": int(grid.n_cells), "grid": grid}
    return centres, k, info


def segment_embeddings_device(emb, bandwidth, threshold, reduction_probability=1.0, seeds=None, rng="numpy",
                              fit_flags=None, method="auto", label_dtype=torch.int32, want_mask=False,
                              philox_seed=0, assign="grid", one_call=False, bin_seeding=False, distinct=False):
    """threshold -> foreground points -> fit subset -> modes -> centres -> labels, all on the device.

    emb: (D+1, *S) CUDA tensor (fp32/fp64), channel D = std.  Returns `(labels (*S), info)`;
    labels are 0 for background and 1..K otherwise (`utils/mean_shift.py:57,101-104`).
    `rng`: "numpy" draws the fit subset exactly like the reference (`np.random.rand(N) < p`,
    :68-70, global RNG) and uploads the flags; "philox" draws it on the device.
    `bin_seeding`: seed with scikit-learn's grid-binned seeds (`MeanShift(bin_seeding=True)`) instead of every fit
    point -- BASELINE configs[3]'s second seeding mode; the reference itself never enables it.
    `one_call`: run the identical sequence inside the library (`cb200_detect_volume`: one C-ABI call, no
    interpreter between the kernels); needs the device RNG (or no subsampling) and the grid kernels, and
    reports counts only (no per-seed modes / iterations in `info`).
    `distinct`: see `cluster_points_device` (same labels, less climbing; per-seed info marks the merged seeds).
    """
    if one_call and seeds is None and not bin_seeding and fit_flags is None and method in ("auto", "grid") and assign == "grid" and (
            rng == "philox" or reduction_probability >= 1.0):
        labels, mask, _, info = K.detect_volume(emb, bandwidth, threshold, reduction_probability, philox_seed,
                                               label_dtype=label_dtype, want_mask=want_mask)
        if want_mask:
            info["mask"] = mask
        return labels, info
    D = emb.shape[0] - 1
    spatial = tuple(emb.shape[1:])
    dev = emb.device
    pts, pix, n, mask = K.fg_compact(emb, threshold, mask_dtype=torch.uint8 if want_mask else None)
    labels = torch.zeros(spatial, dtype=label_dtype, device=dev)
    info = {"n_fg": n}
    if want_mask:
        info["mask"] = mask
    if n == 0:  # utils/mean_shift.py:83-84 -> all background
        info.update({"k": 0})
        return labels, info
    if reduction_probability < 1.0:
        if fit_flags is None:
            if rng == "numpy":
                fit_flags = torch.from_numpy((np.random.rand(n) < reduction_probability).astype(np.uint8)).to(dev)
            elif rng == "philox":
                fit_flags = K.bernoulli_flags(n, reduction_probability, philox_seed, dev)
            else:
                raise ValueError("rng must be 'numpy' or 'philox'")
        fit_pts, n_fit = K.select_points(pts, n, fit_flags)
    else:
        fit_pts, n_fit = pts, n
    centres, k, cinfo = cluster_points_device(pts, n, fit_pts, n_fit, bandwidth, seeds=seeds, method=method,
                                              bin_seeding=bin_seeding, distinct=distinct)
    # predict on ALL foreground (:74), scatter, +1; pruned nearest-centre search over the same cell grid
    K.assign_labels(pts, n, centres, k, pix, labels, grid=cinfo["grid"] if assign == "grid" else None)
    info.update(cinfo)
    info.update({"k": k, "centres": centres[:, :k]})
    return labels, info


def mean_shift_segmentation(
    embedding_mean,
    embedding_std,
    bandwidth,
    min_size,
    reduction_probability,
    threshold,
    seeds,
    device="cuda",
    method="auto",
):
    """Drop-in for `cellulus/utils/mean_shift.py:6-45`.

    embedding_mean (1, D, *S) and embedding_std (*S) numpy arrays -> int32 (*S)
    labels, 0 = background.  Like the reference it (a) adds the pixel
    coordinates to `embedding_mean` in place (callers re-copy, `detect.py:142-160`),
    (b) draws the fit subset with `np.random.rand`, (c) ignores `min_size`.
    """
    del min_size  # unused in the reference too (:10)
    emb_np = np.asarray(embedding_mean)
    if emb_np.ndim not in (4, 5) or emb_np.shape[0] != 1:
        raise ValueError("embedding_mean must be (1, D, *S)")
    D = emb_np.shape[1]
    dtype = torch.float64 if emb_np.dtype == np.float64 or np.asarray(embedding_std).dtype == np.float64 \
        else torch.float32
    dev = torch.device(device)
    emb = torch.empty((D + 1, *emb_np.shape[2:]), dtype=dtype, device=dev)
    emb[:D] = torch.from_numpy(np.ascontiguousarray(emb_np[0])).to(dev)
    emb[D] = torch.from_numpy(np.ascontiguousarray(np.asarray(embedding_std))).to(dev)
    # (a) the in-place coordinate add is part of the reference's observable behaviour (:15-32)
    spatial = emb_np.shape[2:]
    for ch in range(D):
        axis = D - 1 - ch
        shape = [1] * D
        shape[axis] = spatial[axis]
        emb_np[0, ch] += np.arange(spatial[axis]).reshape(shape)
    with torch.cuda.device(dev):
        labels, _ = segment_embeddings_device(emb, float(bandwidth), float(threshold), float(reduction_probability),
                                              seeds=seeds, rng="numpy", method=method, label_dtype=torch.int32,
                                              distinct=True)
    return labels.cpu().numpy()
