"""Embedding -> instance detection on the B200 kernels (`cellulus/detect.py:82-161`).

`detect_embeddings` is the per-sample body of the reference's `detect()` for
`clustering="meanshift"`: foreground threshold (Otsu on the std channel unless
configured), flat-kernel mean-shift per bandwidth, nearest-mode labels.  The
zarr I/O around it stays with the caller.
"""

from __future__ import annotations

import numpy as np
import torch

import os

from cellulus_b200 import kernels as K
from cellulus_b200 import sharding, zarr_lite
from cellulus_b200.utils import mean_shift as MS
from cellulus_b200.utils.device import resolve_device


def otsu_from_histogram(counts: np.ndarray, edges: np.ndarray):
    """The O(nbins) tail of scikit-image's `threshold_otsu` (host, 256 numbers):
    float32 counts, bin centres, argmax of w1*w2*(mu1-mu2)^2, returns a bin centre."""
    return otsu_from_centers(counts, (edges[:-1] + edges[1:]) / 2.0)


def otsu_from_centers(counts: np.ndarray, bin_centers: np.ndarray):
    """Same tail for any histogram given as (counts, bin centres) -- integer images have one bin per value."""
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


def _warn_if_labels_wrap(k: int, sample, bandwidth) -> None:
    """the detection dataset is uint16 (`detect.py:30,161`); int32 labels above 65 535 are
    silently narrowed there.  The dtype is kept (zarr layout contract) but the wrap is reported."""
    if k > 65535:
        import warnings

        warnings.warn(f"sample {sample}, bandwidth {bandwidth}: {k} instances exceed the uint16 range of the "
                      "`detection` dataset (cellulus/detect.py:30); labels above 65535 wrap around, as in the reference")


def detect_embeddings(embeddings, bandwidth, threshold=None, num_bandwidths=1, reduction_probability=0.1,
                      seeds=None, rng="numpy", method="auto", label_dtype=torch.uint16, return_info=False,
                      one_call=None, bin_seeding=False):
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
    if one_call is None:  # per-seed details are only available from the step-by-step sequence
        one_call = not return_info
    out, infos, mask = [], [], None
    for k in range(num_bandwidths):
        labels, info = MS.segment_embeddings_device(
            embeddings, bandwidth / (2**k), threshold, reduction_probability, seeds=seeds, rng=rng, method=method,
            label_dtype=label_dtype, want_mask=(k == 0), one_call=one_call, bin_seeding=bin_seeding,
            distinct=not return_info)
        if k == 0:
            mask = info.pop("mask")
        if label_dtype == torch.uint16:
            _warn_if_labels_wrap(int(info.get("k", 0)), "?", bandwidth / (2**k))
        out.append(labels)
        infos.append(info)
    result = (torch.stack(out, 0), threshold, mask)
    return result + (infos,) if return_info else result


def add_coordinates(embeddings: torch.Tensor) -> torch.Tensor:
    """Copy of a (D+1, *S) embedding tensor with the pixel coordinates added to its first D channels: channel 0 += x
    (last axis), channel 1 += y, channel 2 += z -- what `mean_shift_segmentation` does to its argument in place
    (`utils/mean_shift.py:15-32`)."""
    D = embeddings.shape[0] - 1
    out = embeddings.clone()
    for c in range(D):
        axis = D - 1 - c  # spatial axis of column c
        shape = [1] * D
        shape[axis] = embeddings.shape[1 + axis]
        out[c] += torch.arange(embeddings.shape[1 + axis], device=embeddings.device, dtype=embeddings.dtype).view(shape)
    return out


def detect_with_seeds(centred, bandwidth, threshold, num_bandwidths, reduction_probability,
                      label_dtype=torch.uint16):
    """The `use_seeds=True` bandwidth loop of `detect.py:121-144`, WITH its side effect: the first
    `mean_shift_segmentation` call adds the coordinate grids to `embeddings_centered` through a view, so from the
    second bandwidth on the reference finds its seeds on the shifted channels and clusters a copy of them to which
    the coordinates are added once more.  Reproduced as is (drop-in: same detections for every bandwidth index --
    including the reference's failure mode: with ordinary bandwidths the displaced seeds of index >= 1 reach no
    point and scikit-learn's "No point was within bandwidth" ValueError comes up, here as there).
    Returns `((num_bandwidths, *S) labels, (*S) uint8 mask)`."""
    labels, mask, shifted = [], None, None
    for k in range(num_bandwidths):
        source = centred if k == 0 else shifted
        lab, _, m = detect_embeddings(source, bandwidth / (2**k), threshold, 1, reduction_probability,
                                      seeds=K.find_seeds(source), rng="numpy", label_dtype=label_dtype)
        labels.append(lab[0])
        if k == 0:
            mask = m
            if num_bandwidths > 1:
                shifted = add_coordinates(centred)
    return torch.stack(labels, 0), mask


def _detect_sample_sharded(cfg, ds, sample, nd, device, rank, world, ds_detection, ds_binary, ds_centred):
    """One sample, all ranks: rank 0 does the O(N) preamble on the whole sample (threshold, mask, centring --
    `detect.py:88-119`), every rank compacts its slab of the slowest axis and the mean-shift runs seed-sharded
    (`sharding.sharded_mean_shift`: two all-gathers).  Results are identical to the single-GPU path."""
    import torch.distributed as dist

    spatial = tuple(ds.shape[2:])
    thr_t = torch.zeros(1, dtype=torch.float64, device=device)
    if rank == 0:
        emb = torch.from_numpy(np.ascontiguousarray(ds[sample])).to(device)
        threshold = cfg.threshold if cfg.threshold is not None else threshold_otsu(emb[nd])
        print(f"For sample {sample}, binary threshold {threshold} was used.")
        _, centred = K.centre_embeddings(emb, threshold)
        ds_binary[sample, 0, ...] = (emb[nd] < threshold).cpu().numpy().astype(np.uint16)
        ds_centred[sample] = centred.cpu().numpy()
        thr_t[0] = threshold
        del emb, centred
    dist.broadcast(thr_t, 0)
    threshold = float(thr_t.item())
    rows = sharding.shard_items(spatial[0], rank, world)  # slab of the slowest spatial axis: rank order = raster order
    slab_np = np.ascontiguousarray(ds[(sample, slice(None), slice(rows.start, rows.stop))])
    slab = torch.from_numpy(slab_np).to(device)
    pts, pix, n_local, _ = K.fg_compact(slab, threshold) if len(rows) else (torch.zeros((nd, 2), dtype=torch.float64,
                                                                                          device=device), None, 0, None)
    if n_local:
        pts[nd - 1, :n_local] += float(rows.start)  # column nd-1 is the slowest axis (z in 3-D, y in 2-D)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = n_local
    dist.all_reduce(counts)
    n_total, start = int(counts.sum()), int(counts[:rank].sum())
    out = []
    for k in range(cfg.num_bandwidths):
        bandwidth = cfg.bandwidth / (2**k)
        labels_slab = torch.zeros((len(rows), *spatial[1:]), dtype=torch.int32, device=device)
        if n_total:
            flags_local = None
            if cfg.reduction_probability < 1.0:  # `np.random.rand(N) < p` over ALL points (utils/mean_shift.py:68-70)
                flags = torch.zeros(n_total, dtype=torch.uint8, device=device)
                if rank == 0:
                    flags.copy_(torch.from_numpy((np.random.rand(n_total) < cfg.reduction_probability).astype(np.uint8)))
                dist.broadcast(flags, 0)
                flags_local = flags[start:start + n_local]
            labels, centres = sharding.sharded_mean_shift(pts, n_local, bandwidth, sharding.cuda_ops("auto"), flags_local)
            if n_local:
                labels_slab.view(-1)[pix[:n_local].long()] = labels.to(torch.int32)
            if rank == 0:
                _warn_if_labels_wrap(int(centres.shape[1]), sample, bandwidth)
        out.append(labels_slab.to(torch.uint16))
    mine = torch.stack(out, 0).cpu().numpy()  # (num_bandwidths, rows, ...)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank == 0:
        ds_detection[sample] = np.concatenate(gathered, axis=1)


def detect(inference_config) -> None:
    """`detect(inference_config)` (`cellulus/detect.py:14-192`) for `clustering="meanshift"` without seeds:
    reads the `embeddings` dataset, writes `detection` (uint16, `(s, num_bandwidths, *spatial)`),
    `binary-segmentation` and `centered-embeddings` with the reference's attributes.  Samples are dealt
    round-robin to the ranks under torchrun (each sample is one independent unit, `detect.py:82`)."""
    from cellulus_b200.datasets.meta_data import DatasetMetaData

    meta = DatasetMetaData.from_dataset_config(inference_config.dataset_config)
    nd = meta.num_spatial_dims
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = resolve_device(inference_config.device, "detect")
    cfg = inference_config.detection_dataset_config
    f = zarr_lite.open(cfg.container_path)
    ds = f[cfg.secondary_dataset_name]
    attrs = {"axis_names": ["s", "c"] + ["t", "z", "y", "x"][-nd:], "resolution": (1,) * nd, "offset": (0,) * nd}
    if rank == 0:
        for name, channels, dtype in [(cfg.dataset_name, inference_config.num_bandwidths, np.uint16),
                                      ("binary-segmentation", 1, np.uint16), ("centered-embeddings", nd + 1, float)]:
            d = f.create_dataset(name, shape=(meta.num_samples, channels, *meta.spatial_array), dtype=dtype,
                                 chunks=(1, 1, *meta.spatial_array) if int(np.prod(meta.spatial_array)) < (1 << 24) else None)
            d.attrs.update(attrs)
    if world > 1:
        torch.distributed.barrier()
    ds_detection, ds_binary, ds_centred = f[cfg.dataset_name], f["binary-segmentation"], f["centered-embeddings"]
    if (world > 1 and meta.num_samples < world and inference_config.clustering == "meanshift"
            and not inference_config.use_seeds):
        # fewer samples than GPUs: every sample is split BY SEED over all ranks (north_star: "a single huge
        # volume is split by seed with an NCCL all-gather of the point set") instead of leaving ranks idle
        for sample in range(meta.num_samples):
            _detect_sample_sharded(inference_config, ds, sample, nd, device, rank, world, ds_detection, ds_binary,
                                   ds_centred)
        torch.distributed.barrier()
        return
    for sample in sharding.shard_round_robin(meta.num_samples, rank, world):
        emb = torch.from_numpy(np.ascontiguousarray(ds[sample])).to(device)  # float64, as stored
        threshold = inference_config.threshold
        if threshold is None:
            threshold = threshold_otsu(emb[nd])
        print(f"For sample {sample}, binary threshold {threshold} was used.")
        _, centred = K.centre_embeddings(emb, threshold)
        if inference_config.clustering == "greedy":  # detect.py:162-192
            mask = (emb[nd] < threshold).to(torch.uint8)
            labels = torch.stack([
                K.greedy_cluster(emb, mask, inference_config.bandwidth / (2**k), inference_config.min_size)[0]
                .to(torch.int32).to(torch.uint16) for k in range(inference_config.num_bandwidths)])
        elif inference_config.use_seeds:
            labels, mask = detect_with_seeds(centred, inference_config.bandwidth, threshold,
                                             inference_config.num_bandwidths, inference_config.reduction_probability)
        else:
            labels, _, mask = detect_embeddings(
                emb, inference_config.bandwidth, threshold, inference_config.num_bandwidths,
                inference_config.reduction_probability, rng="numpy", label_dtype=torch.uint16)
        ds_binary[sample, 0, ...] = mask.cpu().numpy().astype(np.uint16)
        ds_centred[sample] = centred.cpu().numpy()
        ds_detection[sample] = labels.cpu().numpy()
    if world > 1:
        torch.distributed.barrier()
