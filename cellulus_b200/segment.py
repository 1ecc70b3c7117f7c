"""`segment(inference_config)` (`cellulus/segment.py:13-108`): post-processing of the detection.

All three pixel-heavy steps run on the device: the grow / shrink of "cell" (`cb200_grow_shrink`: two
thresholded Euclidean distance transforms), the per-instance Otsu + hole filling of "nucleus"
(`cb200_label_stats`, `cb200_label_histogram`, `cb200_label_otsu`, `cb200_nucleus_fill`) and `size_filter` (`cb200_size_filter`, `utils/misc.py:11-25`).
"""

from __future__ import annotations

import numpy as np

from cellulus_b200 import zarr_lite
from cellulus_b200.datasets.meta_data import DatasetMetaData
from cellulus_b200.utils.misc import size_filter


def _bin_edges(lo, hi, dtype, nbins=256):
    """The edges `np.histogram(values, nbins)` builds for values of `dtype` spanning [lo, hi]."""
    first, last = dtype.type(lo), dtype.type(hi)
    return np.linspace(first, last, nbins + 1, endpoint=True, dtype=np.result_type(first, last, dtype))


def nucleus(segmentation: np.ndarray, raw_image: np.ndarray, device="cuda") -> np.ndarray:
    """`segment.py:52-101` on the device; numpy in, numpy out (a new array like the reference's dataset)."""
    import torch

    from cellulus_b200 import kernels as K

    seg_np = np.ascontiguousarray(segmentation)
    out_dtype = seg_np.dtype
    raw_np = np.ascontiguousarray(raw_image)
    if raw_np.dtype not in (np.float32, np.float64, np.uint8, np.uint16):
        raise TypeError(f"nucleus post-processing: raw dtype {raw_np.dtype} is not supported "
                        "(float32, float64, uint8, uint16)")
    seg = torch.from_numpy(seg_np.astype(np.int32)).to(device)
    raw = torch.from_numpy(raw_np).to(device)
    max_label = int(seg.max().item()) if seg.numel() else 0
    if max_label <= 0:
        return np.zeros_like(seg_np)
    mn_t, mx_t, box_t = K.label_stats(seg, raw, max_label)
    mn, mx, box = mn_t.cpu().numpy(), mx_t.cpu().numpy(), box_t.cpu().numpy()
    ids = np.nonzero(box[:, 3] >= 0)[0]
    ids = ids[ids != 0]
    integer = np.issubdtype(raw_np.dtype, np.integer)
    present = np.zeros(max_label + 1, bool)
    present[ids] = True
    varied = present & (mn != mx)  # a constant instance: skimage returns that value (set below)
    if integer:  # skimage: one bin per value in [min, max] of the instance
        span = np.where(present, (mx - mn).astype(np.int64) + 1, 0)
        offset = np.concatenate([[0], np.cumsum(span)[:-1]]).astype(np.int64)
        total_bins = int(span.sum())
        offset_t = torch.from_numpy(offset).to(device)
        hist = torch.zeros(total_bins, dtype=torch.int32, device=device)
        K.label_histogram(seg, raw, max_label, hist, raw_min=mn_t, hist_offset=offset_t)
        thresholds_t = K.label_otsu(hist, offset_t, torch.from_numpy(span).to(device), total_bins, torch.float64,
                                    centre0=mn_t)
    else:  # np.histogram, 256 bins over [min, max] of the instance, edges in the image's dtype
        nbins = 256
        dt = raw_np.dtype
        edges = np.zeros((max_label + 1, nbins + 1), dt)
        sel = np.nonzero(varied)[0]
        if len(sel):
            edges[sel] = np.linspace(mn[sel].astype(dt), mx[sel].astype(dt), nbins + 1, endpoint=True, dtype=dt, axis=-1)
        centres = (edges[:, :-1] + edges[:, 1:]) / 2.0  # stays in the image's dtype, like skimage's bin centres
        hist = torch.zeros((max_label + 1) * nbins, dtype=torch.int32, device=device)
        K.label_histogram(seg, raw, max_label, hist, edges=torch.from_numpy(edges.astype(np.float64)).to(device),
                          nbins=nbins)
        offset_t = torch.arange(max_label + 1, dtype=torch.int64, device=device) * nbins
        num_bins = torch.from_numpy(np.where(varied, nbins, 0).astype(np.int64)).to(device)
        thresholds_t = K.label_otsu(hist, offset_t, num_bins, (max_label + 1) * nbins,
                                    torch.float32 if dt == np.float32 else torch.float64,
                                    centres=torch.from_numpy(centres.astype(np.float64)).to(device))
        thresholds_t = torch.where(torch.from_numpy(varied).to(device), thresholds_t, mn_t)
    thresholds = thresholds_t[torch.from_numpy(ids).to(device)].contiguous()
    boxes = box[ids].astype(np.int32)
    volumes = np.prod(boxes[:, 3:].astype(np.int64) - boxes[:, :3] + 1, axis=1)
    box_offset = np.concatenate([[0], np.cumsum(volumes)]).astype(np.int64)
    out = K.nucleus_fill(seg, raw, torch.from_numpy(ids.astype(np.int32)).to(device),
                         thresholds, torch.from_numpy(boxes).to(device),
                         torch.from_numpy(box_offset).to(device), int(box_offset[-1]))
    return out.cpu().numpy().astype(out_dtype)


def grow_shrink(segmentation: np.ndarray, grow_distance, shrink_distance, device="cuda") -> np.ndarray:
    """`segment.py:46-50` (two Euclidean distance transforms and their thresholds) on the device
    (`cb200_grow_shrink`); numpy in, numpy out, `segmentation` is modified in place like the reference's."""
    import torch

    from cellulus_b200 import kernels as K

    seg = torch.from_numpy(np.ascontiguousarray(segmentation).astype(np.int32)).to(device)
    K.grow_shrink_(seg, grow_distance, shrink_distance)
    segmentation[...] = seg.cpu().numpy().astype(segmentation.dtype)
    return segmentation


def segment(inference_config) -> None:
    meta = DatasetMetaData.from_dataset_config(inference_config.dataset_config)
    nd = meta.num_spatial_dims
    cfg = inference_config.segmentation_dataset_config
    f = zarr_lite.open(cfg.container_path)
    ds = f[cfg.secondary_dataset_name]
    ds_segmented = f.create_dataset(
        cfg.dataset_name, shape=(meta.num_samples, inference_config.num_bandwidths, *meta.spatial_array),
        dtype=np.uint16,
        chunks=(1, 1, *meta.spatial_array) if int(np.prod(meta.spatial_array)) < (1 << 24) else None)
    ds_segmented.attrs.update({"axis_names": ["s", "c"] + ["t", "z", "y", "x"][-nd:], "resolution": (1,) * nd,
                               "offset": (0,) * nd})
    ds_raw = None
    if inference_config.post_processing == "nucleus":
        ds_raw = zarr_lite.open(inference_config.dataset_config.container_path, "r")[
            inference_config.dataset_config.dataset_name]
    for sample in range(meta.num_samples):
        for k in range(inference_config.num_bandwidths):
            segmentation = np.asarray(ds[sample, k])
            if inference_config.post_processing == "cell":  # segment.py:41-51, on the device
                out = grow_shrink(segmentation, inference_config.grow_distance, inference_config.shrink_distance)
            else:  # "nucleus", segment.py:52-101, on the device
                out = nucleus(segmentation, np.asarray(ds_raw[sample, 0]))
            # size filter: remove small objects (segment.py:104-108) -- device connected components
            ds_segmented[sample, k, ...] = size_filter(out, inference_config.min_size).astype(np.uint16)
