"""`segment(inference_config)` (`cellulus/segment.py:13-108`): post-processing of the detection.

Only `size_filter` (`utils/misc.py:11-25`) is on the hot path and runs on the device
(`cb200_size_filter`).  The morphological grow/shrink ("cell") and per-instance Otsu + hole filling
("nucleus") are SURVEY §8f "next" rows; they are carried out here with scipy on the host exactly as the
reference writes them so that `infer()` completes end to end.
"""

from __future__ import annotations

import numpy as np
from scipy.ndimage import binary_fill_holes
from scipy.ndimage import distance_transform_edt as dtedt

from cellulus_b200 import zarr_lite
from cellulus_b200.datasets.meta_data import DatasetMetaData
from cellulus_b200.utils.misc import size_filter


def _otsu(values: np.ndarray):
    from cellulus_b200.detect import otsu_from_histogram

    first = values.reshape(-1)[0]
    if np.all(values == first):
        return first
    counts, edges = np.histogram(values.reshape(-1), 256)
    return otsu_from_histogram(counts, edges)


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
            if inference_config.post_processing == "cell":  # segment.py:41-51
                expanded = dtedt(segmentation == 0) < inference_config.grow_distance
                segmentation[dtedt(expanded) < inference_config.shrink_distance] = 0
                out = segmentation
            else:  # "nucleus", segment.py:52-101
                out = np.zeros_like(segmentation)
                raw_image = np.asarray(ds_raw[sample, 0])
                for id_ in np.unique(segmentation):
                    if id_ == 0:
                        continue
                    m = segmentation == id_
                    idx = np.where(m)
                    box = tuple(slice(int(i.min()), int(i.max()) + 1) for i in idx)
                    mask = m & (raw_image > _otsu(raw_image[m]))
                    mask[box] = binary_fill_holes(mask[box])
                    out[mask] = id_
            # size filter: remove small objects (segment.py:104-108) -- device connected components
            ds_segmented[sample, k, ...] = size_filter(out, inference_config.min_size).astype(np.uint16)
