"""`predict(model, inference_config, normalization_factor)` (`cellulus/predict.py:9-142`) without gunpowder:
the volume is walked in scan blocks (`crop_size` in, valid-convolution output out, reflect-padded by the
context, last block shifted inward -- `cellulus_b200.sharding.scan_blocks`), each block goes through the
model's infer-mode forward (TTA loop + aggregate resident on the device) and lands in the `embeddings`
dataset `(s, D+1, *spatial)` float64 with the reference's attributes.

Multi-GPU (torchrun): samples are dealt to ranks round-robin; scan blocks of one sample stay on one rank so
that no two ranks ever write the same chunk.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from cellulus_b200.utils.device import resolve_device
from cellulus_b200 import sharding, zarr_lite
from cellulus_b200.datasets.meta_data import DatasetMetaData


def _normalize(data: np.ndarray, factor):
    if factor is None:
        factor = {np.dtype(np.uint8): 1.0 / 255, np.dtype(np.uint16): 1.0 / 65535}.get(data.dtype, 1.0)
    return data.astype(np.float32) * np.float32(factor)


def predict(model: torch.nn.Module, inference_config, normalization_factor) -> None:
    dataset_config = inference_config.dataset_config
    meta = DatasetMetaData.from_dataset_config(dataset_config)
    device = resolve_device(inference_config.device, "predict")
    nd = meta.num_spatial_dims
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    model.set_infer(p_salt_pepper=inference_config.p_salt_pepper,
                    num_infer_iterations=inference_config.num_infer_iterations, device=device)
    crop = tuple(inference_config.crop_size)
    with torch.no_grad():
        out_shape = tuple(model(torch.zeros((1, meta.num_channels, *crop), device=device)).shape[2:])
    context = tuple((c - o) // 2 for c, o in zip(crop, out_shape))

    raw_ds = zarr_lite.open(dataset_config.container_path, "r")[dataset_config.dataset_name]
    f = zarr_lite.open(inference_config.prediction_dataset_config.container_path)
    name = inference_config.prediction_dataset_config.dataset_name
    if rank == 0:
        ds = f.create_dataset(name, shape=(meta.num_samples, nd + 1, *meta.spatial_array), dtype=float,
                              chunks=(1, nd + 1, *meta.spatial_array) if int(np.prod(meta.spatial_array)) < (1 << 24) else None)
        ds.attrs.update({"axis_names": ["s", "c"] + ["t", "z", "y", "x"][-nd:], "resolution": (1,) * nd,
                         "offset": (0,) * nd})
    if world > 1:
        torch.distributed.barrier()
    ds = f[name]

    blocks = sharding.scan_blocks(meta.spatial_array, out_shape)
    with torch.no_grad():
        for sample in sharding.shard_round_robin(meta.num_samples, rank, world):
            raw = _normalize(np.asarray(raw_ds[sample]), normalization_factor)  # (c, *spatial)
            # a volume smaller than one output block is padded up to it (the reference's Scan would fail)
            pad = [(0, 0)] + [(ctx, ctx + max(0, o - s)) for ctx, o, s in zip(context, out_shape, meta.spatial_array)]
            raw = torch.from_numpy(np.pad(raw, pad, mode="reflect")).to(device)
            result = torch.empty((nd + 1, *meta.spatial_array), dtype=torch.float32, device=device)
            for off in blocks:
                src = (slice(None),) + tuple(slice(o, o + c) for o, c in zip(off, crop))
                emb = model(raw[src][None])[0]  # (D+1, *out_shape), on the device
                dst = (slice(None),) + tuple(slice(o, min(o + b, s)) for o, b, s in zip(off, out_shape, meta.spatial_array))
                cut = (slice(None),) + tuple(slice(0, d.stop - d.start) for d in dst[1:])
                result[dst] = emb[cut]
            ds[sample] = result.double().cpu().numpy()
    if world > 1:
        torch.distributed.barrier()
