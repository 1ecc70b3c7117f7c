"""`predict(model, inference_config, normalization_factor)` (`cellulus/predict.py:9-142`) without gunpowder:
the volume is walked in scan blocks (`crop_size` in, valid-convolution output out, reflect-padded by the
context, last block shifted inward -- `cellulus_b200.sharding.scan_blocks`), each block goes through the
model's infer-mode forward (TTA loop + aggregate resident on the device) and lands in the `embeddings`
dataset `(s, D+1, *spatial)` float64 with the reference's attributes.

Multi-GPU (torchrun): samples are dealt to ranks round-robin and the scan blocks of one sample stay on one rank
(no two ranks ever write the same chunk) -- unless there are fewer samples than ranks (one mosaic, one volume:
BASELINE configs[4]): then the scan blocks of EVERY sample are dealt to the ranks round-robin, each rank fills the
part of its blocks that no later block of the scan overwrites (`sharding.owned_extents`), the partial volumes are
summed onto rank 0 over NCCL and rank 0 writes the dataset.
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

    blocks = sharding.owned_extents(meta.spatial_array, out_shape)
    share_samples = world > 1 and meta.num_samples < world  # fewer samples than ranks: deal the scan blocks instead
    samples = range(meta.num_samples) if share_samples else sharding.shard_round_robin(meta.num_samples, rank, world)
    with torch.no_grad():
        for sample in samples:
            raw = _normalize(np.asarray(raw_ds[sample]), normalization_factor)  # (c, *spatial)
            # a volume smaller than one output block is padded up to it (the reference's Scan would fail)
            pad = [(0, 0)] + [(ctx, ctx + max(0, o - s)) for ctx, o, s in zip(context, out_shape, meta.spatial_array)]
            raw = torch.from_numpy(np.pad(raw, pad, mode="reflect")).to(device)
            alloc = torch.zeros if share_samples else torch.empty
            result = alloc((nd + 1, *meta.spatial_array), dtype=torch.float32, device=device)
            mine = blocks[rank::world] if share_samples else blocks
            for off, owned in mine:
                src = (slice(None),) + tuple(slice(o, o + c) for o, c in zip(off, crop))
                emb = model(raw[src][None])[0]  # (D+1, *out_shape), on the device
                dst = (slice(None),) + tuple(slice(o, o + e) for o, e in zip(off, owned))
                cut = (slice(None),) + tuple(slice(0, e) for e in owned)
                result[dst] = emb[cut]
            if share_samples:
                torch.distributed.reduce(result, dst=0)
            if not share_samples or rank == 0:
                ds[sample] = result.double().cpu().numpy()
    if world > 1:
        torch.distributed.barrier()
