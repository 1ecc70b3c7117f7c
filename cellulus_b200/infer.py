"""`infer(experiment_config)` -- the reference's inference entry point (`cellulus/infer.py:16-80`): defaults
derived from `object_size`, checkpoint loading, then predict -> detect -> segment -> evaluate, each stage
run when its dataset is configured."""

from __future__ import annotations

import os

import numpy as np
import torch

from cellulus_b200.datasets.meta_data import DatasetMetaData
from cellulus_b200.detect import detect
from cellulus_b200.evaluate import evaluate
from cellulus_b200.models import get_model
from cellulus_b200.predict import predict
from cellulus_b200.segment import segment

torch.backends.cudnn.benchmark = True


def _derive_defaults(inference_config, object_size: float, num_spatial_dims: int) -> None:
    """Bandwidth = half the object size; minimum size = a tenth of a disc / ball of that diameter (`:28-39`)."""
    if inference_config.bandwidth is None:
        inference_config.bandwidth = 0.5 * object_size
    if inference_config.min_size is None:
        if num_spatial_dims == 2:
            inference_config.min_size = int(0.1 * np.pi * object_size**2 / 4)
        else:
            inference_config.min_size = int(0.1 * 4.0 / 3.0 * np.pi * object_size**3 / 8)


def _device(inference_config) -> torch.device:
    """The configured CUDA device, or this rank's GPU under torchrun (one process per GPU, NCCL)."""
    from cellulus_b200.utils.device import resolve_device

    device = resolve_device(inference_config.device, "infer")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    return device


def _restore_model(model_config, in_channels: int, num_spatial_dims: int, device: torch.device):
    model = get_model(in_channels, num_spatial_dims, model_config.num_fmaps, model_config.fmap_inc_factor,
                      model_config.features_in_last_layer, [tuple(f) for f in model_config.downsampling_factors],
                      num_spatial_dims).to(device)
    checkpoint = model_config.checkpoint
    assert checkpoint is not None and os.path.exists(checkpoint), \
        f"Model weights do not exist at this location :{checkpoint}!"
    model.load_state_dict(torch.load(checkpoint, map_location=device)["model_state_dict"], strict=True)
    return model.eval()


def infer(experiment_config):
    print(experiment_config)
    inference_config = experiment_config.inference_config
    meta = DatasetMetaData.from_dataset_config(inference_config.dataset_config)
    _derive_defaults(inference_config, experiment_config.object_size, meta.num_spatial_dims)
    model = _restore_model(experiment_config.model_config, meta.num_channels, meta.num_spatial_dims,
                           _device(inference_config))
    first_rank = int(os.environ.get("RANK", "0")) == 0
    # (stage, its dataset, runs on every rank?) -- predict and detect shard their scan blocks / samples
    stages = (
        (lambda: predict(model, inference_config, experiment_config.normalization_factor),
         inference_config.prediction_dataset_config, True),
        (lambda: detect(inference_config), inference_config.detection_dataset_config, True),
        (lambda: segment(inference_config), inference_config.segmentation_dataset_config, False),
        (lambda: evaluate(inference_config), inference_config.evaluation_dataset_config, False),
    )
    for run, dataset, every_rank in stages:
        if dataset is not None and (every_rank or first_rank):
            run()
