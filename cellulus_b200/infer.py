"""`infer(experiment_config)` -- the reference's inference entry point (`cellulus/infer.py:16-80`):
defaults from `object_size`, checkpoint loading, predict -> detect -> segment."""

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


def infer(experiment_config):
    print(experiment_config)
    inference_config = experiment_config.inference_config
    model_config = experiment_config.model_config
    meta = DatasetMetaData.from_dataset_config(inference_config.dataset_config)
    nd = meta.num_spatial_dims

    if inference_config.bandwidth is None:  # infer.py:28-29
        inference_config.bandwidth = 0.5 * experiment_config.object_size
    if inference_config.min_size is None:  # infer.py:31-39
        s = experiment_config.object_size
        inference_config.min_size = int(0.1 * np.pi * s**2 / 4) if nd == 2 else int(0.1 * 4.0 / 3.0 * np.pi * s**3 / 8)

    device = torch.device(inference_config.device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError(f"infer: device={device!s} -- cellulus_b200 has no CPU fallback; use a CUDA device")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    model = get_model(
        in_channels=meta.num_channels, out_channels=nd, num_fmaps=model_config.num_fmaps,
        fmap_inc_factor=model_config.fmap_inc_factor, features_in_last_layer=model_config.features_in_last_layer,
        downsampling_factors=[tuple(f) for f in model_config.downsampling_factors], num_spatial_dims=nd)
    model = model.to(device)
    if model_config.checkpoint is not None and os.path.exists(model_config.checkpoint):
        state = torch.load(model_config.checkpoint, map_location=device)
        model.load_state_dict(state["model_state_dict"], strict=True)
    else:
        assert False, f"Model weights do not exist at this location :{model_config.checkpoint}!"
    model.eval()

    rank = int(os.environ.get("RANK", "0"))
    if inference_config.prediction_dataset_config is not None:
        predict(model, inference_config, experiment_config.normalization_factor)
    if inference_config.detection_dataset_config is not None:
        detect(inference_config)
    if inference_config.segmentation_dataset_config is not None and rank == 0:
        segment(inference_config)
    if inference_config.evaluation_dataset_config is not None and rank == 0:
        evaluate(inference_config)
