"""`train(experiment_config)` -- the reference's training entry point (`cellulus/train.py:16-157`) with the
loss slice of `train_iteration` (`:160-180`) on the B200 kernels.

Per iteration: raw crops come from the DataLoader (host), the (anchor, reference) pair lists are drawn ON THE
DEVICE (`cb200_sample_pairs`; the reference ships two int64 lists per step over PCIe), the U-Net runs in
channels-last so its output is already in the layout the fused kernel gathers from, and gather x2 + OCE loss
+ backward-to-offsets is one kernel (`oce_loss_fused`).  Checkpoints keep the reference's keys
(`iteration, lowest_loss, model_state_dict, optim_state_dict, logger_data`) and state-dict names.

Multi-GPU: launch with torchrun; every rank trains on its own crops (shard by batch), parameter gradients
are all-reduced by DDP over NCCL.  The loss is a SUM over samples (`criterions/oce_loss.py:58-62`), so the
DDP average is multiplied back by the world size.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from cellulus_b200 import zarr_lite
from cellulus_b200.criterions import get_loss
from cellulus_b200.datasets import get_dataset
from cellulus_b200.models import get_model
from cellulus_b200.utils.logger import get_logger

torch.backends.cudnn.benchmark = True


def _require_cuda(device: torch.device, what: str):
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError(
            f"{what}: device={device!s} -- cellulus_b200 runs its loss / detection kernels on a CUDA device only "
            "(there is no CPU fallback); set `device = \"cuda:0\"` in the config")


def train(experiment_config):
    print(experiment_config)
    train_config = experiment_config.train_config
    model_config = experiment_config.model_config
    device = torch.device(train_config.device)
    _require_cuda(device, "train")

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    else:
        torch.cuda.set_device(device)
    if rank == 0 and not os.path.exists("models"):
        os.makedirs("models")

    train_dataset = get_dataset(
        dataset_config=train_config.train_data_config,
        crop_size=tuple(train_config.crop_size),
        elastic_deform=train_config.elastic_deform,
        control_point_spacing=train_config.control_point_spacing,
        control_point_jitter=train_config.control_point_jitter,
        density=train_config.density,
        kappa=train_config.kappa,
        normalization_factor=experiment_config.normalization_factor,
    )
    train_dataset.sample_pairs = False  # pair lists are drawn on the device
    train_dataloader = torch.utils.data.DataLoader(
        dataset=train_dataset, batch_size=train_config.batch_size, drop_last=True,
        num_workers=train_config.num_workers, pin_memory=True)

    nd = train_dataset.get_num_spatial_dims()
    model = get_model(
        in_channels=train_dataset.get_num_channels(), out_channels=nd, num_fmaps=model_config.num_fmaps,
        fmap_inc_factor=model_config.fmap_inc_factor, features_in_last_layer=model_config.features_in_last_layer,
        downsampling_factors=[tuple(f) for f in model_config.downsampling_factors], num_spatial_dims=nd)
    memory_format = torch.channels_last if nd == 2 else torch.channels_last_3d
    model = model.to(device).to(memory_format=memory_format)
    if model_config.initialize:
        for _name, layer in model.named_modules():
            if isinstance(layer, torch.nn.modules.conv._ConvNd):
                torch.nn.init.kaiming_normal_(layer.weight, nonlinearity="relu")

    criterion = get_loss(regularizer_weight=train_config.regularizer_weight, temperature=train_config.temperature,
                         density=train_config.density, num_spatial_dims=nd, device=device)
    optimizer = torch.optim.Adam(model.parameters(), lr=train_config.initial_learning_rate, weight_decay=0.01)
    logger = get_logger(keys=["loss", "oce_loss"], title="loss")

    start_iteration, lowest_loss, epoch_loss, num_iterations = 0, 1e6, 0, 0
    if model_config.checkpoint is not None:
        print(f"Resuming model from {model_config.checkpoint}")
        state = torch.load(model_config.checkpoint, map_location=device)
        start_iteration = state["iteration"] + 1
        lowest_loss = state["lowest_loss"]
        model.load_state_dict(state["model_state_dict"], strict=True)
        optimizer.load_state_dict(state["optim_state_dict"])
        logger.data = state["logger_data"]

    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index])

    def state_dict(iteration):
        return {"iteration": iteration, "lowest_loss": lowest_loss, "model_state_dict": model.state_dict(),
                "optim_state_dict": optimizer.state_dict(), "logger_data": logger.data}

    for iteration, raw in zip(range(start_iteration, train_config.max_iterations), train_dataloader):
        loss, oce_loss, prediction = train_iteration(
            raw, net, criterion, optimizer, device, train_dataset, memory_format,
            seed=1_000_003 * rank + iteration, grad_scale=float(world))
        if rank == 0:
            print(f"===> loss: {loss:.6f}, oce loss: {oce_loss:.6f}")
            logger.add(key="loss", value=loss)
            logger.add(key="oce_loss", value=oce_loss)
            logger.write()
        epoch_loss += loss
        num_iterations += 1
        if iteration % train_config.save_best_model_every == 0:
            is_lowest = epoch_loss / num_iterations < lowest_loss
            lowest_loss = min(epoch_loss / num_iterations, lowest_loss)
            if is_lowest and rank == 0:
                save_model(state_dict(iteration), iteration, is_lowest)
            epoch_loss, num_iterations = 0, 0
        if (iteration % train_config.save_model_every == 0 or iteration == train_config.max_iterations - 1) and rank == 0:
            save_model(state_dict(iteration), iteration)
        if iteration % train_config.save_snapshot_every == 0 and rank == 0:
            save_snapshot(raw, prediction, iteration)


def train_iteration(raw, model, criterion, optimizer, device, dataset, memory_format, seed, grad_scale=1.0):
    """`train.py:160-180`; returns `(loss, oce_loss, offsets)` like the reference."""
    raw = raw.to(device, non_blocking=True).contiguous(memory_format=memory_format)
    anchors, refs = dataset.sample_coordinates_device(raw.shape[0], device, seed)
    model.train()
    offsets = model(raw)
    loss, oce_loss, _ = criterion.fused(offsets, anchors, refs)  # gather x2 + OCE loss + backward in one kernel
    optimizer.zero_grad()
    (loss * grad_scale if grad_scale != 1.0 else loss).backward()
    optimizer.step()
    return loss.item(), oce_loss.item(), offsets


def save_model(state, iteration, is_lowest=False):
    name = "best_loss.pth" if is_lowest else str(iteration).zfill(6) + ".pth"
    torch.save(state, os.path.join("models", name))
    print(("Best model weights" if is_lowest else "Checkpoint") + f" saved at iteration {iteration}")


def save_snapshot(raw, prediction, iteration):
    """`train.py:194-224`: raw + mean-subtracted offsets into snapshots.zarr."""
    nd = raw.ndim - 2
    axis_names = ["s", "c"] + ["t", "z", "y", "x"][-nd:]
    f = zarr_lite.open("snapshots.zarr", "a")
    f[f"{iteration}/raw"] = raw.detach().cpu().numpy()
    f[f"{iteration}/raw"].attrs.update({"axis_names": axis_names, "resolution": [1] * nd})
    pred = prediction.detach().float().cpu().numpy()
    pred = pred - pred.reshape(pred.shape[0], pred.shape[1], -1).mean(2)[(...,) + (np.newaxis,) * nd]
    f[f"{iteration}/prediction"] = pred
    f[f"{iteration}/prediction"].attrs.update({
        "axis_names": axis_names, "resolution": [1] * nd,
        "offset": [(a - b) / 2 for a, b in zip(raw.shape[-nd:], prediction.shape[-nd:])]})
