"""`train(experiment_config)` -- the reference's training entry point (`cellulus/train.py:16-157`) with the
loss slice of `train_iteration` (`:160-180`) on the B200 kernels.

Per iteration: raw crops come from the DataLoader (host), the U-Net runs in channels-last so its output is
already in the layout the loss kernel gathers from, and the (anchor, reference) pair sampling, gather x2, OCE loss
and backward-to-offsets are ONE kernel (`cb200_oce_loss_sampled`: the pairs are drawn inside it from a counter-based
stream; the reference samples them in DataLoader workers and ships two int64 lists per step over PCIe).  Checkpoints keep the reference's keys
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
from cellulus_b200.utils.device import resolve_device
from cellulus_b200.utils.logger import get_logger

torch.backends.cudnn.benchmark = True


def _join_ranks(spec):
    """(device, rank, world): the configured device with its index resolved (`"cuda"` -> current device); under
    torchrun every rank takes its local GPU and joins the NCCL group."""
    device = resolve_device(spec, "train")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    return device, int(os.environ.get("RANK", "0")), world


class _Checkpoints:
    """The reference's checkpoint dictionary (`train.py:129-157`): same keys, same file names under models/."""

    def __init__(self, model, optimizer, logger):
        self.model, self.optimizer, self.logger = model, optimizer, logger
        self.lowest_loss = 1e6
        self.first_iteration = 0

    def resume(self, path, device):
        print(f"Resuming model from {path}")
        state = torch.load(path, map_location=device)
        self.first_iteration = state["iteration"] + 1
        self.lowest_loss = state["lowest_loss"]
        self.model.load_state_dict(state["model_state_dict"], strict=True)
        self.optimizer.load_state_dict(state["optim_state_dict"])
        self.logger.data = state["logger_data"]

    def save(self, iteration, is_lowest=False):
        save_model({"iteration": iteration, "lowest_loss": self.lowest_loss,
                    "model_state_dict": self.model.state_dict(), "optim_state_dict": self.optimizer.state_dict(),
                    "logger_data": self.logger.data}, iteration, is_lowest)


def train(experiment_config):
    print(experiment_config)
    cfg = experiment_config.train_config
    model_config = experiment_config.model_config
    device, rank, world = _join_ranks(cfg.device)
    main = rank == 0
    if main:
        os.makedirs("models", exist_ok=True)

    dataset = get_dataset(cfg.train_data_config, tuple(cfg.crop_size), cfg.elastic_deform, cfg.control_point_spacing,
                          cfg.control_point_jitter, cfg.density, cfg.kappa, experiment_config.normalization_factor)
    dataset.sample_pairs = False  # the pair lists are drawn on the device, not by the DataLoader workers
    loader = torch.utils.data.DataLoader(dataset=dataset, batch_size=cfg.batch_size, drop_last=True,
                                         num_workers=cfg.num_workers, pin_memory=True)

    nd = dataset.get_num_spatial_dims()
    memory_format = torch.channels_last if nd == 2 else torch.channels_last_3d
    model = get_model(dataset.get_num_channels(), nd, model_config.num_fmaps, model_config.fmap_inc_factor,
                      model_config.features_in_last_layer, [tuple(f) for f in model_config.downsampling_factors], nd)
    model = model.to(device).to(memory_format=memory_format)
    if model_config.initialize:  # train.py:57-61
        for layer in model.modules():
            if isinstance(layer, torch.nn.modules.conv._ConvNd):
                torch.nn.init.kaiming_normal_(layer.weight, nonlinearity="relu")

    criterion = get_loss(cfg.temperature, cfg.regularizer_weight, cfg.density, nd, device)
    optimizer = torch.optim.Adam(model.parameters(), lr=cfg.initial_learning_rate, weight_decay=0.01)
    logger = get_logger(keys=["loss", "oce_loss"], title="loss")
    checkpoints = _Checkpoints(model, optimizer, logger)
    if model_config.checkpoint is not None:
        checkpoints.resume(model_config.checkpoint, device)

    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index]) if world > 1 else model
    running, steps = 0.0, 0  # mean loss since the last "best model" check
    for iteration, raw in zip(range(checkpoints.first_iteration, cfg.max_iterations), loader):
        loss, oce_loss, prediction = train_iteration(
            raw, net, criterion, optimizer, device, dataset, memory_format,
            seed=1_000_003 * rank + iteration, grad_scale=float(world))
        if main:
            print(f"===> loss: {loss:.6f}, oce loss: {oce_loss:.6f}")
            logger.add(key="loss", value=loss)
            logger.add(key="oce_loss", value=oce_loss)
            logger.write()
        running += loss
        steps += 1
        if iteration % cfg.save_best_model_every == 0:
            mean_loss = running / steps
            improved = mean_loss < checkpoints.lowest_loss
            checkpoints.lowest_loss = min(mean_loss, checkpoints.lowest_loss)
            if improved and main:
                checkpoints.save(iteration, is_lowest=True)
            running, steps = 0.0, 0
        last = iteration == cfg.max_iterations - 1
        if main and (iteration % cfg.save_model_every == 0 or last):
            checkpoints.save(iteration)
        if main and iteration % cfg.save_snapshot_every == 0:
            save_snapshot(raw, prediction, iteration)


def train_iteration(raw, model, criterion, optimizer, device, dataset, memory_format, seed, grad_scale=1.0):
    """`train.py:160-180`; returns `(loss, oce_loss, offsets)` like the reference."""
    raw = raw.to(device, non_blocking=True).contiguous(memory_format=memory_format)
    model.train()
    offsets = model(raw)
    if tuple(offsets.shape[2:]) != tuple(dataset.output_shape):
        # the sampler draws pairs for `crop - 16` (hard-coded in the reference, zarr_dataset.py:94); with other
        # downsampling factors the reference fails with an IndexError in its gather -- fail as loudly here
        raise IndexError(f"the model's output shape {tuple(offsets.shape[2:])} differs from the shape the pair "
                         f"sampler assumes {tuple(dataset.output_shape)} (crop_size - 16)")
    # pair sampling + gather x2 + OCE loss + backward in ONE kernel: the lists never exist
    loss, oce_loss, _ = criterion.fused_sampled(offsets, seed=seed, **dataset.pair_stream())
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
