"""Device resolution shared by `train`, `infer`, `predict` and `detect`.

The reference accepts any torch device string in its configs (`configs/train_config.py`, `inference_config.py`:
`device = "cuda:0"` by default, `"cuda"` works too).  The kernels here need a CONCRETE CUDA device: the C ABI
launches on the calling thread's current device, `torch.cuda.set_device` and DDP's `device_ids` want an index.
"""

from __future__ import annotations

import os

import torch


def resolve_device(spec, what: str = "cellulus_b200", set_current: bool = True) -> torch.device:
    """`torch.device(spec)` with the index filled in (`"cuda"` -> the current device); under torchrun the rank's
    own GPU (`LOCAL_RANK`) wins over the configured index.  Raises for non-CUDA devices: no CPU fallback."""
    device = torch.device(spec)
    if device.type != "cuda":
        raise RuntimeError(
            f"{what}: device={device!s} -- cellulus_b200 runs its loss / detection kernels on a CUDA device only "
            "(there is no CPU fallback); set `device = \"cuda:0\"` in the config")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    elif device.index is None:
        index = torch.cuda.current_device() if torch.cuda.is_available() else 0
        device = torch.device("cuda", index)
    if set_current:
        if not torch.cuda.is_available():
            raise RuntimeError(f"{what}: device={device!s} requested but no CUDA device is available")
        torch.cuda.set_device(device)
    return device
