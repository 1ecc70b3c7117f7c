"""The two hot-path pieces that live on the reference's model class
(`cellulus/models/unet.py`): the neighbour gather `select_and_add_coordinates`
(:108-124) and the test-time-augmentation aggregate of the infer-mode forward
(:90-98).  The U-Net backbone itself is not the product and is not here.
"""

from __future__ import annotations

import torch

from cellulus_b200 import kernels as K


class _GatherAddCoords(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, coordinates):
        ctx.save_for_backward(coordinates)
        ctx.shape = tuple(outputs.shape)
        ctx.in_dtype = outputs.dtype
        return K.gather_add_coords(outputs, coordinates)

    @staticmethod
    def backward(ctx, grad_out):
        (coordinates,) = ctx.saved_tensors
        return K.scatter_add_coords(grad_out, coordinates, ctx.shape).to(ctx.in_dtype), None


class UNetModel:
    """Namespace mirror of the reference class for the static gather
    (`train.py:170-173` calls `model.select_and_add_coordinates(...)`)."""

    @staticmethod
    def select_and_add_coordinates(outputs, coordinates):
        """`models/unet.py:108-124`: outputs (B,C,H,W)/(B,C,D,H,W), coordinates
        (B,P,D) integer with columns (x, y[, z]) -> (B,P,C) fp32 = gathered offset
        + coordinate.  Differentiable w.r.t. `outputs` (scatter-add backward)."""
        return _GatherAddCoords.apply(outputs, coordinates)


def tta_aggregate(predictions: torch.Tensor) -> torch.Tensor:
    """`models/unet.py:90-98`: (T, C, *S) fp32 stack of noisy predictions ->
    (C+1, *S): channel means, then the per-channel population std summed over
    channels.  One streaming kernel instead of stack + std_mean + sum + cat."""
    return K.tta_aggregate(predictions)


class TTAAccumulator:
    """Streaming form for an on-device TTA loop: fold each prediction in as it
    is produced (no T-deep stack, no `.cpu()` per pass as in `models/unet.py:83-88`)."""

    def __init__(self, channels: int, spatial, device):
        self.channels = channels
        self.spatial = tuple(int(s) for s in spatial)
        self.state = torch.empty((2 * channels, *self.spatial), dtype=torch.float32, device=device)
        self.t = 0

    def add(self, prediction: torch.Tensor) -> None:
        assert tuple(prediction.shape) == (self.channels, *self.spatial) and prediction.dtype == torch.float32
        K.tta_accumulate(self.state, prediction, self.t)
        self.t += 1

    def result(self) -> torch.Tensor:
        if self.t == 0:
            raise RuntimeError("no prediction was accumulated")
        return K.tta_finalize(self.state, self.t, self.channels, self.spatial)
