"""Oracle: neighbour gather + object-centric-embedding loss (torch CPU).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows
`cellulus/models/unet.py:108-124` and `cellulus/criterions/oce_loss.py:45-63`.
"""

from __future__ import annotations

import torch


def select_and_add_coordinates(outputs: torch.Tensor, coordinates: torch.Tensor) -> torch.Tensor:
    """`cellulus/models/unet.py:108-124`.

    `outputs` (B,C,H,W) or (B,C,D,H,W); `coordinates` (B,P,D) int64 with
    column 0 = x (last axis).  Gathers the per-pixel offsets, transposes to
    (P,C), adds the integer coordinate into the float selection in place
    (channel k += coordinate column k) and stacks to (B,P,C).
    """
    selections = []
    for output, coordinate in zip(outputs, coordinates):
        if output.ndim == 3:
            selection = output[:, coordinate[:, 1], coordinate[:, 0]]
        elif output.ndim == 4:
            selection = output[:, coordinate[:, 2], coordinate[:, 1], coordinate[:, 0]]
        else:  # the reference would raise UnboundLocalError here
            raise ValueError("outputs must be (B,C,H,W) or (B,C,D,H,W)")
        selection = selection.transpose(1, 0)
        selection += coordinate
        selections.append(selection)
    return torch.stack(selections, dim=0)


def distance_function(e0: torch.Tensor, e1: torch.Tensor) -> torch.Tensor:
    """`cellulus/criterions/oce_loss.py:45-48`."""
    return (e0 - e1).norm(2, dim=-1)


def oce_loss(anchor_embedding, reference_embedding, temperature, regularization_weight):
    """`cellulus/criterions/oce_loss.py:53-63`: returns (loss, oce, reg), all sums."""
    distance = distance_function(anchor_embedding, reference_embedding.detach())
    non_linear_distance = 1 - (-distance.pow(2) / temperature).exp()  # :50-51
    oce = non_linear_distance.sum()
    reg = regularization_weight * anchor_embedding.norm(2, dim=-1).sum()
    return oce + reg, oce, reg


def loss_step(offsets, anchor_coordinates, reference_coordinates, temperature, regularization_weight):
    """The loss slice of `cellulus/train.py:169-178`: gather x2, criterion,
    backward.  Returns `(loss, oce, reg, d loss / d offsets)`."""
    offsets = offsets.detach().clone().requires_grad_(True)
    ea = select_and_add_coordinates(offsets, anchor_coordinates)
    er = select_and_add_coordinates(offsets, reference_coordinates)
    loss, oce, reg = oce_loss(ea, er, temperature, regularization_weight)
    loss.backward()
    return loss.detach(), oce.detach(), reg.detach(), offsets.grad


def loss_step_float64(offsets, anchor_coordinates, reference_coordinates, temperature, regularization_weight):
    """Same arithmetic carried in float64 on the fp32 inputs: the "true" value
    both the fp32 reference and the fp32 CUDA kernel approximate.  Used to show
    the kernel is at least as close to the exact result as the reference is."""
    return loss_step(
        offsets.double(), anchor_coordinates, reference_coordinates, temperature, regularization_weight
    )
