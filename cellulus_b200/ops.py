"""`torch.library` registration of the hot-path kernels: `torch.ops.cellulus_b200.*`.

The C ABI (`libcellulus_b200.so`, bound with ctypes in `_cabi.py`) stays the loader; this module makes the
same calls visible to the PyTorch dispatcher -- schema, CUDA implementation, fake (meta) implementation and
autograd formula -- so that they can be traced (`torch.compile`, `torch.export`), checked with
`torch.library.opcheck`, and called as `torch.ops.cellulus_b200.<op>`:

    oce_loss_fused(offsets, anchor_coordinates, reference_coordinates, temperature, regularization_weight)
        -> (result (4,) = [loss, oce_loss, regularization_loss, skipped pairs], d loss / d offsets)
        replaces cellulus/train.py:169-178 (gather x2, OCELoss.forward, backward to the offsets)
    oce_loss_sampled(offsets, kappa, num_anchors, num_references, seed, sequence, temperature, regularization_weight)
        -> same pair of tensors; the pairs are drawn inside the kernel (zarr_dataset.py:177-251 as well)
    tta_aggregate(stack (T, C, *S)) -> (C + 1, *S)                      cellulus/models/unet.py:90-98
    detect_volume(embeddings (D + 1, *S), bandwidth, threshold, reduction_probability, philox_seed)
        -> int32 labels (*S)                                            cellulus/utils/mean_shift.py:6-45

Autograd: `result` is differentiable w.r.t. `offsets` in all of its first three entries (exact: the gradient
of the data term alone is recomputed with the regulariser switched off when it is asked for).  The eager entry
points `criterions.oce_loss_fused` / `oce_loss_fused_sampled` keep their `autograd.Function` fast path, which
skips that recomputation for the usual `loss.backward()`; they compute the same numbers with the same kernels.
"""

from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from cellulus_b200 import kernels as K

_NS = "cellulus_b200"


# ------------------------------------------------------------------------------------------ loss, explicit lists
@torch.library.custom_op(f"{_NS}::oce_loss_fused", mutates_args=(), device_types="cuda")
def oce_loss_fused(offsets: Tensor, anchor_coordinates: Tensor, reference_coordinates: Tensor, temperature: float,
                   regularization_weight: float) -> Tuple[Tensor, Tensor]:
    out, grad = K.oce_loss_fwd_bwd(offsets, anchor_coordinates, reference_coordinates, temperature,
                                   regularization_weight, want_grad=True)
    return out, grad


@oce_loss_fused.register_fake
def _(offsets, anchor_coordinates, reference_coordinates, temperature, regularization_weight):
    return offsets.new_empty((4,), dtype=torch.float32), torch.empty_like(offsets, dtype=torch.float32)


def _combine(g_result, grad_loss, grad_oce, in_dtype):
    """d/d offsets of sum_i g_result[i] * result[i]: loss = oce + reg, so d reg = d loss - d oce."""
    g = g_result.to(torch.float32)
    total = (g[0] + g[1]) * grad_oce + (g[0] + g[2]) * (grad_loss - grad_oce)
    return total.to(in_dtype)


def _fused_setup(ctx, inputs, output):
    offsets, anchors, refs, temperature, weight = inputs
    ctx.save_for_backward(output[1], offsets, anchors, refs)
    ctx.args = (temperature, weight)


def _fused_backward(ctx, g_result, _g_grad):
    grad_loss, offsets, anchors, refs = ctx.saved_tensors
    temperature, _ = ctx.args
    if g_result is None:
        return None, None, None, None, None
    # the data term alone: the same kernel with the regulariser switched off
    _, grad_oce = torch.ops.cellulus_b200.oce_loss_fused(offsets.detach(), anchors, refs, temperature, 0.0)
    return _combine(g_result, grad_loss, grad_oce, offsets.dtype), None, None, None, None


oce_loss_fused.register_autograd(_fused_backward, setup_context=_fused_setup)


# ------------------------------------------------------------------------------------------ loss, pairs drawn in the kernel
@torch.library.custom_op(f"{_NS}::oce_loss_sampled", mutates_args=(), device_types="cuda")
def oce_loss_sampled(offsets: Tensor, kappa: float, num_anchors: int, num_references: int, seed: int, sequence: int,
                     temperature: float, regularization_weight: float) -> Tuple[Tensor, Tensor]:
    out, grad, _ = K.oce_loss_sampled(offsets, kappa, num_anchors, num_references, seed, sequence, temperature,
                                      regularization_weight, want_grad=True)
    return out, grad


@oce_loss_sampled.register_fake
def _(offsets, kappa, num_anchors, num_references, seed, sequence, temperature, regularization_weight):
    return offsets.new_empty((4,), dtype=torch.float32), torch.empty_like(offsets, dtype=torch.float32)


def _sampled_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1], inputs[0])
    ctx.args = inputs[1:]


def _sampled_backward(ctx, g_result, _g_grad):
    grad_loss, offsets = ctx.saved_tensors
    kappa, na, nr, seed, sequence, temperature, _ = ctx.args
    none = (None,) * 7
    if g_result is None:
        return (None,) + none
    _, grad_oce = torch.ops.cellulus_b200.oce_loss_sampled(offsets.detach(), kappa, na, nr, seed, sequence, temperature, 0.0)
    return (_combine(g_result, grad_loss, grad_oce, offsets.dtype),) + none


oce_loss_sampled.register_autograd(_sampled_backward, setup_context=_sampled_setup)


# ------------------------------------------------------------------------------------------ inference
@torch.library.custom_op(f"{_NS}::tta_aggregate", mutates_args=(), device_types="cuda")
def tta_aggregate(stack: Tensor) -> Tensor:
    return K.tta_aggregate(stack)


@tta_aggregate.register_fake
def _(stack):
    return stack.new_empty((stack.shape[1] + 1, *stack.shape[2:]))


@torch.library.custom_op(f"{_NS}::detect_volume", mutates_args=(), device_types="cuda")
def detect_volume(embeddings: Tensor, bandwidth: float, threshold: float, reduction_probability: float,
                  philox_seed: int) -> Tensor:
    labels, _, _, _ = K.detect_volume(embeddings, bandwidth, threshold, reduction_probability, philox_seed)
    return labels


@detect_volume.register_fake
def _(embeddings, bandwidth, threshold, reduction_probability, philox_seed):
    return embeddings.new_empty(tuple(embeddings.shape[1:]), dtype=torch.int32)
