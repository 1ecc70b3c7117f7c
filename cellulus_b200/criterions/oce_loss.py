"""Object-centric-embedding loss on the B200 kernels.

`OCELoss` keeps the reference's constructor and `forward(anchor_embedding,
reference_embedding) -> (loss, oce_loss, regularization_loss)` contract
(`cellulus/criterions/oce_loss.py:5-63`); `oce_loss_fused` replaces the three
calls of `cellulus/train.py:169-176` (gather, gather, criterion) with ONE
kernel that also produces the gradient w.r.t. the offsets.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from cellulus_b200 import kernels as K


class _PairLoss(torch.autograd.Function):
    """criterions/oce_loss.py:53-63 on materialised (.., D) embeddings; gradient
    flows to the anchor side only (the reference side is `.detach()`ed, :55)."""

    @staticmethod
    def forward(ctx, anchor_embedding, reference_embedding, temperature, regularization_weight):
        ctx.set_materialize_grads(False)
        need = ctx.needs_input_grad[0]
        out, grad = K.oce_pair_loss(anchor_embedding, reference_embedding, temperature, regularization_weight, need)
        ctx.args = (temperature, regularization_weight)
        ctx.in_dtype = anchor_embedding.dtype
        ctx.save_for_backward(grad, anchor_embedding, reference_embedding)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_loss, g_oce, g_reg):
        grad, ea, er = ctx.saved_tensors
        T, w = ctx.args
        if grad is None:
            return None, None, None, None
        total = _combine(grad, g_loss, g_oce, g_reg,
                         lambda ww: K.oce_pair_loss(ea, er, T, ww, True)[1])
        return total.reshape(ea.shape).to(ctx.in_dtype), None, None, None


def _combine(grad, g_loss, g_oce, g_reg, recompute):
    """d/d(input) of g_loss*loss + g_oce*oce + g_reg*reg from the stored d loss/d(input)."""
    if g_oce is None and g_reg is None:
        if g_loss is None:
            return torch.zeros_like(grad)
        # common case (loss.backward()): one scale kernel that exits at once when g_loss == 1.
        # The stored gradient is left untouched so that retain_graph=True keeps working.
        scale = g_loss.detach().to(torch.float32).reshape(1).contiguous()
        return K.scale_inplace(grad.clone(), scale)
    # rare: somebody differentiates oce_loss / regularization_loss on their own
    g_oce_only = recompute(0.0)  # regulariser off
    g_reg_only = grad - g_oce_only
    zero = torch.zeros((), device=grad.device)
    gl = zero if g_loss is None else g_loss
    go = zero if g_oce is None else g_oce
    gr = zero if g_reg is None else g_reg
    return (gl + go) * g_oce_only + (gl + gr) * g_reg_only


class _FusedLoss(torch.autograd.Function):
    """cellulus/train.py:169-176 in one kernel (cb200_oce_loss_fwd_bwd)."""

    @staticmethod
    def forward(ctx, offsets, anchor_coordinates, reference_coordinates, temperature, regularization_weight):
        ctx.set_materialize_grads(False)
        need = ctx.needs_input_grad[0]
        out, grad = K.oce_loss_fwd_bwd(offsets, anchor_coordinates, reference_coordinates, temperature,
                                       regularization_weight, want_grad=need)
        ctx.args = (temperature, regularization_weight)
        ctx.in_dtype = offsets.dtype
        ctx.save_for_backward(grad, offsets, anchor_coordinates, reference_coordinates)
        ctx.mark_non_differentiable(out)
        return out[0], out[1], out[2], out

    @staticmethod
    def backward(ctx, g_loss, g_oce, g_reg, _g_out):
        grad, offsets, anchors, refs = ctx.saved_tensors
        T, w = ctx.args
        if grad is None:
            return None, None, None, None, None
        if g_oce is None and g_reg is None and g_loss is not None:
            # d loss / d offsets was produced by the forward pass; apply the upstream scalar on the
            # device (the kernel returns immediately when it is exactly 1, as for loss.backward()).
            if getattr(ctx, "consumed", False):  # second backward through a retained graph
                grad = K.oce_loss_fwd_bwd(offsets, anchors, refs, T, w, True)[1]
            ctx.consumed = True
            total = K.scale_inplace(grad, g_loss.detach().to(torch.float32).reshape(1).contiguous())
        else:
            total = _combine(grad, g_loss, g_oce, g_reg,
                             lambda ww: K.oce_loss_fwd_bwd(offsets, anchors, refs, T, ww, True)[1])
        return total.to(ctx.in_dtype), None, None, None, None


def oce_loss_fused(offsets, anchor_coordinates, reference_coordinates, temperature, regularizer_weight,
                   return_raw: bool = False):
    """Fused replacement of

        ea = model.select_and_add_coordinates(offsets, anchor_coordinates)
        er = model.select_and_add_coordinates(offsets, reference_coordinates)
        loss, oce_loss, regularization_loss = criterion(ea, er)

    (`cellulus/train.py:169-176`).  `offsets` (B, D, *S) fp32/bf16 on a CUDA device,
    coordinates (B, P, D) int64/int32/int16 with columns (x, y[, z]).  Returns three
    zero-dim fp32 tensors; `loss` (and the other two) are differentiable w.r.t. `offsets`.
    With `return_raw=True` a 4th value is returned: the raw 4-float result whose last entry
    counts pairs skipped because a coordinate was out of range (the reference would raise
    IndexError for those; read it with `.item()` when you want that check).
    """
    loss, oce, reg, raw = _FusedLoss.apply(offsets, anchor_coordinates, reference_coordinates,
                                           float(temperature), float(regularizer_weight))
    if return_raw:
        return loss, oce, reg, raw
    return loss, oce, reg


class _SampledLoss(torch.autograd.Function):
    """The loss slice on the device pair stream (cb200_oce_loss_sampled): sampling + gather + loss + backward."""

    @staticmethod
    def forward(ctx, offsets, kappa, num_anchors, num_references, seed, sequence, temperature,
                regularization_weight, extent_xyz):
        ctx.set_materialize_grads(False)
        need = ctx.needs_input_grad[0]
        out, grad, _ = K.oce_loss_sampled(offsets, kappa, num_anchors, num_references, seed, sequence, temperature,
                                          regularization_weight, extent_xyz, want_grad=need)
        ctx.args = (kappa, num_anchors, num_references, seed, sequence, temperature, regularization_weight, extent_xyz)
        ctx.in_dtype = offsets.dtype
        ctx.save_for_backward(grad, offsets)
        ctx.mark_non_differentiable(out)
        return out[0], out[1], out[2], out

    @staticmethod
    def backward(ctx, g_loss, g_oce, g_reg, _g_out):
        grad, offsets = ctx.saved_tensors
        kappa, na, nr, seed, seq, T, w, ext = ctx.args
        none = (None,) * 8
        if grad is None:
            return (None,) + none

        def recompute(ww):  # the stream is a pure function of (seed, sequence): same pairs again
            return K.oce_loss_sampled(offsets, kappa, na, nr, seed, seq, T, ww, ext, want_grad=True)[1]

        if g_oce is None and g_reg is None and g_loss is not None:
            if getattr(ctx, "consumed", False):  # second backward through a retained graph
                grad = recompute(w)
            ctx.consumed = True
            total = K.scale_inplace(grad, g_loss.detach().to(torch.float32).reshape(1).contiguous())
        else:
            total = _combine(grad, g_loss, g_oce, g_reg, recompute)
        return (total.to(ctx.in_dtype),) + none


def oce_loss_fused_sampled(offsets, kappa, num_anchors, num_references, seed, sequence, temperature,
                           regularizer_weight, extent_xyz=None, return_raw: bool = False):
    """The whole loss slice of a training step INCLUDING the pair sampler, in one kernel:

        anchors, refs = dataset.sample_coordinates()              # zarr_dataset.py:198-242, per sample
        ea = model.select_and_add_coordinates(offsets, anchors)   # train.py:169-176
        er = model.select_and_add_coordinates(offsets, refs)
        loss, oce_loss, regularization_loss = criterion(ea, er)

    The pairs are those of the device pair stream `(seed, sequence)` -- the lists
    `kernels.sample_pairs(..., seed, sequence)` would write; `oce_loss_fused` fed those lists returns the same
    values.  No coordinate list ever exists in memory.  `extent_xyz` are the sampling extents in column order
    (default: the reversed spatial shape of `offsets`)."""
    loss, oce, reg, raw = _SampledLoss.apply(offsets, float(kappa), int(num_anchors), int(num_references), int(seed),
                                             int(sequence), float(temperature), float(regularizer_weight),
                                             None if extent_xyz is None else tuple(int(e) for e in extent_xyz))
    if return_raw:
        return loss, oce, reg, raw
    return loss, oce, reg


class GraphedLossStep:
    """One fused loss step (zero-fill + gather + loss + backward) captured in a CUDA graph.

    The step is two tiny launches; replaying them as a graph takes the host (python, ctypes, the
    caching allocator) out of the loop, which is what a launch-bound inner loop wants on B200.
    Buffers are static: write new data into `.offsets` / `.anchors` / `.refs` (or pass tensors that
    are already final), call `replay()`, read `.loss`, `.oce_loss`, `.regularization_loss`, `.grad`
    (`d loss / d offsets`, in the memory layout of `offsets`).
    """

    def __init__(self, offsets, anchor_coordinates, reference_coordinates, temperature, regularizer_weight,
                 sampled=None):
        """`sampled`: dict(kappa, num_anchors, num_references, seed[, sequence, extent_xyz]) captures the
        sampled-loss kernel instead (the coordinate arguments are then None; the stream is fixed per graph)."""
        self.offsets = offsets.detach()
        self.anchors = anchor_coordinates
        self.refs = reference_coordinates
        self.sampled = sampled
        self.temperature = float(temperature)
        self.regularizer_weight = float(regularizer_weight)
        dev = self.offsets.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside the capture (workspace allocation, occupancy query)
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.raw, self.grad = self._run()
        self.loss, self.oce_loss, self.regularization_loss = self.raw[0], self.raw[1], self.raw[2]

    def _run(self):
        if self.sampled is not None:
            sp = self.sampled
            out, grad, _ = K.oce_loss_sampled(self.offsets, sp["kappa"], sp["num_anchors"], sp["num_references"],
                                              sp["seed"], sp.get("sequence", 0), self.temperature,
                                              self.regularizer_weight, sp.get("extent_xyz"), want_grad=True)
            return out, grad
        return K.oce_loss_fwd_bwd(self.offsets, self.anchors, self.refs, self.temperature, self.regularizer_weight,
                                  want_grad=True)

    def replay(self):
        self.graph.replay()
        K.launch_counter["calls"] += 1
        return self.loss


class GraphedLossCycle:
    """Several loss steps (each with its own static buffers) captured back to back in ONE CUDA graph.

    Consecutive graph launches leave ~1.5 us of front-end gap on the stream (measured at configs[1]: 47.3 us per
    step replayed one graph per step, 45.8 us as a three-step graph); inside a graph the steps are kernel -> kernel
    edges.  `replay()` runs every step once, in order; `cycle.results[i]` = `(raw, grad)` of step i (`raw` =
    [loss, oce, reg, n_bad]).  This is how a training loop that captures its whole iteration sees the loss step:
    in the middle of a graph, not at the head of one."""

    def __init__(self, steps):
        self.steps = list(steps)
        dev = self.steps[0].offsets.device
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.results = [s._run() for s in self.steps]
        torch.cuda.current_stream(dev).synchronize()

    def __len__(self):
        return len(self.steps)

    def replay(self):
        self.graph.replay()
        K.launch_counter["calls"] += len(self.steps)


class OCELoss(nn.Module):  # type: ignore
    def __init__(
        self,
        temperature: float,
        regularization_weight: float,
        density: float,
        num_spatial_dims: int,
        device: torch.device,
    ):
        """Same constructor as the reference (`criterions/oce_loss.py:6-43`);
        `density`, `num_spatial_dims` and `device` are stored and unused there too."""
        super().__init__()
        self.temperature = temperature
        self.regularization_weight = regularization_weight
        self.density = density
        self.num_spatial_dims = num_spatial_dims
        self.device = device

    @staticmethod
    def distance_function(embedding_0, embedding_1):
        difference = embedding_0 - embedding_1
        return difference.norm(2, dim=-1)

    def non_linearity(self, distance):
        return 1 - (-distance.pow(2) / self.temperature).exp()

    def forward(self, anchor_embedding, reference_embedding):
        """(B, P, D) x 2 -> (loss, oce_loss, regularization_loss), all sums (:53-63)."""
        return _PairLoss.apply(anchor_embedding, reference_embedding, float(self.temperature),
                               float(self.regularization_weight))

    def fused_sampled(self, offsets, kappa, num_anchors, num_references, seed, sequence=0, extent_xyz=None,
                      return_raw=False):
        """`fused` with the pair sampler inside the kernel (`oce_loss_fused_sampled`)."""
        return oce_loss_fused_sampled(offsets, kappa, num_anchors, num_references, seed, sequence, self.temperature,
                                      self.regularization_weight, extent_xyz, return_raw)

    def fused(self, offsets, anchor_coordinates, reference_coordinates):
        """The whole loss slice of `train_iteration` in one kernel."""
        return oce_loss_fused(offsets, anchor_coordinates, reference_coordinates, self.temperature,
                              self.regularization_weight)
