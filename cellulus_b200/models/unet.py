"""`UNetModel` (`cellulus/models/unet.py`): backbone + 1x1 head, with the two hot-path pieces that live on
the reference's model class running on the B200 kernels:

* `select_and_add_coordinates` (`:108-124`) -- the neighbour gather, differentiable (scatter-add backward);
* the infer-mode forward (`:73-100`) -- the TTA loop keeps everything on the device: salt/pepper noise from a
  device Philox stream (`cb200_salt_pepper`) instead of a CPU `torch.rand` + H2D copy per pass, and every
  prediction is folded into a running Welford state (`cb200_tta_accumulate`) instead of `.cpu()` + stack +
  `std_mean` on the host.
"""

from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn

from cellulus_b200 import kernels as K
from cellulus_b200.models.backbone import UNet


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


def tta_aggregate(predictions: torch.Tensor) -> torch.Tensor:
    """`models/unet.py:90-98`: (T, C, *S) fp32 stack of noisy predictions -> (C+1, *S): channel means, then
    the per-channel population std summed over channels.  One streaming kernel."""
    return K.tta_aggregate(predictions)


class TTAAccumulator:
    """Streaming form: fold each prediction in as it is produced (no T-deep stack, no `.cpu()` per pass)."""

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


class UNetModel(nn.Module):  # type: ignore
    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        num_fmaps: int,
        fmap_inc_factor: int,
        features_in_last_layer: int,
        downsampling_factors: List[Tuple[int, ...]],
        num_spatial_dims: int,
    ):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.features_in_last_layer = features_in_last_layer
        self.mode = "train"
        kernels = [[(3,) * num_spatial_dims, (1,) * num_spatial_dims, (1,) * num_spatial_dims,
                    (3,) * num_spatial_dims]]
        self.backbone = UNet(
            in_channels=in_channels, num_fmaps=num_fmaps, fmap_inc_factor=fmap_inc_factor,
            downsample_factors=[tuple(f) for f in downsampling_factors], num_fmaps_out=features_in_last_layer,
            kernel_size_down=kernels * (len(downsampling_factors) + 1),
            kernel_size_up=kernels * len(downsampling_factors))
        conv = nn.Conv2d if num_spatial_dims == 2 else nn.Conv3d
        self.head = nn.Sequential(
            conv(features_in_last_layer, features_in_last_layer, 1), nn.ReLU(),
            conv(features_in_last_layer, out_channels, 1))
        self.tta_seed = 0
        self.tta_cuda_graph = True
        self._tta_graphs = {}

    def head_forward(self, backbone_output):
        return self.head(backbone_output)

    def forward(self, raw):
        if self.mode == "train":
            return self.head_forward(self.backbone(raw))
        # mode == "infer": `models/unet.py:73-100` with the loop resident on the device
        embeddings = []
        for sample in range(raw.shape[0]):
            raw_sample = raw[sample: sample + 1].detach().float().contiguous()
            if self.tta_cuda_graph and raw_sample.is_cuda and not torch.is_grad_enabled():
                embeddings.append(self._tta_replay(raw_sample))
            else:
                embeddings.append(self._tta_loop(raw_sample, self.tta_seed))
                self.tta_seed += 1
        return torch.stack(embeddings, dim=0)

    def _tta_loop(self, raw_sample, seed):
        """The 2 x `num_infer_iterations` noisy passes of one sample (`models/unet.py:76-89`) folded into the
        running mean / M2 as they are produced; `seed`: int, or a device tensor read by the noise kernel."""
        acc, pass_index = None, 0
        for val in [0.5, 1.0]:
            for _ in range(self.num_infer_iterations):
                noisy = K.salt_pepper(raw_sample, self.p_salt_pepper, val, seed, pass_index)
                prediction = self.head_forward(self.backbone(noisy))[0].detach().float().contiguous()
                if acc is None:
                    acc = TTAAccumulator(prediction.shape[0], prediction.shape[1:], prediction.device)
                acc.add(prediction)
                pass_index += 1
        return acc.result()

    def _tta_replay(self, raw_sample):
        """The whole loop of one sample as ONE CUDA graph per input shape: ~30 small kernels per pass
        x 32 passes are launch-bound when driven from the interpreter.  The noise seed lives in device memory and is
        set before every replay, so each call draws fresh noise -- the same streams, in the same order, as the
        eager loop.  Weights are read through their storage: loading a checkpoint in place is seen by the graph."""
        key = (tuple(raw_sample.shape), raw_sample.device.index, float(self.p_salt_pepper), int(self.num_infer_iterations))
        entry = self._tta_graphs.get(key)
        if entry is None:
            device = raw_sample.device
            static_in = raw_sample.clone()
            seed = torch.tensor([self.tta_seed], dtype=torch.int64, device=device)
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):  # warm-up outside the capture (cuDNN plans, allocator); noise seed untouched
                self._tta_loop(static_in, seed)
            torch.cuda.current_stream(device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._tta_loop(static_in, seed)
            entry = (graph, static_in, static_out, seed)
            self._tta_graphs[key] = entry
        graph, static_in, static_out, seed = entry
        static_in.copy_(raw_sample)
        seed.fill_(self.tta_seed)
        graph.replay()
        self.tta_seed += 1
        return static_out.clone()

    def set_infer(self, p_salt_pepper, num_infer_iterations, device, cuda_graph: bool = True):
        """`models/unet.py:102-106`; `cuda_graph=False` drives the test-time-augmentation loop eagerly."""
        self.mode = "infer"
        self.p_salt_pepper = p_salt_pepper
        self.num_infer_iterations = num_infer_iterations
        self.device: torch.device = device
        self.tta_cuda_graph = bool(cuda_graph)
        self._tta_graphs = {}

    @staticmethod
    def select_and_add_coordinates(outputs, coordinates):
        """`models/unet.py:108-124`: outputs (B,C,H,W)/(B,C,D,H,W), coordinates (B,P,D) integer with columns
        (x, y[, z]) -> (B,P,C) fp32 = gathered offset + coordinate; differentiable w.r.t. `outputs`."""
        return _GatherAddCoords.apply(outputs, coordinates)
