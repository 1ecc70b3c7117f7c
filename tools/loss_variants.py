"""Time the fused loss kernel variants on config #2 (planar / channels-last offsets, int64/int32/int16 lists)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cellulus_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
offsets = torch.randn(bench.B, bench.D, *bench.OUT, device=dev)
anchors, refs = K.sample_pairs(bench.B, (bench.OUT[1], bench.OUT[0]), bench.KAPPA, bench.N_ANCHORS, bench.N_REFS,
                               seed=1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10, inner=20, do_flush=False):
    """Back-to-back launches between one event pair (the GPU never waits for the host); us per call."""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()  # also keeps the queue non-empty while the host enqueues
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(1 if do_flush else inner):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (1 if do_flush else inner))
    return float(np.median(ts)) * 1e3, float(np.min(ts)) * 1e3


from cellulus_b200 import _cabi  # noqa: E402

ref_out, ref_grad = K.oce_loss_fwd_bwd(offsets, anchors, refs, bench.TEMP, bench.REGW)
for tma in [1, 0]:
  _cabi.load().cb200_oce_loss_set_variant(tma)
  print("TMA pipeline for int64 lists:", bool(tma))
  for layout in ["planar", "channels_last"]:
      off = offsets if layout == "planar" else offsets.contiguous(memory_format=torch.channels_last)
      for cdt in [torch.int64, torch.int32, torch.int16]:
          a, r = anchors.to(cdt), refs.to(cdt)
          for bwd in [True, False]:
              med, mn = timeit(lambda: K.oce_loss_fwd_bwd(off, a, r, bench.TEMP, bench.REGW, want_grad=bwd))
              medf, _ = timeit(lambda: K.oce_loss_fwd_bwd(off, a, r, bench.TEMP, bench.REGW, want_grad=bwd), do_flush=True)
              nbytes = bench.B * bench.P * 2 * 2 * a.element_size() + bench.N_PX * 2 * 4 * (2 if bwd else 1)
              out, grad = K.oce_loss_fwd_bwd(off, a, r, bench.TEMP, bench.REGW, want_grad=bwd)
              err = abs(out[0].item() - ref_out[0].item()) / abs(ref_out[0].item())
              gerr = (grad - ref_grad).abs().max().item() / ref_grad.abs().max().item() if bwd else 0.0
              print(f"{layout:13s} {str(cdt):12s} bwd={bwd!s:5s} median {med:7.1f} us  min {mn:7.1f} us  single+L2-flush {medf:7.1f} us  "
                    f"{nbytes / med / 1e3:7.0f} GB/s  loss-relerr {err:.1e} grad-relerr {gerr:.1e}")
