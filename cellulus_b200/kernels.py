"""Thin torch-tensor wrappers over the C ABI (device pointers + current stream).

torch is used for device memory and streams only; every computation below is
a call into `libcellulus_b200.so`.  Inputs must live on a CUDA device: there
is no CPU fallback (a CPU tensor raises).
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import Grid, check, spatial_array

_DTYPE_CODE = {
    torch.float32: _cabi.F32,
    torch.bfloat16: _cabi.BF16,
    torch.float64: _cabi.F64,
    torch.int64: _cabi.I64,
    torch.int32: _cabi.I32,
    torch.int16: _cabi.I16,
    torch.uint8: _cabi.U8,
    torch.uint16: _cabi.U16,
}

# count of kernel-launching C-ABI calls made by this process (bench.py reports it)
launch_counter = {"calls": 0}


def _lib():
    return _cabi.load()


def _stream(t: torch.Tensor) -> C.c_void_p:
    # the C ABI launches on the calling thread's current device, like any CUDA runtime call
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(
            f"tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
            "wrap the call in `with torch.cuda.device(tensor.device):` (one process per GPU sets it once)")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(
                "cellulus_b200 kernels run on a CUDA device only (no CPU fallback); got "
                f"{type(t).__name__} on {getattr(t, 'device', None)}"
            )


def _code(t: torch.Tensor, allowed) -> int:
    if t.dtype not in allowed:
        raise TypeError(f"unsupported dtype {t.dtype}; expected one of {sorted(str(a) for a in allowed)}")
    return _DTYPE_CODE[t.dtype]


_workspaces = {}


def _zero_workspace(kind: str, nbytes: int, device: torch.device) -> torch.Tensor:
    """Small persistent zero-initialised scratch, one per (kind, device, stream)."""
    key = (kind, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 64), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def fma_peak_tflops(dtype=torch.float64, device=None, iters: int = 4096, reps: int = 5) -> float:
    """Measured FMA peak (TFLOP/s) of the FP32 / FP64 pipe of `device` (`cb200_fma_peak`, CUDA events, best of reps)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.zeros(2, dtype=torch.float64, device=device)
    flop = C.c_int64(0)
    code = _DTYPE_CODE[dtype]
    best = 0.0
    with torch.cuda.device(device):
        st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(_lib().cb200_fma_peak(code, iters, 8, _ptr(out), C.byref(flop), st), "cb200_fma_peak")
            e1.record()
            e1.synchronize()
            best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


# --------------------------------------------------------------------------- loss slice
_COORD_DTYPES = (torch.int64, torch.int32, torch.int16)
_OFFSET_DTYPES = (torch.float32, torch.bfloat16)


def _check_loss_inputs(offsets, coords_list):
    _require_cuda(offsets, *coords_list)
    if offsets.ndim not in (4, 5):
        raise ValueError("offsets must be (B, C, H, W) or (B, C, D, H, W)")
    B, Cc = offsets.shape[:2]
    D = offsets.ndim - 2
    if Cc != D:
        raise ValueError(f"offsets has {Cc} channels but {D} spatial dims; the embedding is one offset per dim")
    for c in coords_list:
        if c.ndim != 3 or c.shape[0] != B or c.shape[2] != D:
            raise ValueError(f"coordinates must be (B={B}, P, {D}); got {tuple(c.shape)}")
        if c.shape != coords_list[0].shape or c.dtype != coords_list[0].dtype:
            raise ValueError("anchor and reference coordinates must have identical shape and dtype")
    return B, D


def _offsets_layout(offsets):
    """Planar (contiguous NCHW/NCDHW) or channels-last, both zero-copy; anything else is made contiguous."""
    if offsets.is_contiguous():
        return offsets, _cabi.LAYOUT_PLANAR
    cl = torch.channels_last if offsets.ndim == 4 else torch.channels_last_3d
    if offsets.is_contiguous(memory_format=cl):
        return offsets, _cabi.LAYOUT_CHANNELS_LAST
    return offsets.contiguous(), _cabi.LAYOUT_PLANAR


def oce_loss_fwd_bwd(offsets, anchors, refs, temperature, regularization_weight, want_grad=True, staged=True):
    """`cb200_oce_loss_fwd_bwd_staged`: returns `(out4, grad)`; out4 = [loss, oce, reg, n_bad] (fp32, device).
    `staged=False` calls `cb200_oce_loss_fwd_bwd` (no staging scratch: planar offsets are gathered in place)."""
    B, D = _check_loss_inputs(offsets, [anchors, refs])
    offsets, layout = _offsets_layout(offsets)
    anchors = anchors.contiguous()
    refs = refs.contiguous()
    odt = _code(offsets, _OFFSET_DTYPES)
    cdt = _code(anchors, _COORD_DTYPES)
    out = torch.empty(4, dtype=torch.float32, device=offsets.device)
    # the gradient is produced in the memory layout of `offsets` (empty_like preserves channels_last)
    grad = torch.empty_like(offsets, dtype=torch.float32) if want_grad else None
    ws = _zero_workspace("loss", _lib().cb200_oce_loss_workspace_bytes(), offsets.device)
    spatial = spatial_array(offsets.shape[2:])
    # planar 2-D offsets (the reference's NCHW output): scratch for the kernel's own channels-last copy
    staging_bytes = _lib().cb200_oce_loss_staging_bytes(odt, layout, B, D, spatial) if staged else 0
    staging = torch.empty(staging_bytes, dtype=torch.uint8, device=offsets.device) if staging_bytes > 0 else None
    rc = _lib().cb200_oce_loss_fwd_bwd_staged(
        _ptr(offsets), odt, layout, _ptr(anchors), _ptr(refs), cdt, B, D, spatial,
        anchors.shape[1], float(temperature), float(regularization_weight), _ptr(grad), _ptr(out), _ptr(ws),
        _ptr(staging), staging_bytes, _stream(offsets))
    check(rc, "cb200_oce_loss_fwd_bwd_staged")
    launch_counter["calls"] += 1
    return out, grad


def scale_inplace(grad: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    _require_cuda(grad, scale)
    dense = grad.is_contiguous() or (grad.ndim in (4, 5) and grad.is_contiguous(
        memory_format=torch.channels_last if grad.ndim == 4 else torch.channels_last_3d))
    assert grad.dtype == torch.float32 and dense and scale.dtype == torch.float32
    check(_lib().cb200_scale_inplace(_ptr(grad), grad.numel(), _ptr(scale), _stream(grad)), "cb200_scale_inplace")
    launch_counter["calls"] += 1
    return grad


def gather_add_coords(offsets, coords):
    B, D = _check_loss_inputs(offsets, [coords])
    offsets = offsets.contiguous()
    coords = coords.contiguous()
    out = torch.empty((B, coords.shape[1], D), dtype=torch.float32, device=offsets.device)
    rc = _lib().cb200_gather_add_coords(
        _ptr(offsets), _code(offsets, _OFFSET_DTYPES), _ptr(coords), _code(coords, _COORD_DTYPES), B, D,
        spatial_array(offsets.shape[2:]), coords.shape[1], _ptr(out), _stream(offsets))
    check(rc, "cb200_gather_add_coords")
    launch_counter["calls"] += 1
    return out


def scatter_add_coords(grad_out, coords, offsets_shape):
    _require_cuda(grad_out, coords)
    grad_out = grad_out.contiguous().float()
    coords = coords.contiguous()
    B, D = offsets_shape[0], len(offsets_shape) - 2
    grad = torch.empty(tuple(offsets_shape), dtype=torch.float32, device=grad_out.device)
    rc = _lib().cb200_scatter_add_coords(
        _ptr(grad_out), _ptr(coords), _code(coords, _COORD_DTYPES), B, D, spatial_array(offsets_shape[2:]),
        coords.shape[1], _ptr(grad), _stream(grad_out))
    check(rc, "cb200_scatter_add_coords")
    launch_counter["calls"] += 1
    return grad


def oce_pair_loss(ea, er, temperature, regularization_weight, want_grad=True):
    _require_cuda(ea, er)
    if ea.shape != er.shape or ea.shape[-1] not in (2, 3):
        raise ValueError("embeddings must both be (..., D) with D in {2, 3}")
    ea = ea.contiguous().float()
    er = er.contiguous().float()
    n = ea.numel() // ea.shape[-1]
    out = torch.empty(4, dtype=torch.float32, device=ea.device)
    grad = torch.empty_like(ea) if want_grad else None
    ws = _zero_workspace("loss", _lib().cb200_oce_loss_workspace_bytes(), ea.device)
    rc = _lib().cb200_oce_pair_loss(_ptr(ea), _ptr(er), n, ea.shape[-1], float(temperature),
                                    float(regularization_weight), _ptr(grad), _ptr(out), _ptr(ws), _stream(ea))
    check(rc, "cb200_oce_pair_loss")
    launch_counter["calls"] += 1
    return out, grad


def sample_pairs(batch, extent_xyz, kappa, num_anchors, num_references, seed, sequence=0,
                 dtype=torch.int64, device="cuda"):
    """`cb200_sample_pairs`: (anchors, refs), each (B, num_anchors*num_references, D)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("cellulus_b200 kernels run on a CUDA device only (no CPU fallback)")
    D = len(extent_xyz)
    P = int(num_anchors) * int(num_references)
    anchors = torch.empty((batch, P, D), dtype=dtype, device=device)
    refs = torch.empty_like(anchors)
    with torch.cuda.device(device):
        rc = _lib().cb200_sample_pairs(
            _ptr(anchors), _ptr(refs), _code(anchors, _COORD_DTYPES), batch, D, spatial_array(extent_xyz),
            float(kappa), int(num_anchors), int(num_references), int(seed) & (2**64 - 1), int(sequence),
            _stream(anchors))
    check(rc, "cb200_sample_pairs")
    launch_counter["calls"] += 1
    return anchors, refs


def oce_loss_sampled(offsets, kappa, num_anchors, num_references, seed, sequence, temperature,
                     regularization_weight, extent_xyz=None, want_grad=True, dump_dtype=None, staged=True):
    """`cb200_oce_loss_sampled_staged`: the loss on the pair stream (seed, sequence), pairs drawn inside the kernel.
    Returns `(out4, grad, (anchors, refs) | None)`; the lists are only written when `dump_dtype` is given.
    `staged=False`: planar offsets are gathered in place (no staging scratch)."""
    _require_cuda(offsets)
    if offsets.ndim not in (4, 5) or offsets.shape[1] != offsets.ndim - 2:
        raise ValueError("offsets must be (B, D, *S) with one offset channel per spatial dim")
    B, D = offsets.shape[0], offsets.ndim - 2
    offsets, layout = _offsets_layout(offsets)
    spatial = tuple(offsets.shape[2:])
    if extent_xyz is None:
        extent_xyz = spatial[::-1]
    if len(extent_xyz) != D:
        raise ValueError(f"extent_xyz must have {D} entries (x, y[, z])")
    odt = _code(offsets, _OFFSET_DTYPES)
    out = torch.empty(4, dtype=torch.float32, device=offsets.device)
    grad = torch.empty_like(offsets, dtype=torch.float32) if want_grad else None
    ws = _zero_workspace("loss", _lib().cb200_oce_loss_workspace_bytes(), offsets.device)
    lists, ddt = None, 0
    if dump_dtype is not None:
        P = int(num_anchors) * int(num_references)
        lists = (torch.empty((B, P, D), dtype=dump_dtype, device=offsets.device),
                 torch.empty((B, P, D), dtype=dump_dtype, device=offsets.device))
        ddt = _code(lists[0], _COORD_DTYPES)
    staging_bytes = _lib().cb200_oce_loss_staging_bytes(odt, layout, B, D, spatial_array(spatial)) if staged else 0
    staging = torch.empty(staging_bytes, dtype=torch.uint8, device=offsets.device) if staging_bytes > 0 else None
    rc = _lib().cb200_oce_loss_sampled_staged(
        _ptr(offsets), odt, layout, B, D, spatial_array(spatial), spatial_array(extent_xyz), float(kappa),
        int(num_anchors), int(num_references), int(seed) & (2**64 - 1), int(sequence), float(temperature),
        float(regularization_weight), _ptr(grad), _ptr(out), _ptr(ws), _ptr(lists[0] if lists else None),
        _ptr(lists[1] if lists else None), ddt, _ptr(staging), staging_bytes, _stream(offsets))
    check(rc, "cb200_oce_loss_sampled_staged")
    launch_counter["calls"] += 1
    return out, grad, lists


# --------------------------------------------------------------------------- TTA
def tta_aggregate(stack: torch.Tensor) -> torch.Tensor:
    """(T, C, *S) fp32 -> (C+1, *S) fp32 (`models/unet.py:90-98`)."""
    _require_cuda(stack)
    if stack.dtype != torch.float32 or stack.ndim < 3:
        raise TypeError("stack must be fp32 (T, C, *S)")
    stack = stack.contiguous()
    T, Cc = stack.shape[:2]
    spatial = stack.shape[2:]
    n = int(np.prod(spatial))
    out = torch.empty((Cc + 1, *spatial), dtype=torch.float32, device=stack.device)
    check(_lib().cb200_tta_aggregate(_ptr(stack), T, Cc, n, _ptr(out), _stream(stack)), "cb200_tta_aggregate")
    launch_counter["calls"] += 1
    return out


def tta_accumulate(state, prediction, t):
    _require_cuda(state, prediction)
    prediction = prediction.contiguous()
    Cc = prediction.shape[0]
    n = prediction.numel() // Cc
    check(_lib().cb200_tta_accumulate(_ptr(state), _ptr(prediction), int(t), Cc, n, _stream(state)),
          "cb200_tta_accumulate")
    launch_counter["calls"] += 1


def tta_finalize(state, num_passes, channels, spatial):
    n = int(np.prod(spatial))
    out = torch.empty((channels + 1, *spatial), dtype=torch.float32, device=state.device)
    check(_lib().cb200_tta_finalize(_ptr(state), int(num_passes), channels, n, _ptr(out), _stream(state)),
          "cb200_tta_finalize")
    launch_counter["calls"] += 1
    return out


_FLOAT_DTYPES = (torch.float32, torch.float64)


def salt_pepper(raw: torch.Tensor, p: float, value: float, seed, sequence: int) -> torch.Tensor:
    """`noisy[rnd <= p] = value` of `models/unet.py:80-82` with the uniform draw on the device.
    `seed`: an int, or a one-element int64 CUDA tensor that the kernel reads when it runs (CUDA-graph replays)."""
    _require_cuda(raw)
    raw = raw.contiguous()
    assert raw.dtype == torch.float32
    out = torch.empty_like(raw)
    if isinstance(seed, torch.Tensor):
        _require_cuda(seed)
        assert seed.dtype == torch.int64 and seed.numel() == 1 and seed.device == raw.device
        check(_lib().cb200_salt_pepper_device_seed(_ptr(raw), raw.numel(), float(p), float(value), _ptr(seed),
                                                   int(sequence), _ptr(out), _stream(raw)), "cb200_salt_pepper_device_seed")
    else:
        check(_lib().cb200_salt_pepper(_ptr(raw), raw.numel(), float(p), float(value), int(seed) & (2**64 - 1),
                                       int(sequence), _ptr(out), _stream(raw)), "cb200_salt_pepper")
    launch_counter["calls"] += 1
    return out


# --------------------------------------------------------------------------- detect preamble
def centre_embeddings(emb: torch.Tensor, threshold: float, want_centred: bool = True):
    """`detect.py:97-119`.  Returns `(means (D,) float64 device, centred copy of emb or None)`."""
    _require_cuda(emb)
    emb = emb.contiguous()
    D = emb.shape[0] - 1
    n = emb[0].numel()
    means = torch.empty(D, dtype=torch.float64, device=emb.device)
    out = torch.empty_like(emb) if want_centred else None
    ws = _zero_workspace("centre", _lib().cb200_centre_workspace_bytes(), emb.device)
    check(_lib().cb200_centre_embeddings(_ptr(emb), _code(emb, _FLOAT_DTYPES), D, n, float(threshold), _ptr(means),
                                         _ptr(out), _ptr(ws), _stream(emb)), "cb200_centre_embeddings")
    launch_counter["calls"] += 1
    return means, out





def gaussian_weights(sigma: float, truncate: float = 4.0):
    """scipy's `_gaussian_kernel1d` (order 0): `(w[0..radius] with w[0] the centre, radius)`."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x**2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:]), radius


def find_seeds(centred: torch.Tensor, sigma: float = 2.0) -> np.ndarray:
    """`detect.py:129-132` on the device: (n_seeds, D) int64 numpy array in (x, y[, z]) order, sorted by
    descending peak intensity (stable), i.e. `np.flip(peak_local_max(-gaussian_filter(norm(...), 2)), 1)`."""
    _require_cuda(centred)
    centred = centred.contiguous()
    D = centred.shape[0] - 1
    spatial = tuple(centred.shape[1:])
    n = int(np.prod(spatial))
    dev = centred.device
    st = _stream(centred)
    mag = torch.empty(spatial, dtype=torch.float64, device=dev)
    check(_lib().cb200_channel_norm(_ptr(centred), _code(centred, _FLOAT_DTYPES), D, n, _ptr(mag), st),
          "cb200_channel_norm")
    w, radius = gaussian_weights(sigma)
    smooth, scratch = torch.empty_like(mag), torch.empty_like(mag)
    check(_lib().cb200_gaussian_blur(_ptr(mag), _ptr(smooth), _ptr(scratch), D, spatial_array(spatial),
                                     w.ctypes.data_as(C.POINTER(C.c_double)), radius, 1, st), "cb200_gaussian_blur")
    lo = float(minmax(smooth)[0].item())  # threshold_abs = image.min()
    cap = max(1024, n // 27 + 16)
    ws = torch.empty(_lib().cb200_peaks_workspace_bytes(n), dtype=torch.uint8, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    while True:
        idx = torch.empty(cap, dtype=torch.int32, device=dev)
        val = torch.empty(cap, dtype=torch.float64, device=dev)
        check(_lib().cb200_local_peaks(_ptr(smooth), D, spatial_array(spatial), lo, _ptr(idx), _ptr(val), cap,
                                       _ptr(n_out), _ptr(ws), st), "cb200_local_peaks")
        k = int(n_out.item())
        if k <= cap:
            break
        cap = k
    launch_counter["calls"] += 4
    idx = idx[:k].cpu().numpy().astype(np.int64)
    val = val[:k].cpu().numpy()
    order = np.argsort(-val, kind="stable")
    coords = np.stack(np.unravel_index(idx[order], spatial), axis=1) if k else np.zeros((0, D), np.int64)
    return np.flip(coords, 1).astype(np.int64)


def minmax(x: torch.Tensor) -> torch.Tensor:
    """{min, max} of x as a 2-element float64 device tensor."""
    _require_cuda(x)
    x = x.contiguous()
    out = torch.empty(2, dtype=torch.float64, device=x.device)
    ws = _zero_workspace("reduce", _lib().cb200_reduce_workspace_bytes(), x.device)
    check(_lib().cb200_minmax(_ptr(x), _code(x, _FLOAT_DTYPES), x.numel(), _ptr(out), _ptr(ws), _stream(x)),
          "cb200_minmax")
    launch_counter["calls"] += 1
    return out


def histogram(x: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """np.histogram(x, len(edges)-1, range=(edges[0], edges[-1])) counts, uint64-exact (returned as int64)."""
    _require_cuda(x, edges)
    x = x.contiguous()
    assert edges.dtype == torch.float64 and edges.is_contiguous()
    nbins = edges.numel() - 1
    counts = torch.zeros(nbins, dtype=torch.int64, device=x.device)
    check(_lib().cb200_histogram(_ptr(x), _code(x, _FLOAT_DTYPES), x.numel(), _ptr(edges), nbins, _ptr(counts),
                                 _stream(x)), "cb200_histogram")
    launch_counter["calls"] += 1
    return counts


def fg_compact(emb: torch.Tensor, threshold: float, capacity: Optional[int] = None, mask_dtype=None):
    """Foreground compaction.  Returns `(points (D, capacity) f64 SoA, pix_index (capacity) i32,
    n_fg (python int), mask or None)`.  One host sync (the count)."""
    _require_cuda(emb)
    emb = emb.contiguous()
    D = emb.shape[0] - 1
    spatial = emb.shape[1:]
    if len(spatial) != D or D not in (2, 3):
        raise ValueError("emb must be (D+1, *S) with D spatial dims, D in {2, 3}")
    n_pix = int(np.prod(spatial))
    dev = emb.device
    ws = torch.empty(_lib().cb200_compact_workspace_bytes(n_pix), dtype=torch.uint8, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    mask = None
    mcode = 0
    if mask_dtype is not None:
        mask = torch.empty(spatial, dtype=mask_dtype, device=dev)
        mcode = _code(mask, (torch.uint8, torch.uint16))

    def run(cap):
        cap_even = max(2, (cap + 1) & ~1)  # even stride: the brute-force kernel moves 16-byte granules
        pts = torch.empty((D, cap_even), dtype=torch.float64, device=dev)
        pix = torch.empty(cap_even, dtype=torch.int32, device=dev)
        rc = _lib().cb200_fg_compact(_ptr(emb), _code(emb, _FLOAT_DTYPES), D, spatial_array(spatial),
                                     float(threshold), _ptr(pts), _ptr(pix), cap_even, _ptr(n_out), _ptr(mask), mcode,
                                     _ptr(ws), _stream(emb))
        check(rc, "cb200_fg_compact")
        launch_counter["calls"] += 1
        return pts, pix, int(n_out.item())

    cap = n_pix if capacity is None else int(capacity)
    pts, pix, n = run(cap)
    if n > pts.shape[1]:
        pts, pix, n = run(n)
    return pts, pix, n, mask


def select_points(points: torch.Tensor, n: int, flags: torch.Tensor):
    """Rows of an SoA point set where flags != 0 (order kept).  Returns `(subset (D, cap) SoA, count)`."""
    _require_cuda(points, flags)
    D = points.shape[0]
    assert flags.dtype == torch.uint8 and flags.numel() >= n
    dev = points.device
    cap = max(2, (n + 1) & ~1)
    dst = torch.empty((D, cap), dtype=torch.float64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(_lib().cb200_compact_workspace_bytes(max(n, 1)), dtype=torch.uint8, device=dev)
    rc = _lib().cb200_select_points(_ptr(points), n, points.stride(0), D, _ptr(flags), _ptr(dst), cap, _ptr(n_out),
                                    _ptr(ws), _stream(points))
    check(rc, "cb200_select_points")
    launch_counter["calls"] += 1
    return dst, int(n_out.item())


def bernoulli_flags(n: int, p: float, seed: int, device) -> torch.Tensor:
    flags = torch.empty(max(n, 1), dtype=torch.uint8, device=device)
    check(_lib().cb200_bernoulli_flags(_ptr(flags), n, float(p), int(seed) & (2**64 - 1), _stream(flags)),
          "cb200_bernoulli_flags")
    launch_counter["calls"] += 1
    return flags


# --------------------------------------------------------------------------- mean-shift
def plan_grid(lo, hi, bandwidth: float, max_cells: int = 1 << 26) -> Grid:
    D = len(lo)
    g = Grid()
    lo_a = (C.c_double * D)(*[float(v) for v in lo])
    hi_a = (C.c_double * D)(*[float(v) for v in hi])
    check(_lib().cb200_grid_plan(lo_a, hi_a, D, float(bandwidth), int(max_cells), C.byref(g)), "cb200_grid_plan")
    return g


def bounding_box(points: torch.Tensor, n: int):
    """Per-column (min, max) of an SoA point set; one host sync."""
    D = points.shape[0]
    mm = torch.stack([minmax(points[k, :n]) for k in range(D)])  # (D, 2)
    mm = mm.cpu().numpy()
    return mm[:, 0], mm[:, 1]


def grid_build(points: torch.Tensor, n: int, grid: Grid, want_order: bool = False):
    """Sort an SoA point set by grid cell.  Returns `(sorted SoA (D, cap), cell_start (n_cells+1) i32, order|None)`."""
    _require_cuda(points)
    D = points.shape[0]
    dev = points.device
    cap = max(2, (n + 1) & ~1)
    sorted_pts = torch.empty((D, cap), dtype=torch.float64, device=dev)
    cell_start = torch.empty(grid.n_cells + 1, dtype=torch.int32, device=dev)
    order = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if want_order else None
    nbytes = _lib().cb200_grid_build_workspace_bytes(n, grid.n_cells)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = _lib().cb200_grid_build(_ptr(points), n, points.stride(0), C.byref(grid), _ptr(sorted_pts), cap, _ptr(order),
                                 _ptr(cell_start), _ptr(ws), nbytes, _stream(points))
    check(rc, "cb200_grid_build")
    launch_counter["calls"] += 1
    return sorted_pts, cell_start, order


# the work buffer of the most recent `ms_grid_modes` call; `grid_modes_distance_tests()` reads its statistic
last_grid_modes_work = [None]


def grid_modes_distance_tests() -> int:
    """Distance tests (seed x candidate point evaluations) made by the most recent `ms_grid_modes` call."""
    w = last_grid_modes_work[0]
    return 0 if w is None else int(w[2:4].view(torch.int64).item())


def grid_modes_climb_steps() -> int:
    """Sum over seeds of (iterations + 1) of the most recent `ms_grid_modes` call."""
    w = last_grid_modes_work[0]
    return 0 if w is None else int(w[4:6].view(torch.int64).item())


def ms_grid_modes(sorted_pts, n, grid, cell_start, seeds_soa, n_seeds, bandwidth, max_iter=300):
    """Climb every seed to convergence (grid-hash form).  `seeds_soa` (D, cap) is updated IN PLACE to the modes.
    Returns `(counts, iters)` int32 device tensors."""
    dev = sorted_pts.device
    counts = torch.zeros(max(n_seeds, 1), dtype=torch.int32, device=dev)
    iters = torch.zeros(max(n_seeds, 1), dtype=torch.int32, device=dev)
    work = torch.zeros(8, dtype=torch.int32, device=dev)  # [0] claim counter, [2:4] / [4:6] uint64 statistics
    rc = _lib().cb200_ms_grid_modes(_ptr(sorted_pts), n, sorted_pts.stride(0), C.byref(grid), _ptr(cell_start),
                                    _ptr(seeds_soa), seeds_soa.stride(0), n_seeds, float(bandwidth), int(max_iter),
                                    _ptr(counts), _ptr(iters), _ptr(work), _stream(sorted_pts))
    check(rc, "cb200_ms_grid_modes")
    launch_counter["calls"] += 1
    last_grid_modes_work[0] = work
    return counts, iters


def default_merge_rounds(n_seeds: int) -> int:
    """How often `ms_grid_modes_distinct` merges identical trajectories: once for ordinary seed counts (every further
    round costs two small launches and a merge pass), more often for millions of dense seeds (CB200_MS_ROUNDS overrides)."""
    env = os.environ.get("CB200_MS_ROUNDS")
    if env:
        return max(1, min(30, int(env)))
    return 1 if n_seeds < 1_000_000 else 4


def ms_grid_modes_distinct(sorted_pts, n, grid, cell_start, seeds_soa, n_seeds, bandwidth, max_iter=300, merge_rounds=None):
    """`cb200_ms_grid_modes_distinct`: one window evaluation for every seed, then only one representative of every
    distinct unfinished mean climbs on.  `seeds_soa` is updated in place; returns `(counts, iters)` where the merged
    copies have count 0 and a negative `iters` (they would end as copies of their representative's mode)."""
    dev = sorted_pts.device
    counts = torch.zeros(max(n_seeds, 1), dtype=torch.int32, device=dev)
    iters = torch.zeros(max(n_seeds, 1), dtype=torch.int32, device=dev)
    work = torch.zeros(16, dtype=torch.int32, device=dev)
    nbytes = _lib().cb200_ms_distinct_workspace_bytes(n_seeds)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = _lib().cb200_ms_grid_modes_distinct(_ptr(sorted_pts), n, sorted_pts.stride(0), C.byref(grid), _ptr(cell_start),
                                             _ptr(seeds_soa), seeds_soa.stride(0), n_seeds, float(bandwidth),
                                             int(max_iter), int(merge_rounds or default_merge_rounds(n_seeds)),
                                             _ptr(counts), _ptr(iters), _ptr(work), _ptr(ws), nbytes,
                                             _stream(sorted_pts))
    check(rc, "cb200_ms_grid_modes_distinct")
    launch_counter["calls"] += 1
    w = work.view(torch.int64)  # [1] tests, [2] steps of pass 1; [5], [6] of pass 2
    stats = torch.stack([w[1] + w[5], w[2] + w[6]])
    last_grid_modes_work[0] = torch.cat([work[:2], stats.view(torch.int32)])
    return counts, iters


def ms_brute_modes(points, n, seeds_soa, n_seeds, bandwidth, max_iter=300):
    """Climb every seed to convergence (brute-force n-body form); one accumulate + one update launch per
    iteration over the still-active seeds.  `seeds_soa` is updated in place.  Returns `(counts, iters)`."""
    dev = points.device
    D = points.shape[0]
    counts = torch.zeros(max(n_seeds, 1), dtype=torch.int32, device=dev)
    iters = torch.zeros(max(n_seeds, 1), dtype=torch.int32, device=dev)
    if n_seeds == 0 or n == 0:
        return counts, iters
    active = torch.arange(n_seeds, dtype=torch.int32, device=dev)
    nxt = torch.empty_like(active)
    n_next = torch.zeros(1, dtype=torch.int32, device=dev)
    n_active = n_seeds
    partial = torch.empty(_lib().cb200_ms_brute_partial_bytes(n_active, n, D), dtype=torch.uint8, device=dev)
    st = _stream(points)
    for _ in range(max_iter + 2):
        if n_active == 0:
            break
        rc = _lib().cb200_ms_brute_accumulate(_ptr(points), n, points.stride(0), D, _ptr(seeds_soa),
                                              seeds_soa.stride(0), _ptr(active), n_active, float(bandwidth),
                                              _ptr(partial), st)
        check(rc, "cb200_ms_brute_accumulate")
        n_next.zero_()
        rc = _lib().cb200_ms_update(_ptr(seeds_soa), seeds_soa.stride(0), D, _ptr(counts), _ptr(iters), _ptr(active),
                                    n_active, n, _ptr(partial), float(bandwidth), int(max_iter), _ptr(nxt),
                                    _ptr(n_next), st)
        check(rc, "cb200_ms_update")
        launch_counter["calls"] += 2
        n_active = int(n_next.item())
        active, nxt = nxt, active
    return counts, iters


def bin_seeds(points: torch.Tensor, n: int, bin_size: float):
    """`cb200_bin_seeds`: sklearn `get_bin_seeds(points, bin_size)` on the device.  Returns `(seeds SoA (D, cap)
    float64, n_seeds)`; one host sync (the count)."""
    _require_cuda(points)
    D = points.shape[0]
    dev = points.device
    cap = max(2, (n + 1) & ~1)
    seeds = torch.zeros((D, cap), dtype=torch.float64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = _lib().cb200_bin_seeds_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = _lib().cb200_bin_seeds(_ptr(points), n, points.stride(0), D, float(bin_size), _ptr(seeds), cap, _ptr(n_out),
                                _ptr(overflow), _ptr(ws), nbytes, _stream(points))
    check(rc, "cb200_bin_seeds")
    launch_counter["calls"] += 1
    k, bad = int(n_out.item()), int(overflow.item())
    if bad:
        raise _cabi.CellulusB200Error("bin_seeds: a bin index exceeds 2^20 (coordinates / bin_size too large)")
    return seeds, k


def unique_modes(modes_soa, counts, n_seeds):
    """`cb200_unique_modes`: one copy of every bit-identical mode with count > 0 (the copy the suppression would keep).
    Returns `(modes (D, cap) SoA, counts, n_unique)`; one host sync for the count."""
    dev = modes_soa.device
    D = modes_soa.shape[0]
    cap = max(2, (n_seeds + 1) & ~1)
    out = torch.empty((D, cap), dtype=torch.float64, device=dev)
    counts_out = torch.empty(max(n_seeds, 1), dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    nbytes = _lib().cb200_unique_modes_workspace_bytes(n_seeds)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = _lib().cb200_unique_modes(_ptr(modes_soa), modes_soa.stride(0), D, _ptr(counts), n_seeds, _ptr(out), cap,
                                   _ptr(counts_out), _ptr(n_out), _ptr(ws), nbytes, _stream(modes_soa))
    check(rc, "cb200_unique_modes")
    launch_counter["calls"] += 1
    return out, counts_out, int(n_out.item())


def nms_centres(modes_soa, counts, n_seeds, bandwidth, grid: Grid, rounds_per_call: int = 4, dedupe: bool = True):
    """sklearn:511-547 on the device.  Returns `(centres (D, cap) SoA in `cluster_centers_` order, K)`.
    One host sync per call of `rounds_per_call` rounds (the fix-point takes 2-3 rounds in practice).
    `dedupe`: merge bit-identical modes first (`cb200_unique_modes`; same centres, in the same order, from far fewer
    candidates)."""
    dev = modes_soa.device
    D = modes_soa.shape[0]
    if dedupe and n_seeds > 1024:
        modes_soa, counts, n_seeds = unique_modes(modes_soa, counts, n_seeds)
        if n_seeds == 0:
            return torch.zeros((D, 2), dtype=torch.float64, device=dev), 0
    out2 = torch.zeros(2, dtype=torch.int32, device=dev)
    nbytes = _lib().cb200_nms_workspace_bytes(n_seeds, C.byref(grid), float(bandwidth))
    if nbytes < 0:
        raise _cabi.CellulusB200Error("the bounding box of the modes is too large for the suppression grid")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    resume = 0
    for _ in range(64):
        rc = _lib().cb200_nms_suppress(_ptr(modes_soa), modes_soa.stride(0), D, _ptr(counts), n_seeds,
                                       float(bandwidth), C.byref(grid), int(rounds_per_call), resume, _ptr(out2),
                                       _ptr(ws), nbytes, _stream(modes_soa))
        check(rc, "cb200_nms_suppress")
        launch_counter["calls"] += 1
        k, undecided = (int(x) for x in out2.tolist())
        if undecided == 0:
            break
        resume = 1
    else:
        raise _cabi.CellulusB200Error("centre suppression did not reach its fix-point")
    cap = max(2, (k + 1) & ~1)
    centres = torch.zeros((D, cap), dtype=torch.float64, device=dev)
    if k:
        rc = _lib().cb200_nms_emit(_ptr(modes_soa), modes_soa.stride(0), D, _ptr(counts), n_seeds, float(bandwidth),
                                   C.byref(grid), k, _ptr(centres), cap, _ptr(ws), nbytes, _stream(modes_soa))
        check(rc, "cb200_nms_emit")
        launch_counter["calls"] += 1
    return centres, k


def assign_labels(points, n, centres, k, pix_index, labels_out, grid: Optional[Grid] = None):
    """labels_out[pix_index[i]] = 1 + nearest centre (ties -> lowest index); labels_out int32 or uint16.
    With `grid` (cells of edge >= bandwidth covering the centres) the search is pruned to 3^D cells."""
    ws = None
    if grid is not None:
        nbytes = _lib().cb200_assign_workspace_bytes(n, int(k), grid.n_cells)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
    rc = _lib().cb200_assign_labels(_ptr(points), n, points.stride(0), points.shape[0], _ptr(centres),
                                    centres.stride(0), int(k), C.byref(grid) if grid is not None else None,
                                    _ptr(pix_index), _ptr(labels_out),
                                    _code(labels_out, (torch.int32, torch.uint16)), _ptr(ws), _stream(points))
    check(rc, "cb200_assign_labels")
    launch_counter["calls"] += 1
    return labels_out


def greedy_cluster(emb: torch.Tensor, fg_mask: torch.Tensor, bandwidth: float, min_object_size: int,
                   seed_thresh: float = 0.9, min_unclustered_sum: int = 0, compute_dtype=None):
    """`Cluster2d/3d.cluster` (utils/greedy_cluster.py) on the device.  emb (D+1, *S) fp32/fp64, fg_mask (*S)
    uint8/bool.  Returns `(instance_map (*S) int16, n_objects, n_seeds_tried)`."""
    _require_cuda(emb, fg_mask)
    emb = emb.contiguous()
    D = emb.shape[0] - 1
    spatial = tuple(emb.shape[1:])
    n_pix = int(np.prod(spatial))
    dev = emb.device
    if compute_dtype is None:  # the reference: `.float()` in 2-D (:84), the stored dtype in 3-D (:240)
        compute_dtype = torch.float32 if D == 2 else emb.dtype
    mask = fg_mask.to(torch.uint8).contiguous()
    seed = emb[D].to(compute_dtype)
    mm = minmax(seed.to(torch.float32) if compute_dtype == torch.float32 else seed).cpu().numpy()
    n_fg = int(mask.sum().item())
    out = torch.zeros(spatial, dtype=torch.int16, device=dev)
    if n_fg == 0:
        return out, 0, 0
    cap = max(2, (n_fg + 1) & ~1)
    emb_m = torch.empty((D, cap), dtype=compute_dtype, device=dev)
    seed_m = torch.empty(cap, dtype=compute_dtype, device=dev)
    pix = torch.empty(cap, dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(_lib().cb200_compact_workspace_bytes(n_pix), dtype=torch.uint8, device=dev)
    st = _stream(emb)
    check(_lib().cb200_greedy_prepare(_ptr(emb), _code(emb, _FLOAT_DTYPES), D, spatial_array(spatial), _ptr(mask),
                                      _DTYPE_CODE[compute_dtype], float(mm[0]), float(mm[1]), _ptr(emb_m), cap,
                                      _ptr(seed_m), _ptr(pix), _ptr(n_out), _ptr(ws), st), "cb200_greedy_prepare")
    inst = torch.empty(cap, dtype=torch.int16, device=dev)
    res = torch.zeros(2, dtype=torch.int32, device=dev)
    gws = torch.empty(_lib().cb200_greedy_workspace_bytes(n_fg), dtype=torch.uint8, device=dev)
    check(_lib().cb200_greedy_cluster(_ptr(emb_m), cap, _ptr(seed_m), n_fg, D, _DTYPE_CODE[compute_dtype],
                                      float(bandwidth), int(min_object_size), float(seed_thresh),
                                      int(min_unclustered_sum), _ptr(inst), _ptr(res), _ptr(gws), st),
          "cb200_greedy_cluster")
    check(_lib().cb200_scatter_i16(_ptr(inst), _ptr(pix), n_fg, _ptr(out), st), "cb200_scatter_i16")
    launch_counter["calls"] += 3
    n_obj, n_iter = (int(x) for x in res.tolist())
    return out, n_obj, n_iter


# --------------------------------------------------------------------------- size filter
def label_components(seg: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _require_cuda(seg)
    seg = seg.contiguous().to(torch.int32)
    labels = torch.empty_like(seg)
    n_labels = torch.zeros(1, dtype=torch.int32, device=seg.device)
    ws = torch.empty(_lib().cb200_cc_workspace_bytes(seg.numel()), dtype=torch.uint8, device=seg.device)
    rc = _lib().cb200_label_components(_ptr(seg), seg.ndim, spatial_array(seg.shape), _ptr(labels), _ptr(n_labels),
                                       _ptr(ws), _stream(seg))
    check(rc, "cb200_label_components")
    launch_counter["calls"] += 1
    return labels, n_labels


def size_filter_(seg: torch.Tensor, min_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """In-place on `seg` (int32, contiguous); returns `(relabelled, n_labels)`."""
    _require_cuda(seg)
    assert seg.dtype == torch.int32 and seg.is_contiguous()
    labels = torch.empty_like(seg)
    n_labels = torch.zeros(1, dtype=torch.int32, device=seg.device)
    ws = torch.empty(_lib().cb200_cc_workspace_bytes(seg.numel()), dtype=torch.uint8, device=seg.device)
    rc = _lib().cb200_size_filter(_ptr(seg), seg.ndim, spatial_array(seg.shape), int(min_size), _ptr(labels),
                                  _ptr(n_labels), _ptr(ws), _stream(seg))
    check(rc, "cb200_size_filter")
    launch_counter["calls"] += 1
    return labels, n_labels


# --------------------------------------------------------------------------- "cell" post-processing
def edt_within(mask: torch.Tensor, radius: float) -> torch.Tensor:
    """`cb200_edt_within`: `distance_transform_edt(mask) < radius` as a uint8 mask (2-D / 3-D)."""
    _require_cuda(mask)
    assert mask.dtype in (torch.uint8, torch.bool) and mask.is_contiguous() and mask.ndim in (2, 3)
    src = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    out = torch.empty_like(src)
    ws = torch.empty(_lib().cb200_edt_workspace_bytes(src.numel()), dtype=torch.uint8, device=src.device)
    rc = _lib().cb200_edt_within(_ptr(src), src.ndim, spatial_array(src.shape), float(radius), _ptr(out), _ptr(ws),
                                 _stream(src))
    check(rc, "cb200_edt_within")
    launch_counter["calls"] += 1
    return out


def grow_shrink_(seg: torch.Tensor, grow_distance: float, shrink_distance: float) -> torch.Tensor:
    """`cb200_grow_shrink`, in place on int32 labels: segment.py:47-50."""
    _require_cuda(seg)
    assert seg.dtype == torch.int32 and seg.is_contiguous() and seg.ndim in (2, 3)
    ws = torch.empty(_lib().cb200_edt_workspace_bytes(seg.numel()), dtype=torch.uint8, device=seg.device)
    rc = _lib().cb200_grow_shrink(_ptr(seg), seg.ndim, spatial_array(seg.shape), float(grow_distance),
                                  float(shrink_distance), _ptr(ws), _stream(seg))
    check(rc, "cb200_grow_shrink")
    launch_counter["calls"] += 1
    return seg


# --------------------------------------------------------------------------- "nucleus" post-processing
_RAW_DTYPES = (torch.float32, torch.float64, torch.uint8, torch.uint16)


def label_stats(seg: torch.Tensor, raw: torch.Tensor, max_label: int):
    """`cb200_label_stats`: per label id in [0, max_label] -> (raw_min, raw_max) float64 and box (.., 6) int32
    = lo z,y,x, hi z,y,x inclusive (hi < 0: label absent)."""
    _require_cuda(seg, raw)
    assert seg.dtype == torch.int32 and seg.is_contiguous() and raw.is_contiguous() and raw.shape == seg.shape
    dev = seg.device
    mn = torch.empty(max_label + 1, dtype=torch.float64, device=dev)
    mx = torch.empty_like(mn)
    box = torch.empty((max_label + 1, 6), dtype=torch.int32, device=dev)
    ws = torch.empty(_lib().cb200_label_stats_workspace_bytes(max_label), dtype=torch.uint8, device=dev)
    rc = _lib().cb200_label_stats(_ptr(seg), _ptr(raw), _code(raw, _RAW_DTYPES), seg.ndim, spatial_array(seg.shape),
                                  int(max_label), _ptr(mn), _ptr(mx), _ptr(box), _ptr(ws), _stream(seg))
    check(rc, "cb200_label_stats")
    launch_counter["calls"] += 1
    return mn, mx, box


def label_histogram(seg, raw, max_label, hist, raw_min=None, hist_offset=None, edges=None, nbins=0):
    """`cb200_label_histogram` into the zeroed int32 tensor `hist`."""
    _require_cuda(seg, raw, hist)
    assert hist.dtype == torch.int32 and hist.is_contiguous()
    rc = _lib().cb200_label_histogram(_ptr(seg), _ptr(raw), _code(raw, _RAW_DTYPES), seg.numel(), int(max_label),
                                      _ptr(raw_min), _ptr(hist_offset), _ptr(edges), int(nbins), _ptr(hist),
                                      _stream(seg))
    check(rc, "cb200_label_histogram")
    launch_counter["calls"] += 1
    return hist


def label_otsu(hist, hist_offset, num_bins, total_bins, arithmetic_dtype, centre0=None, centres=None):
    """`cb200_label_otsu`: float64 threshold per label (labels with num_bins <= 0 keep 0)."""
    _require_cuda(hist, hist_offset, num_bins)
    assert hist.dtype == torch.int32 and hist_offset.dtype == torch.int64 and num_bins.dtype == torch.int64
    n_labels = int(num_bins.numel())
    thresholds = torch.zeros(n_labels, dtype=torch.float64, device=hist.device)
    ws = torch.empty(_lib().cb200_label_otsu_workspace_bytes(int(total_bins)), dtype=torch.uint8, device=hist.device)
    rc = _lib().cb200_label_otsu(_ptr(hist), _ptr(hist_offset), _ptr(num_bins), _ptr(centre0), _ptr(centres),
                                 _DTYPE_CODE[arithmetic_dtype], n_labels, int(total_bins), _ptr(thresholds), _ptr(ws),
                                 _stream(hist))
    check(rc, "cb200_label_otsu")
    launch_counter["calls"] += 1
    return thresholds


def nucleus_fill(seg, raw, ids, thresholds, boxes, box_offset, total_box_voxels):
    """`cb200_nucleus_fill`: thresholded instance masks with their holes filled -> int32 label image."""
    _require_cuda(seg, raw)
    out = torch.empty_like(seg)
    n_inst = int(ids.numel())
    ws = torch.empty(_lib().cb200_nucleus_fill_workspace_bytes(int(total_box_voxels), n_inst), dtype=torch.uint8,
                     device=seg.device)
    rc = _lib().cb200_nucleus_fill(_ptr(seg), _ptr(raw), _code(raw, _RAW_DTYPES), seg.ndim, spatial_array(seg.shape),
                                   n_inst, _ptr(ids), _ptr(thresholds), _ptr(boxes), _ptr(box_offset),
                                   int(total_box_voxels), _ptr(out), _ptr(ws), _stream(seg))
    check(rc, "cb200_nucleus_fill")
    launch_counter["calls"] += 1
    return out


# --------------------------------------------------------------------------- evaluation counts
_LABEL_DTYPES = (torch.uint16, torch.int32)


def label_presence(labels: torch.Tensor, max_value: int) -> torch.Tensor:
    """`cb200_label_presence`: uint8 (max_value + 1), 1 where the value occurs."""
    _require_cuda(labels)
    assert labels.is_contiguous()
    present = torch.empty(max_value + 1, dtype=torch.uint8, device=labels.device)
    rc = _lib().cb200_label_presence(_ptr(labels), _code(labels, _LABEL_DTYPES), labels.numel(), int(max_value),
                                     _ptr(present), _stream(labels))
    check(rc, "cb200_label_presence")
    launch_counter["calls"] += 1
    return present


def contingency(pred: torch.Tensor, gt: torch.Tensor, rank_pred: torch.Tensor, rank_gt: torch.Tensor, rows: int,
                cols: int) -> torch.Tensor:
    """`cb200_contingency`: (rows, cols) int32 table of joint label counts."""
    _require_cuda(pred, gt, rank_pred, rank_gt)
    assert pred.dtype == gt.dtype and pred.shape == gt.shape and pred.is_contiguous() and gt.is_contiguous()
    assert rank_pred.dtype == torch.int32 and rank_gt.dtype == torch.int32
    table = torch.empty((rows, cols), dtype=torch.int32, device=pred.device)
    assert rank_pred.numel() == rank_gt.numel()
    rc = _lib().cb200_contingency(_ptr(pred), _ptr(gt), _code(pred, _LABEL_DTYPES), pred.numel(), _ptr(rank_pred),
                                  _ptr(rank_gt), int(rank_pred.numel()) - 1, int(rows), int(cols), _ptr(table),
                                  _stream(pred))
    check(rc, "cb200_contingency")
    launch_counter["calls"] += 1
    return table


# --------------------------------------------------------------------------- one call per volume
_detect_workspaces = {}  # (device index, stream) -> (uint8 workspace tensor, foreground capacity it was sized for)


def detect_volume(emb: torch.Tensor, bandwidth: float, threshold: float, reduction_probability: float = 1.0,
                  philox_seed: int = 0, max_iter: int = 300, label_dtype=torch.int32, want_mask: bool = False,
                  centre_capacity: int = 0):
    """`cb200_detect_volume`: threshold -> labels in ONE C-ABI call (the count reads are handled inside the
    library, the scratch is a torch buffer kept per device and grown on demand).  Returns `(labels (*S), mask | None, centres (D, centre_capacity) | None, info dict)`.
    Raises ValueError with scikit-learn's messages where `MeanShift.fit` would."""
    _require_cuda(emb)
    emb = emb.contiguous()
    D = emb.shape[0] - 1
    spatial = tuple(emb.shape[1:])
    if len(spatial) != D or D not in (2, 3):
        raise ValueError("emb must be (D+1, *S) with D spatial dims, D in {2, 3}")
    dev = emb.device
    labels = torch.empty(spatial, dtype=label_dtype, device=dev)
    mask = torch.empty(spatial, dtype=torch.uint8, device=dev) if want_mask else None
    centres = torch.zeros((D, centre_capacity), dtype=torch.float64, device=dev) if centre_capacity > 0 else None
    info = _cabi.DetectInfo()
    n_pix = int(np.prod(spatial))
    # scratch is the caller's: one buffer per device and stream, sized for the largest foreground count seen so far
    # (first guess: a quarter of the pixels); CB200_ENOSPACE reports what the counts found on the device call for
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws, capacity = _detect_workspaces.get(key, (None, 0))
    if ws is None:
        capacity = max(n_pix // 4, 1 << 16)
        ws = torch.empty(_lib().cb200_detect_volume_workspace_bytes(D, spatial_array(spatial), capacity,
                                                                    float(reduction_probability)),
                         dtype=torch.uint8, device=dev)
    for _ in range(6):
        rc = _lib().cb200_detect_volume(
            _ptr(emb), _code(emb, _FLOAT_DTYPES), D, spatial_array(spatial), float(threshold), float(bandwidth),
            float(reduction_probability), int(philox_seed) & (2**64 - 1), int(max_iter), _ptr(labels),
            _code(labels, (torch.int32, torch.uint16)), _ptr(mask), _DTYPE_CODE[torch.uint8] if want_mask else 0,
            _ptr(centres), int(centre_capacity), _ptr(ws), ws.numel(), min(capacity, n_pix), C.byref(info),
            _stream(emb))
        if rc != _cabi.ENOSPACE:
            break
        capacity = max(capacity, int(info.n_foreground) + (int(info.n_foreground) >> 3))
        need = max(int(info.workspace_needed) + (int(info.workspace_needed) >> 3),
                   _lib().cb200_detect_volume_workspace_bytes(D, spatial_array(spatial), capacity,
                                                              float(reduction_probability)))
        ws = None  # release before growing
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
    _detect_workspaces[key] = (ws, capacity)
    launch_counter["calls"] += 1
    if rc == _cabi.ENOFIT:
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by MeanShift.")
    if rc == _cabi.ENOCENTRE:
        raise ValueError(
            "No point was within bandwidth=%f of any seed. Try a different seeding strategy "
            "                             or increase the bandwidth." % bandwidth)
    check(rc, "cb200_detect_volume")
    return labels, mask, centres, {"n_fg": int(info.n_foreground), "n_fit": int(info.n_fit),
                                   "n_seeds": int(info.n_seeds), "k": int(info.n_centres), "method": "grid",
                                   "grid_cells": int(info.grid.n_cells), "suppress_calls": int(info.suppress_calls),
                                   "distance_tests": int(info.distance_tests), "climb_steps": int(info.climb_steps),
                                   "distinct_modes": int(info.n_distinct_modes)}


def release_scratch(device=None) -> None:
    """Drop the workspace buffers `detect_volume` keeps (all devices, or one)."""
    for key in [k for k in _detect_workspaces if device is None or k[0] == torch.device(device).index]:
        del _detect_workspaces[key]
