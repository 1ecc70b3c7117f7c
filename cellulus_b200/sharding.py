"""Multi-GPU sharding of the hot path (one process per GPU, `torch.distributed`).

The path shards three ways, none of which needs a collective on the data path
except the last:

* training: by batch -- the loss kernel is rank-local (the loss is a plain sum over samples,
  `criterions/oce_loss.py:58-62`); only the U-Net's parameter gradients are all-reduced (DDP).
* inference over many samples / scan blocks: independent units (`detect.py:82`, `predict.py:129`),
  dealt round-robin or in contiguous ranges -> `shard_items`, `shard_round_robin`, `scan_blocks`.
* ONE huge volume: seeds are independent given the full point set (sklearn `_mean_shift.py:506-509`).
  Each rank compacts its slab, the point sets are all-gathered, every rank climbs its slice of the
  seeds, the converged (mode, count) lists are all-gathered, centre suppression is replicated
  (deterministic) and labels are assigned per slab -> `sharded_mean_shift`.

The functions take the compute steps as callables so the same composition runs on the CUDA kernels
(NCCL) and, in the CPU tests, on the oracle (gloo, world_size 2).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_items(n_items: int, rank: int, world: int) -> range:
    """Contiguous balanced range of `n_items` units for `rank` (the first n % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def shard_round_robin(n_items: int, rank: int, world: int) -> range:
    """Units rank, rank + world, ... (scan blocks of similar cost are dealt like cards)."""
    return range(rank, n_items, world)


def scan_blocks(spatial: Sequence[int], block: Sequence[int]) -> List[Tuple[int, ...]]:
    """Offsets of the output blocks a gunpowder `Scan` visits (`predict.py:62-93,129`): strides of the
    block size, the last block of every axis shifted INWARD (not shrunk) so that it ends at the border.
    A block larger than the volume is clamped to offset 0."""
    axes = []
    for s, b in zip(spatial, block):
        if b >= s:
            axes.append([0])
            continue
        offs = list(range(0, s - b + 1, b))
        if offs[-1] + b < s:
            offs.append(s - b)
        axes.append(offs)
    out = [()]
    for offs in axes:
        out = [o + (v,) for o in out for v in offs]
    return out


def owned_extents(spatial: Sequence[int], block: Sequence[int]) -> List[Tuple[Tuple[int, ...], Tuple[int, ...]]]:
    """For every block of `scan_blocks(spatial, block)`, in the same order: `(offset, extent)` of the part of the
    block that no LATER block of the scan overwrites.  Walking the blocks in scan order and writing each one whole
    (what a single process does, `predict.py:129`) leaves every output pixel with the value of the last block that
    covers it; blocks only overlap where the last block of an axis is shifted inward, and there the later block is
    the one with the larger offset.  Writing only the owned part of every block therefore gives the same volume in
    ANY order -- which is what lets the blocks of one sample be dealt to different ranks and summed."""
    axes = []
    for s, b in zip(spatial, block):
        offs = sorted({o[len(axes)] for o in scan_blocks(spatial, block)})
        ends = {o: min(o + b, s, offs[i + 1] if i + 1 < len(offs) else s) for i, o in enumerate(offs)}
        axes.append(ends)
    return [(off, tuple(axes[a][o] - o for a, o in enumerate(off))) for off in scan_blocks(spatial, block)]


SEED_CHUNK = 2048


def seed_indices(n_seeds: int, rank: int, world: int, device=None) -> torch.Tensor:
    """Indices of the seeds `rank` climbs: chunks rank, rank + world, ... of SEED_CHUNK consecutive seeds."""
    if world == 1:
        return torch.arange(n_seeds, device=device)
    n_chunks = -(-n_seeds // SEED_CHUNK)
    if rank >= n_chunks:
        return torch.zeros(0, dtype=torch.int64, device=device)
    starts = torch.arange(rank, n_chunks, world, device=device) * SEED_CHUNK
    idx = (starts[:, None] + torch.arange(SEED_CHUNK, device=device)[None, :]).reshape(-1)
    return idx[idx < n_seeds]


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_columns(local: torch.Tensor, n_local: int, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather of a ragged SoA array: `local` is (R, >= n_local); returns ((R, sum n), counts per rank),
    rank-major.  Works on CPU tensors over gloo and CUDA tensors over NCCL (pads to the largest shard)."""
    rank, world = _world(group)
    if world == 1:
        return local[:, :n_local].contiguous(), [n_local]
    counts_t = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts_t[rank] = n_local
    dist.all_reduce(counts_t, group=group)
    counts = [int(c) for c in counts_t.tolist()]
    width = max(max(counts), 1)
    send = torch.zeros((local.shape[0], width), dtype=local.dtype, device=local.device)
    send[:, :n_local] = local[:, :n_local]
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([r[:, :c] for r, c in zip(recv, counts)], dim=1).contiguous(), counts


@dataclass
class MeanShiftOps:
    """The compute steps `sharded_mean_shift` composes (CUDA kernels in production, oracle in CPU tests).

    climb(points (D,n), seeds (D,s), bandwidth)           -> (modes (D,s), counts (s,), iters (s,))
    suppress(modes (D,s), counts (s,), bandwidth, points)  -> centres (D,k) in priority order
    assign(points (D,n), centres (D,k))                    -> labels (n,) int, 1 + nearest centre
    dedupe(modes (D,s), counts (s,))                       -> (modes (D,u), counts (u,)): one copy of every
        bit-identical mode with count > 0 (optional; seeds that end in the same window end in the SAME mean, so a
        rank's share of the modes shrinks by orders of magnitude before it is gathered and suppressed)
    """

    climb: Callable
    suppress: Callable
    assign: Callable
    dedupe: Optional[Callable] = None


class _PhaseTimer:
    """CUDA-event stopwatch for the phases of `sharded_mean_shift` (device time, this rank)."""

    def __init__(self, enabled, device):
        self.on = enabled and device.type == "cuda"
        self.marks = []
        self.device = device

    def mark(self, name):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def result(self):
        if not self.on or len(self.marks) < 2:
            return {}
        torch.cuda.synchronize(self.device)
        return {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks[:-1], self.marks[1:])}


def sharded_mean_shift(local_points: torch.Tensor, n_local: int, bandwidth: float, ops: MeanShiftOps,
                       local_fit_flags: Optional[torch.Tensor] = None, group=None, timings: Optional[dict] = None):
    """Mean-shift of one point set that is spread over the ranks (slab order = rank order).

    local_points (D, >= n_local) float64 SoA: this rank's foreground points in raster order.
    local_fit_flags (n_local,) uint8/bool or None: which of them take part in `fit` (the
    `reduction_probability` subset, `utils/mean_shift.py:68-70`).
    Returns `(labels_local (n_local,), centres (D, k))`; identical to running the single-GPU pipeline on
    the concatenated point set.  `timings` (optional dict) receives the device milliseconds per phase.
    """
    rank, world = _world(group)
    D = local_points.shape[0]
    clock = _PhaseTimer(timings is not None, local_points.device)
    clock.mark("start")
    if local_fit_flags is None:
        fit_local, n_fit_local = local_points[:, :n_local], n_local
    else:
        keep = local_fit_flags[:n_local].to(torch.bool)
        fit_local = local_points[:, :n_local][:, keep]
        n_fit_local = int(keep.sum())
    # exchange 1: every rank needs the whole fit set (N x D doubles; O(ms) over NVSwitch)
    fit_all, _ = all_gather_columns(fit_local.contiguous(), n_fit_local, group)
    clock.mark("gather_points")
    n_fit = fit_all.shape[1]
    if n_fit == 0:
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by MeanShift.")
    # seeds = all fit points (sklearn:491-496), dealt to the ranks in chunks of SEED_CHUNK consecutive seeds
    # (block-cyclic: climbing cost varies over the volume, a contiguous split leaves ranks waiting for the slowest)
    idx = [seed_indices(n_fit, r, world, fit_all.device) for r in range(world)]
    mine = idx[rank]
    seeds = fit_all[:, mine].contiguous() if world > 1 else fit_all.clone()
    modes, counts, _ = ops.climb(fit_all, seeds, bandwidth)
    clock.mark("climb")
    # exchange 2: converged (mode, count) of every seed.  With a dedupe op every rank first merges its bit-identical
    # modes (exact: the suppression keeps the same copy); the order of the gathered modes does not matter to the
    # suppression (it orders them by (count, coordinates) itself).  Without one they go back into global seed order.
    n_mine = int(mine.numel())
    modes_mine, counts_mine = modes[:, :n_mine], counts[:n_mine]
    deduped = ops.dedupe is not None and n_mine > 0
    if deduped:
        modes_mine, counts_mine = ops.dedupe(modes_mine, counts_mine)
        n_mine = int(modes_mine.shape[1])
    packed = torch.cat([modes_mine[:, :n_mine], counts_mine[:n_mine].to(modes.dtype)[None]], dim=0)
    packed_all, _ = all_gather_columns(packed.contiguous(), n_mine, group)
    if world > 1 and ops.dedupe is None:
        ordered = torch.empty_like(packed_all)
        ordered[:, torch.cat(idx)] = packed_all
        packed_all = ordered
    modes_all = packed_all[:D].contiguous()
    counts_all = packed_all[D].round().to(torch.int32).contiguous()
    clock.mark("gather_modes")
    # centre suppression is deterministic: replicate it instead of broadcasting its result
    centres = ops.suppress(modes_all, counts_all, bandwidth, fit_all)
    clock.mark("suppress")
    labels = ops.assign(local_points[:, :n_local].contiguous(), centres)
    clock.mark("assign")
    if timings is not None:
        timings.update(clock.result())
    return labels, centres


def cuda_ops(method: str = "auto") -> MeanShiftOps:
    """`MeanShiftOps` on the B200 kernels.  The three steps share one cell grid (edge >= bandwidth over the
    bounding box of the fit points): the hill climb hashes the points into it, the suppression derives its fine
    grid from it, the label assignment hashes the centres into it (pruned nearest-centre search)."""
    from cellulus_b200 import kernels as K
    from cellulus_b200.utils import mean_shift as MS

    ctx = {}

    def pad(t):  # kernels want an even stride (16-byte bulk copies)
        n = t.shape[1]
        cap = max(2, (n + 1) & ~1)
        if t.stride(0) == cap and t.is_contiguous():
            return t
        out = torch.zeros((t.shape[0], cap), dtype=torch.float64, device=t.device)
        out[:, :n] = t
        return out

    def grid_of(points, bandwidth):
        key = (points.data_ptr(), points.shape[1], float(bandwidth))
        if ctx.get("key") != key:
            lo, hi = K.bounding_box(pad(points), points.shape[1])
            ctx["key"], ctx["grid"] = key, K.plan_grid(lo, hi, bandwidth)
        return ctx["grid"]

    def climb(points, seeds, bandwidth):
        n, s = points.shape[1], seeds.shape[1]
        pts, sd = pad(points), pad(seeds)
        grid = grid_of(points, bandwidth)
        use = method
        if use == "auto":
            use = "brute" if n * s <= MS._BRUTE_PAIR_LIMIT else "grid"
        if use == "grid":
            sorted_pts, cell_start, _ = K.grid_build(pts, n, grid)
            # distinct trajectories only: the merged copies come back with count 0 and are dropped by `dedupe`
            counts, iters = K.ms_grid_modes_distinct(sorted_pts, n, grid, cell_start, sd, s, bandwidth,
                                                     merge_rounds=K.default_merge_rounds(max(n, s)))
        else:
            counts, iters = K.ms_brute_modes(pts, n, sd, s, bandwidth)
        return sd[:, :s], counts[:s], iters[:s]

    def suppress(modes, counts, bandwidth, points):
        grid = grid_of(points, bandwidth)
        m = pad(modes)
        centres, k = K.nms_centres(m, counts.contiguous(), modes.shape[1], bandwidth, grid)
        if k == 0:
            raise ValueError("No point was within bandwidth=%f of any seed." % bandwidth)
        return centres[:, :k].contiguous()

    def assign(points, centres):
        n = points.shape[1]
        labels = torch.zeros(max(n, 1), dtype=torch.int32, device=points.device)
        if n:
            K.assign_labels(pad(points), n, pad(centres), centres.shape[1], None, labels, grid=ctx.get("grid"))
        return labels[:n]

    def dedupe(modes, counts):
        s = modes.shape[1]
        m, c, u = K.unique_modes(pad(modes), counts.contiguous(), s)
        return m[:, :u], c[:u]

    return MeanShiftOps(climb, suppress, assign, dedupe)
