"""BASELINE configs[4]: blockwise inference of a 16k x 16k 2-D mosaic and a 1024^3 3-D volume, scan blocks
dealt round-robin to 1/2/4/8 GPUs (one process per GPU, no data-path collective -- blocks are independent).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/blockwise_bench.py [2d|3d|both] [scale]

Per block (the hot path only; the U-Net forward is not the product): a stack of T = 32 noisy per-pixel
embeddings, generated on the device outside the timed region, goes through
    TTA aggregate -> threshold -> foreground compaction -> mean-shift -> centre suppression -> labels.
Timed with CUDA events per block; the job time is the MAX over ranks of the summed block times.
`scale` < 1 shrinks the volume (fewer blocks) for quick runs.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellulus_b200 import sharding  # noqa: E402
from cellulus_b200.detect import detect_embeddings  # noqa: E402
from cellulus_b200.models import tta_aggregate  # noqa: E402
from cellulus_b200.synthetic import block_stack  # noqa: E402


def run(kind, scale, dev, rank, world):
    if kind == "2d":
        total, block, radius, bw = (int(16384 * scale),) * 2, (1024, 1024), 10.0, 7.0
    else:
        total, block, radius, bw = (int(1024 * scale),) * 3, (128, 256, 256), 10.0, 7.0
    total = tuple(max(t, b) for t, b in zip(total, block))
    blocks = sharding.scan_blocks(total, block)
    mine = [blocks[i] for i in sharding.shard_round_robin(len(blocks), rank, world)]
    D = len(block)
    thr = 0.5 * D
    ms, labelled = 0.0, 0
    # warm-up block
    st = block_stack(block, radius, 32, 10_000 + rank, dev)
    detect_embeddings(tta_aggregate(st), bandwidth=bw, threshold=thr, reduction_probability=0.1, rng="philox")
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    for b in mine:
        seed = int(np.ravel_multi_index(tuple(o // s for o, s in zip(b, block)), tuple(t // s + 1 for t, s in zip(total, block))))
        st = block_stack(block, radius, 32, seed, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        emb = tta_aggregate(st)
        labels, _, _ = detect_embeddings(emb, bandwidth=bw, threshold=thr, reduction_probability=0.1, rng="philox")
        e1.record()
        torch.cuda.synchronize(dev)
        ms += e0.elapsed_time(e1)
        labelled += int((labels.to(torch.int32) > 0).sum())
        del st, emb, labels
    t = torch.tensor([ms], device=dev)
    n = torch.tensor([labelled], device=dev, dtype=torch.int64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(n)
    px = float(np.prod(block)) * len(blocks)
    if rank == 0:
        print(json.dumps({"workload": f"configs[4] blockwise {kind}: {'x'.join(map(str, total))} in {len(blocks)} blocks of "
                                      f"{'x'.join(map(str, block))}, T=32 TTA aggregate + detect (bw {bw}, rp 0.1)",
                          "n_gpus": world, "job_ms": t.item(), "Mpx_per_s": px / t.item() / 1e3,
                          "foreground_px": int(n.item()), "blocks_per_gpu": len(mine)}), flush=True)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "both"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    world, rank, local = (int(os.environ.get(k, d)) for k, d in [("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    for k in (["2d", "3d"] if kind == "both" else [kind]):
        run(k, scale, dev, rank, world)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
