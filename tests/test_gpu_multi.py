"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): seed-sharded mean-shift on the CUDA kernels equals
the single-GPU pipeline bit for bit."""

import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, emb_np, results):
    import torch.distributed as dist

    from cellulus_b200 import kernels as K
    from cellulus_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        emb = torch.from_numpy(emb_np).to(dev)
        pts, pix, n, _ = K.fg_compact(emb, 0.5)
        flags = K.bernoulli_flags(n, 0.3, 11, dev)
        mine = sharding.shard_items(n, rank, world)  # this rank's slab of the raster-ordered foreground
        local = pts[:, mine.start:mine.stop].contiguous()
        labels, centres = sharding.sharded_mean_shift(local, local.shape[1], 5.0, sharding.cuda_ops("grid"),
                                                      flags[mine.start:mine.stop], None)
        results[rank] = (labels.cpu().numpy(), centres.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_seed_sharded_mean_shift_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from cellulus_b200 import kernels as K
    from cellulus_b200 import synthetic
    from cellulus_b200.utils.mean_shift import segment_embeddings_device

    emb_np, _, _ = synthetic.blob_scene((40, 96, 96), 30, radius=8.0, seed=2)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(2, port, emb_np, results), nprocs=2, join=True)
    # single-GPU pipeline with the same device-generated fit subset
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    emb = torch.from_numpy(emb_np).to(dev)
    _, _, n, _ = K.fg_compact(emb, 0.5)
    flags = K.bernoulli_flags(n, 0.3, 11, dev)
    labels, info = segment_embeddings_device(emb, 5.0, 0.5, 0.3, fit_flags=flags, method="grid")
    ref = labels[emb[3] < 0.5].cpu().numpy()
    got = np.concatenate([results[r][0] for r in range(2)])
    assert np.array_equal(got, ref)
    for r in range(2):
        assert np.array_equal(results[r][1], info["centres"].cpu().numpy())
