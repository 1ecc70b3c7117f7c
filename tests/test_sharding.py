"""CPU, world_size 2 over gloo: the multi-GPU composition (seed-sharded mean-shift with two all-gathers,
block / sample sharding) produces exactly what the single-process oracle does."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cellulus_b200 import sharding, synthetic
from oracle import mean_shift as oms


def test_shard_ranges_cover_everything():
    for n in [0, 1, 7, 64, 1001]:
        for world in [1, 2, 3, 8]:
            parts = [list(sharding.shard_items(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            rr = sorted(sum([list(sharding.shard_round_robin(n, r, world)) for r in range(world)], []))
            assert rr == list(range(n))


def test_seed_chunks_are_dealt_block_cyclic():
    for n in [0, 1, 2047, 2048, 2049, 10_000]:
        for world in [1, 2, 3, 8]:
            parts = [sharding.seed_indices(n, r, world) for r in range(world)]
            assert sorted(torch.cat(parts).tolist()) == list(range(n))
            if world > 1 and n > sharding.SEED_CHUNK:
                assert parts[1][0].item() == sharding.SEED_CHUNK  # rank 1 starts with the second chunk


def test_scan_blocks_shift_last_block_inward():
    # gunpowder Scan semantics (predict.py:129): stride = block, last block shifted inward, never shrunk
    assert sharding.scan_blocks((10,), (4,)) == [(0,), (4,), (6,)]
    assert sharding.scan_blocks((8,), (4,)) == [(0,), (4,)]
    assert sharding.scan_blocks((3,), (4,)) == [(0,)]
    blocks = sharding.scan_blocks((1000, 700), (236, 236))
    covered = np.zeros((1000, 700), bool)
    for y, x in blocks:
        assert 0 <= y <= 1000 - 236 and 0 <= x <= 700 - 236
        covered[y:y + 236, x:x + 236] = True
    assert covered.all()
    assert len(blocks) == 5 * 3


def _oracle_ops():
    def climb(points, seeds, bandwidth):
        m, c, i = oms.mean_shift_modes(points.numpy().T, seeds.numpy().T, bandwidth)
        return torch.from_numpy(np.ascontiguousarray(m.T)), torch.from_numpy(c), torch.from_numpy(i)

    def suppress(modes, counts, bandwidth, points):
        c = oms.nms_centres(modes.numpy().T, counts.numpy(), bandwidth)
        return torch.from_numpy(np.ascontiguousarray(c.T))

    def assign(points, centres):
        return torch.from_numpy(oms.predict_labels(points.numpy().T, centres.numpy().T) + 1)

    return sharding.MeanShiftOps(climb, suppress, assign)


def _scene():
    emb, _, _ = synthetic.blob_scene((48, 56), 6, radius=6.0, seed=3)
    mask = emb[2].astype(np.float64) < 0.5
    X = oms.points_from_embedding(emb[:2], mask)
    rng = np.random.default_rng(0)
    fit = rng.random(len(X)) < 0.4
    return X, fit


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding.SEED_CHUNK = 16  # a few hundred seeds here: small chunks so that both ranks climb interleaved ranges
    try:
        X, fit = _scene()
        mine = sharding.shard_items(len(X), rank, world)  # slab = contiguous raster range
        local = torch.from_numpy(np.ascontiguousarray(X[mine.start:mine.stop].T))
        flags = torch.from_numpy(fit[mine.start:mine.stop].astype(np.uint8))
        labels, centres = sharding.sharded_mean_shift(local, local.shape[1], 4.0, _oracle_ops(), flags)
        # ragged all-gather round trip
        gathered, counts = sharding.all_gather_columns(local, local.shape[1])
        assert counts == [len(sharding.shard_items(len(X), r, world)) for r in range(world)]
        assert np.array_equal(gathered.numpy().T, X)
        results[rank] = (labels.numpy(), centres.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_seed_sharded_mean_shift_matches_single_process():
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    X, fit = _scene()
    ref_labels, ref_centres = oms.segment_points(X, fit, 4.0)
    got = np.concatenate([results[r][0] for r in range(world)])
    assert np.array_equal(got, ref_labels + 1)
    for r in range(world):
        assert np.array_equal(results[r][1].T, ref_centres)  # replicated, bit-identical on every rank


def test_single_process_path_needs_no_process_group():
    X, fit = _scene()
    local = torch.from_numpy(np.ascontiguousarray(X.T))
    labels, centres = sharding.sharded_mean_shift(local, len(X), 4.0, _oracle_ops(),
                                                  torch.from_numpy(fit.astype(np.uint8)))
    ref_labels, ref_centres = oms.segment_points(X, fit, 4.0)
    assert np.array_equal(labels.numpy(), ref_labels + 1)
    assert np.array_equal(centres.numpy().T, ref_centres)


@pytest.mark.parametrize("spatial,block", [((10,), (4,)), ((50, 70), (16, 32)), ((9, 20, 33), (4, 8, 16)), ((5, 5), (8, 8)),
                                           ((32, 32), (16, 16)), ((33, 17), (16, 16))])
def test_owned_extents_partition_the_volume_like_a_sequential_scan(spatial, block):
    """Writing only the owned part of every scan block (in any order, by any rank) leaves the volume exactly as a
    sequential scan that writes every block whole (`predict.py:129`): each pixel belongs to the LAST block covering it."""
    from cellulus_b200 import sharding

    blocks = sharding.scan_blocks(spatial, block)
    owned = sharding.owned_extents(spatial, block)
    assert [o for o, _ in owned] == blocks
    sequential = np.zeros(spatial, np.int64)
    for i, off in enumerate(blocks):
        sequential[tuple(slice(o, min(o + b, s)) for o, b, s in zip(off, block, spatial))] = i + 1
    summed, cover = np.zeros(spatial, np.int64), np.zeros(spatial, np.int64)
    for rank in range(3):  # three "ranks", blocks dealt round-robin, partial volumes summed
        part = np.zeros(spatial, np.int64)
        for i in range(rank, len(owned), 3):
            off, ext = owned[i]
            sl = tuple(slice(o, o + e) for o, e in zip(off, ext))
            part[sl] = i + 1
            cover[sl] += 1
        summed += part
    assert (cover == 1).all()
    assert np.array_equal(summed, sequential)
