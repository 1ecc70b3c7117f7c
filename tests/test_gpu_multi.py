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


def _detect_worker(rank, world, port, container, results):
    """One rank of `detect(inference_config)` under a torchrun-like environment."""
    import torch.distributed as dist

    from cellulus_b200.configs import DatasetConfig, InferenceConfig
    from cellulus_b200.detect import detect

    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world)})
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        np.random.seed(5)  # rank 0 draws the fit subset (np.random.rand over all foreground points)
        cfg = InferenceConfig(
            dataset_config=DatasetConfig(container_path=container, dataset_name="embeddings"),
            prediction_dataset_config=DatasetConfig(container_path=container, dataset_name="embeddings"),
            detection_dataset_config=DatasetConfig(container_path=container, dataset_name=f"detection_w{world}",
                                                   secondary_dataset_name="embeddings"),
            segmentation_dataset_config=DatasetConfig(container_path=container, dataset_name="segmentation"),
            evaluation_dataset_config=DatasetConfig(container_path=container, dataset_name="gt"),
            device=f"cuda:{rank}", threshold=0.5, bandwidth=5.0, num_bandwidths=2, reduction_probability=0.3,
            crop_size=[40, 96, 96])
        detect(cfg)
    finally:
        if world > 1:
            dist.destroy_process_group()
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
            os.environ.pop(k, None)


@pytest.mark.timeout(600)
def test_detect_splits_one_sample_by_seed_over_two_gpus(tmp_path):
    """`detect()` with ONE sample on TWO ranks takes the seed-sharded path (two all-gathers); the `detection`
    dataset it writes equals the one a single rank writes (same numpy seed for the fit subset)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from cellulus_b200 import synthetic, zarr_lite

    emb_np, _, _ = synthetic.blob_scene((40, 96, 96), 30, radius=8.0, seed=2)
    container = str(tmp_path / "one_sample.zarr")
    g = zarr_lite.open(container)
    a = g.create_dataset("embeddings", shape=(1, 4, 40, 96, 96), dtype=float)
    a[0] = emb_np.astype(np.float64)
    a.attrs["axis_names"] = ["s", "c", "z", "y", "x"]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    results = mp.Manager().dict()
    # binary-segmentation / centered-embeddings are created with fixed names (detect.py:40-80): one container each
    import shutil

    container2 = str(tmp_path / "one_sample_w2.zarr")
    shutil.copytree(container, container2)
    mp.spawn(_detect_worker, args=(1, port, container, results), nprocs=1, join=True)
    mp.spawn(_detect_worker, args=(2, port + 1, container2, results), nprocs=2, join=True)
    one = zarr_lite.open(container, "r")["detection_w1"][...]
    two = zarr_lite.open(container2, "r")["detection_w2"][...]
    assert one.shape == (1, 2, 40, 96, 96) and one.dtype == np.uint16 and one.max() > 5
    assert np.array_equal(one, two)
    for name in ("binary-segmentation", "centered-embeddings"):
        assert np.array_equal(zarr_lite.open(container, "r")[name][...], zarr_lite.open(container2, "r")[name][...])


def _predict_worker(rank, world, port, container):
    """One rank of `predict(model, inference_config, None)` under a torchrun-like environment."""
    import torch.distributed as dist

    from cellulus_b200.configs import DatasetConfig, InferenceConfig
    from cellulus_b200.models import get_model
    from cellulus_b200.predict import predict

    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world)})
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        torch.manual_seed(0)  # same random-init weights on every rank and in both runs
        model = get_model(1, 2, 4, 2, 8, [(2, 2)], 2).cuda().eval()
        cfg = InferenceConfig(
            dataset_config=DatasetConfig(container_path=container, dataset_name="raw"),
            prediction_dataset_config=DatasetConfig(container_path=container, dataset_name=f"embeddings_w{world}"),
            detection_dataset_config=DatasetConfig(container_path=container, dataset_name="detection",
                                                   secondary_dataset_name=f"embeddings_w{world}"),
            segmentation_dataset_config=DatasetConfig(container_path=container, dataset_name="segmentation"),
            evaluation_dataset_config=DatasetConfig(container_path=container, dataset_name="gt"),
            device=f"cuda:{rank}", crop_size=[60, 60], p_salt_pepper=0.0, num_infer_iterations=2)
        predict(model, cfg, None)
    finally:
        if world > 1:
            dist.destroy_process_group()
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
            os.environ.pop(k, None)


@pytest.mark.timeout(600)
def test_predict_deals_scan_blocks_of_one_sample_over_two_gpus(tmp_path):
    """`predict()` with ONE sample on TWO ranks deals the scan blocks (BASELINE configs[4]) and sums the partial
    volumes onto rank 0; with the test-time noise switched off (p = 0) the `embeddings` dataset equals the one a
    single rank writes, inward-shifted border blocks included."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from cellulus_b200 import zarr_lite

    container = str(tmp_path / "mosaic.zarr")
    g = zarr_lite.open(container)
    rng = np.random.default_rng(0)
    a = g.create_dataset("raw", shape=(1, 1, 150, 170), dtype=np.float32)
    a[0] = rng.random((1, 150, 170), dtype=np.float32)
    a.attrs["axis_names"] = ["s", "c", "y", "x"]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_predict_worker, args=(1, port, container), nprocs=1, join=True)
    mp.spawn(_predict_worker, args=(2, port + 1, container), nprocs=2, join=True)
    f = zarr_lite.open(container, "r")
    one, two = f["embeddings_w1"][...], f["embeddings_w2"][...]
    assert one.shape == (1, 3, 150, 170) and np.isfinite(one).all() and np.abs(one[0, :2]).max() > 0
    assert np.array_equal(one, two)
    assert f["embeddings_w2"].attrs["axis_names"] == ["s", "c", "y", "x"]
