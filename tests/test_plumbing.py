"""CPU: the drop-in boundary around the hot path -- TOML schema, zarr layout, dataset/sampler, U-Net stand-in
geometry and checkpoint key names, scan blocks.  No kernels are launched here."""

import os
import tomllib

import numpy as np
import pytest
import torch

from cellulus_b200 import zarr_lite
from cellulus_b200.configs import DatasetConfig, ExperimentConfig
from cellulus_b200.datasets import get_dataset
from cellulus_b200.datasets.meta_data import DatasetMetaData
from cellulus_b200.models import get_model
from oracle import sampler as osampler

# tests/train.toml of the reference, repaired as SURVEY §4 describes (object_size must be an int)
TRAIN_TOML = """
experiment_name = "Train test"
object_size = 10

[model_config]
num_fmaps = 12
fmap_inc_factor = 2

[train_config]
batch_size = 32

[train_config.train_data_config]
container_path = "test_data.zarr"
dataset_name = "train"

[train_config.validate_data_config]
container_path = "test_data.zarr"
dataset_name = "validate"
"""


def test_config_schema_matches_reference_defaults():
    cfg = ExperimentConfig(**tomllib.loads(TRAIN_TOML))
    assert cfg.model_config.num_fmaps == 12 and cfg.model_config.features_in_last_layer == 64
    assert cfg.model_config.downsampling_factors == [[2, 2]] and cfg.model_config.initialize is True
    t = cfg.train_config
    assert (t.crop_size, t.batch_size, t.max_iterations) == ([252, 252], 32, 100_000)
    assert (t.initial_learning_rate, t.density, t.kappa, t.temperature, t.regularizer_weight) == (4e-5, 0.1, 10.0, 10.0, 1e-5)
    assert (t.save_model_every, t.save_best_model_every, t.save_snapshot_every, t.num_workers) == (1000, 100, 1000, 8)
    assert t.device == "cuda:0" and t.elastic_deform is True
    assert isinstance(t.train_data_config, DatasetConfig) and t.train_data_config.dataset_name == "train"
    assert cfg.inference_config is None and cfg.normalization_factor is None
    with pytest.raises(TypeError):  # the reference's own toml has object_size = 10.0 and fails this validator
        ExperimentConfig(**tomllib.loads(TRAIN_TOML.replace("object_size = 10", "object_size = 10.0")))
    inf = ExperimentConfig(**tomllib.loads(TRAIN_TOML + """
[inference_config]
num_bandwidths = 2
[inference_config.dataset_config]
container_path = "x.zarr"
dataset_name = "test"
""")).inference_config
    assert (inf.p_salt_pepper, inf.num_infer_iterations, inf.reduction_probability) == (0.01, 16, 0.1)
    assert (inf.clustering, inf.use_seeds, inf.post_processing, inf.grow_distance, inf.shrink_distance) == (
        "meanshift", False, "cell", 3, 6)
    assert inf.threshold is None and inf.bandwidth is None and inf.min_size is None


def test_zarr_lite_layout_roundtrip(tmp_path):
    g = zarr_lite.open(tmp_path / "c.zarr")
    x = np.random.default_rng(0).random((3, 2, 40, 50)).astype(np.float32)
    a = g.create_dataset("raw", shape=x.shape, dtype=np.float32, chunks=(1, 2, 16, 32))
    a[...] = x
    a.attrs["axis_names"] = ["s", "c", "y", "x"]
    b = zarr_lite.open(tmp_path / "c.zarr", "r")["raw"]
    assert np.array_equal(b[...], x) and np.array_equal(b[1], x[1]) and np.array_equal(b[2, :, 5:33, 7:41], x[2, :, 5:33, 7:41])
    a[1, 0, 3:20, 10:45] = 7
    x[1, 0, 3:20, 10:45] = 7
    assert np.array_equal(b[...], x)
    with pytest.raises(ValueError):  # zarr v2: create_dataset on an existing array without overwrite fails
        g.create_dataset("raw", shape=(1,), dtype=float)
    # on-disk format: .zgroup / .zarray (zarr_format 2) / .zattrs / dotted chunk keys
    assert os.path.exists(tmp_path / "c.zarr" / ".zgroup") and os.path.exists(tmp_path / "c.zarr" / "raw" / "0.0.0.0")
    meta = DatasetMetaData.from_dataset_config(DatasetConfig(container_path=tmp_path / "c.zarr", dataset_name="raw"))
    assert (meta.num_samples, meta.num_channels, meta.num_spatial_dims, meta.spatial_array) == (3, 2, 2, (40, 50))
    with pytest.raises(RuntimeError, match="no dataset"):
        DatasetMetaData.from_dataset_config(DatasetConfig(container_path=tmp_path / "c.zarr", dataset_name="nope"))


def _make_container(path, shape=(2, 1, 96, 96)):
    g = zarr_lite.open(path)
    rng = np.random.default_rng(1)
    for name in ["train", "test"]:
        a = g.create_dataset(name, shape=shape, dtype=np.uint8)
        a[...] = (rng.random(shape) * 255).astype(np.uint8)
        a.attrs["axis_names"] = ["s", "c", "y", "x"]
    return g


def test_dataset_crops_and_host_sampler_match_reference_order(tmp_path):
    _make_container(tmp_path / "d.zarr")
    ds = get_dataset(DatasetConfig(container_path=tmp_path / "d.zarr", dataset_name="train"), crop_size=(60, 60),
                     elastic_deform=False, control_point_spacing=64, control_point_jitter=2.0, density=0.1, kappa=10.0,
                     normalization_factor=None)
    assert ds.output_shape == (44, 44) and ds.get_num_anchors() == int(0.1 * 24 * 24) and ds.get_num_references() == 31
    np.random.seed(3)
    crop, anchors, refs = next(iter(ds))
    assert crop.shape == (1, 60, 60) and crop.dtype == np.float32 and 0 < crop.max() <= 1.0  # uint8 -> /255
    np.random.seed(3)
    a_ref, r_ref = osampler.sample_coordinates((44, 44), 10.0, 0.1, 2)
    assert np.array_equal(anchors, a_ref) and np.array_equal(refs, r_ref)
    assert anchors.dtype == np.int64  # the reference's list format
    ds.coordinate_dtype = np.dtype(np.int16)  # narrowed in the DataLoader worker: a quarter of the PCIe bytes
    np.random.seed(3)
    _, a16, r16 = next(iter(ds))
    assert a16.dtype == np.int16 and np.array_equal(a16, a_ref) and np.array_equal(r16, r_ref)


def test_elastic_augmentation(tmp_path):
    """`elastic_deform=True` (the reference default, zarr_dataset.py:122-131) deforms the crops: identity
    parameters reproduce the array (bilinear on a linear ramp is exact), a quarter turn is `np.rot90`, the
    jittered transform is deterministic under its generator, and the dataset serves normalised deformed crops."""
    import math

    from cellulus_b200.datasets.augment import elastic_crop

    yy, xx = np.mgrid[:80, :90].astype(np.float32)
    ramp = (3.0 * yy + xx)[None, None]  # (s, c, y, x)
    out = elastic_crop(ramp, 0, (24, 24), (80, 90), 8, 0.0, np.random.default_rng(4), (0.0, 0.0), (1.0, 1.0))
    d = np.diff(out[0], axis=0), np.diff(out[0], axis=1)
    assert out.shape == (1, 24, 24) and np.allclose(d[0], 3.0, atol=1e-3) and np.allclose(d[1], 1.0, atol=1e-3)
    rnd = np.random.default_rng(0).random((1, 1, 80, 90)).astype(np.float32)
    o0 = elastic_crop(rnd, 0, (21, 21), (80, 90), 8, 0.0, np.random.default_rng(5), (0.0, 0.0), (1.0, 1.0))
    o90 = elastic_crop(rnd, 0, (21, 21), (80, 90), 8, 0.0, np.random.default_rng(5), (math.pi / 2, math.pi / 2), (1.0, 1.0))
    assert np.allclose(np.rot90(o0[0], -1), o90[0], atol=1e-5)
    a = elastic_crop(rnd, 0, (32, 32), (80, 90), 8, 2.0, np.random.default_rng(6))
    b = elastic_crop(rnd, 0, (32, 32), (80, 90), 8, 2.0, np.random.default_rng(6))
    assert np.array_equal(a, b) and not np.allclose(a, elastic_crop(rnd, 0, (32, 32), (80, 90), 8, 2.0, np.random.default_rng(7)))
    vol = np.random.default_rng(1).random((1, 2, 20, 40, 40)).astype(np.float32)
    assert elastic_crop(vol, 0, (12, 24, 24), (20, 40, 40), 8, 2.0, np.random.default_rng(8)).shape == (2, 12, 24, 24)
    _make_container(tmp_path / "a.zarr")
    ds = get_dataset(DatasetConfig(container_path=tmp_path / "a.zarr", dataset_name="train"), crop_size=(60, 60),
                     elastic_deform=True, control_point_spacing=32, control_point_jitter=2.0, density=0.1, kappa=10.0,
                     normalization_factor=None)
    crop, anchors, refs = next(iter(ds))
    assert crop.shape == (1, 60, 60) and crop.dtype == np.float32 and 0 < crop.max() <= 1.0
    assert ds.pair_stream() == dict(kappa=10.0, num_anchors=int(0.1 * 24 * 24), num_references=31, extent_xyz=(44, 44))


def test_device_strings_are_resolved():
    """`device = "cuda"` (no index) is a valid config value in the reference; the kernels need a concrete index."""
    from cellulus_b200.utils.device import resolve_device

    assert resolve_device("cuda", set_current=False) == torch.device("cuda", 0)
    assert resolve_device("cuda:3", set_current=False) == torch.device("cuda", 3)
    os.environ["WORLD_SIZE"], os.environ["LOCAL_RANK"] = "4", "2"
    try:
        assert resolve_device("cuda:0", set_current=False) == torch.device("cuda", 2)  # the rank's own GPU wins
    finally:
        del os.environ["WORLD_SIZE"], os.environ["LOCAL_RANK"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        resolve_device("cpu")


@pytest.mark.parametrize("codec", [None, "zlib", "gzip"])
def test_zarr_lite_compressor_roundtrip(tmp_path, codec):
    """Chunks are written in the container format the array's compressor id names (numcodecs' GZip.decode needs a
    gzip member, Zlib.decode a zlib stream)."""
    import gzip
    import zlib

    g = zarr_lite.open(tmp_path / "z.zarr")
    x = np.random.default_rng(2).integers(0, 9, size=(2, 1, 20, 24)).astype(np.uint16)
    comp = None if codec is None else {"id": codec, "level": 1}
    a = g.create_dataset("x", shape=x.shape, dtype=np.uint16, chunks=(1, 1, 20, 24), compressor=comp)
    a[...] = x
    assert np.array_equal(zarr_lite.open(tmp_path / "z.zarr", "r")["x"][...], x)
    raw = open(tmp_path / "z.zarr" / "x" / "0.0.0.0", "rb").read()
    if codec == "gzip":
        assert raw[:2] == b"\x1f\x8b" and np.array_equal(np.frombuffer(gzip.decompress(raw), np.uint16), x[0].ravel())
    elif codec == "zlib":
        assert np.array_equal(np.frombuffer(zlib.decompress(raw), np.uint16), x[0].ravel())
    else:
        assert raw == x[0].tobytes()


def test_torch_library_ops_are_registered_with_fake_impls():
    """`torch.ops.cellulus_b200.*`: schema + fake (meta) implementations, checked without a GPU by shape propagation."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    import cellulus_b200.ops  # noqa: F401

    ns = torch.ops.cellulus_b200
    assert "Tensor offsets, Tensor anchor_coordinates" in str(ns.oce_loss_fused.default._schema)
    with FakeTensorMode():
        o = torch.empty(2, 2, 16, 20, device="cuda").contiguous(memory_format=torch.channels_last)
        a = torch.empty(2, 50, 2, dtype=torch.int64, device="cuda")
        res, grad = ns.oce_loss_fused(o, a, a, 10.0, 1e-5)
        assert res.shape == (4,) and grad.shape == o.shape and grad.dtype == torch.float32
        assert grad.is_contiguous(memory_format=torch.channels_last)
        res, grad = ns.oce_loss_sampled(torch.empty(1, 3, 8, 9, 10, device="cuda"), 3.0, 40, 5, 1, 0, 10.0, 1e-5)
        assert res.shape == (4,) and grad.shape == (1, 3, 8, 9, 10)
        assert ns.tta_aggregate(torch.empty(8, 2, 5, 6, device="cuda")).shape == (3, 5, 6)
        lab = ns.detect_volume(torch.empty(4, 5, 6, 7, device="cuda"), 3.0, 0.5, 1.0, 0)
        assert lab.shape == (5, 6, 7) and lab.dtype == torch.int32


def test_unet_stand_in_geometry_and_checkpoint_keys():
    m = get_model(in_channels=1, out_channels=2, num_fmaps=12, fmap_inc_factor=2, features_in_last_layer=64,
                  downsampling_factors=[(2, 2)], num_spatial_dims=2)
    with torch.no_grad():
        assert m(torch.zeros(1, 1, 252, 252)).shape == (1, 2, 236, 236)  # crop - 16, as zarr_dataset.py:94 assumes
        assert m(torch.zeros(1, 1, 76, 76)).shape == (1, 2, 60, 60)
    keys = set(m.state_dict())
    for k in ["backbone.l_conv.0.conv_pass.0.weight", "backbone.l_conv.1.conv_pass.6.bias",
              "backbone.r_conv.0.0.conv_pass.6.bias", "head.0.weight", "head.2.bias"]:
        assert k in keys, k
    m3 = get_model(1, 3, 4, 2, 8, [(1, 2, 2), (1, 2, 2)], 3)
    with torch.no_grad():
        assert m3(torch.zeros(1, 1, 40, 92, 92)).shape == (1, 3, 20, 52, 52)


def test_entry_points_refuse_cpu(tmp_path):
    from cellulus_b200.train import train

    _make_container(tmp_path / "e.zarr")
    cfg = ExperimentConfig(**tomllib.loads(TRAIN_TOML))
    cfg.train_config.device = "cpu"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        train(cfg)


# ---------------------------------------------------------------------------------------------------
# host-side pieces of the post-processing / evaluation mirrors (pure numpy: no device needed)
def test_evaluate_host_tail_matches_reference_golden(golden):
    from cellulus_b200.evaluate import _ranks, compute_F1

    g = golden("evaluate")
    for case in ("2d", "3d"):
        F1, TP, FP, FN = compute_F1(g[f"{case}_IoU"])
        assert np.array_equal(np.array([F1, TP, FP, FN], dtype=np.float64), g[f"{case}_scalars"][2:])
    present = np.zeros(65536, np.uint8)
    present[[0, 7, 300, 65535]] = 1
    ids, rank = _ranks(present)
    assert ids.tolist() == [7, 300, 65535] and rank[7] == 1 and rank[300] == 2 and rank[65535] == 3 and rank[0] == 0
    assert rank.dtype == np.int32 and rank.sum() == 6


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bin_edges_are_numpy_histogram_edges(dtype):
    from cellulus_b200.segment import _bin_edges

    rng = np.random.default_rng(4)
    for _ in range(20):
        values = (rng.random(500) * rng.uniform(0.1, 1e4) + rng.uniform(-50, 50)).astype(dtype)
        _, edges = np.histogram(values, 256)
        mine = _bin_edges(float(values.min()), float(values.max()), np.dtype(dtype))
        assert mine.dtype == edges.dtype and np.array_equal(mine, edges)
    # the vectorised form used for all instances at once gives the same numbers as one call per instance
    lo = rng.random(9).astype(dtype)
    hi = (lo + rng.random(9) + 0.1).astype(dtype)
    rows = np.linspace(lo, hi, 257, endpoint=True, dtype=dtype, axis=-1)
    for k in range(9):
        assert np.array_equal(rows[k], _bin_edges(float(lo[k]), float(hi[k]), np.dtype(dtype)))


def test_otsu_tail_of_the_product_equals_the_oracle():
    from cellulus_b200.detect import otsu_from_centers, otsu_from_histogram
    from oracle import otsu as ootsu

    rng = np.random.default_rng(9)
    for dtype in (np.float32, np.float64):
        img = np.concatenate([rng.normal(0.2, 0.05, 3000), rng.normal(0.8, 0.1, 2000)]).astype(dtype)
        counts, edges = np.histogram(img, 256)
        assert otsu_from_histogram(counts, edges) == ootsu.threshold_otsu(img)
    ints = np.concatenate([rng.integers(5, 60, 4000), rng.integers(140, 250, 1500)]).astype(np.uint16)
    lo, hi = int(ints.min()), int(ints.max())
    counts = np.bincount(ints.astype(np.int64) - lo, minlength=hi - lo + 1)
    assert otsu_from_centers(counts, np.arange(lo, hi + 1)) == ootsu.threshold_otsu(ints)
