"""GPU: `train()` / `infer()` end to end on a generated zarr container (BASELINE configs[0], repaired as in
SURVEY §4), the device-resident TTA loop, and the two small detect-preamble kernels against the oracle."""

import os
import tomllib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cellulus_b200 import synthetic, zarr_lite  # noqa: E402
from cellulus_b200.configs import ExperimentConfig  # noqa: E402
from oracle import otsu as ootsu  # noqa: E402


def test_centring_matches_oracle():
    from cellulus_b200 import kernels as K

    emb, _, _ = synthetic.blob_scene((90, 110), 14, radius=8.0, seed=6)
    for arr in [emb, emb.astype(np.float64)]:
        arr = arr.copy()
        arr[0, 5, :] = 0.0  # exact zeros inside the mask are excluded from the mean (detect.py:106-107)
        d = torch.from_numpy(arr).cuda()
        means, centred = K.centre_embeddings(d, 0.5)
        mask = arr[2].astype(np.float64) < 0.5
        ref = ootsu.centre_embeddings(arr.astype(np.float64), mask)
        ref_means = (arr[:2].astype(np.float64) - ref[:2]).reshape(2, -1)[:, 0]
        assert np.allclose(means.cpu().numpy(), ref_means, rtol=1e-12, atol=1e-12)
        tol = 1e-6 if arr.dtype == np.float32 else 1e-12
        assert np.abs(centred.cpu().numpy().astype(np.float64) - ref).max() <= tol
        assert np.array_equal(centred[2].cpu().numpy(), arr[2])


@pytest.mark.parametrize("shape", [(90, 110), (20, 48, 52), (6, 9)])
def test_seed_finder_matches_scipy_and_restated_skimage(shape):
    """detect.py:128-132: norm -> gaussian_filter (bit-exact vs scipy) -> peak_local_max -> flip."""
    from scipy import ndimage

    from cellulus_b200 import kernels as K
    from oracle import seeds as oseeds

    D = len(shape)
    emb, _, _ = synthetic.blob_scene(shape, 10, radius=1.5 if min(shape) < 10 else 6.0, seed=7, dtype=np.float64)
    centred = ootsu.centre_embeddings(emb, emb[D] < 0.5)
    ref = oseeds.find_seeds(centred)
    got = K.find_seeds(torch.from_numpy(centred).cuda())
    assert got.dtype == np.int64 and got.shape == ref.shape
    assert np.array_equal(got, ref)  # integer seeds: exact, same order
    # the blur itself, bit for bit
    mag = np.linalg.norm(centred[:-1], axis=0)
    from cellulus_b200._cabi import check, load, spatial_array
    import ctypes as C
    w, radius = K.gaussian_weights(2.0)
    d_in = torch.from_numpy(mag).cuda()
    out, scratch = torch.empty_like(d_in), torch.empty_like(d_in)
    check(load().cb200_gaussian_blur(d_in.data_ptr(), out.data_ptr(), scratch.data_ptr(), D, spatial_array(shape),
                                     w.ctypes.data_as(C.POINTER(C.c_double)), radius, 0,
                                     torch.cuda.current_stream().cuda_stream), "blur")
    assert np.array_equal(out.cpu().numpy(), ndimage.gaussian_filter(mag, sigma=2))


def test_mean_shift_with_integer_seeds_matches_oracle():
    """use_seeds branch end to end on the device: seeds -> mean-shift(seeds) -> labels, vs the oracle port."""
    from cellulus_b200 import kernels as K
    from cellulus_b200.detect import detect_embeddings
    from oracle import mean_shift as oms
    from oracle import seeds as oseeds

    emb, _, _ = synthetic.blob_scene((100, 120), 16, radius=8.0, seed=11, dtype=np.float64)
    thr, bw, rp = 0.5, 5.0, 0.4
    centred = ootsu.centre_embeddings(emb, emb[2] < thr)
    seeds = oseeds.find_seeds(centred)
    np.random.seed(4)
    ref = oms.mean_shift_segmentation(centred[np.newaxis, :2].copy(), centred[2], bw, 0, rp, thr, seeds)
    d = torch.from_numpy(centred).cuda()
    np.random.seed(4)
    labels, _, _ = detect_embeddings(d, bw, thr, 1, rp, seeds=K.find_seeds(d), rng="numpy", label_dtype=torch.int32)
    assert np.array_equal(labels[0].cpu().numpy(), ref)


def test_use_seeds_bandwidth_loop_reproduces_reference_side_effect():
    """`detect.py:121-144` with `use_seeds=True` and several bandwidths: the first clustering call shifts
    `embeddings_centered` in place, later bandwidths find their seeds on, and cluster, the shifted data (SURVEY
    quirk Q9).  The device loop against the oracle restatement of those lines, detection for detection -- and,
    where the reference's second bandwidth finds no point near its (displaced) seeds, the same ValueError."""
    from cellulus_b200.detect import add_coordinates, detect_with_seeds
    from oracle import seeds as oseeds

    emb, _, _ = synthetic.blob_scene((90, 110), 12, radius=8.0, seed=21, dtype=np.float64)
    thr, rp = 0.5, 0.5
    centred = ootsu.centre_embeddings(emb, emb[2] < thr)
    d = torch.from_numpy(centred).cuda()
    # (a) a bandwidth wide enough that the displaced seeds of index >= 1 still reach points
    ref_in = centred.copy()
    np.random.seed(9)
    ref = oseeds.use_seeds_detections(ref_in, thr, 160.0, 3, rp)
    # the in-place side effect itself: one coordinate grid added to the first D channels
    assert np.array_equal(add_coordinates(d).cpu().numpy(), ref_in) and not np.array_equal(ref_in, centred)
    np.random.seed(9)
    labels, mask = detect_with_seeds(d, 160.0, thr, 3, rp, label_dtype=torch.int32)
    assert np.array_equal(labels.cpu().numpy(), ref)
    assert np.array_equal(mask.cpu().numpy().astype(bool), emb[2] < thr)
    assert torch.equal(d, torch.from_numpy(centred).cuda())  # the device loop leaves its input alone
    # (b) index 0 alone is the ordinary seeded detection
    np.random.seed(9)
    ref0 = oseeds.use_seeds_detections(centred.copy(), thr, 6.0, 1, rp)
    np.random.seed(9)
    labels0, _ = detect_with_seeds(d, 6.0, thr, 1, rp, label_dtype=torch.int32)
    assert np.array_equal(labels0.cpu().numpy(), ref0) and ref0.max() > 1
    # (c) an ordinary bandwidth: the reference's second pass fails (scikit-learn: no point within bandwidth of any
    # seed) -- so does the drop-in
    with pytest.raises(ValueError, match="No point was within bandwidth"):
        np.random.seed(9)
        oseeds.use_seeds_detections(centred.copy(), thr, 6.0, 2, rp)
    with pytest.raises(ValueError, match="No point was within bandwidth"):
        np.random.seed(9)
        detect_with_seeds(d, 6.0, thr, 2, rp, label_dtype=torch.int32)


def test_pair_list_stager_narrows_on_the_host():
    """int64 host lists -> pinned int16 staging -> device: same coordinates, same loss as the int64 lists; a second
    use of the same slot waits for the first copy; extents beyond int16 are refused."""
    from cellulus_b200.criterions import oce_loss_fused
    from cellulus_b200.datasets import PairListStager
    from oracle import sampler as osampler

    np.random.seed(2)
    pairs = [osampler.sample_coordinates((60, 60), 10.0, 0.1, 2) for _ in range(2)]
    anchors = torch.from_numpy(np.stack([p[0] for p in pairs])).long()
    refs = torch.from_numpy(np.stack([p[1] for p in pairs])).long()
    stager = PairListStager(anchors.shape, "cuda:0", max_extent=60)
    for slot in (0, 1, 0):
        a16, r16 = stager.upload(anchors, refs, slot)
        assert a16.dtype == torch.int16 and a16.is_cuda
        assert torch.equal(a16.long().cpu(), anchors) and torch.equal(r16.long().cpu(), refs)
    offsets = torch.from_numpy(synthetic.loss_offsets(2, 2, (60, 60), seed=1)).cuda()
    l16, _, _ = oce_loss_fused(offsets, a16, r16, 10.0, 1e-5)
    l64, _, _ = oce_loss_fused(offsets, anchors.cuda(), refs.cuda(), 10.0, 1e-5)
    assert abs(l16.item() - l64.item()) <= 1e-6 * abs(l64.item())
    with pytest.raises(ValueError):
        PairListStager(anchors.shape, "cuda:0", max_extent=40000)


def test_salt_pepper_statistics():
    from cellulus_b200 import kernels as K

    raw = torch.rand(1, 1, 512, 512, device="cuda") * 0.4 + 0.05
    out = K.salt_pepper(raw, 0.05, 1.0, seed=3, sequence=0)
    hit = out != raw
    assert (out[hit] == 1.0).all() and torch.equal(out[~hit], raw[~hit])
    frac = hit.float().mean().item()
    assert abs(frac - 0.05) < 4 * np.sqrt(0.05 * 0.95 / raw.numel())
    assert torch.equal(out, K.salt_pepper(raw, 0.05, 1.0, seed=3, sequence=0))
    assert not torch.equal(out, K.salt_pepper(raw, 0.05, 1.0, seed=3, sequence=1))


def test_infer_mode_forward_matches_reference_formula():
    """The device TTA loop == the reference's formula applied to the very predictions it made."""
    from cellulus_b200 import kernels as K
    from cellulus_b200.models import get_model
    from oracle import tta as otta

    torch.manual_seed(0)
    model = get_model(1, 2, 4, 2, 8, [(2, 2)], 2).cuda().eval()
    model.set_infer(p_salt_pepper=0.05, num_infer_iterations=3, device=torch.device("cuda"))
    raw = torch.rand(1, 1, 60, 60, device="cuda")
    with torch.no_grad():
        out = model(raw)
        preds = []
        for i, val in enumerate([0.5] * 3 + [1.0] * 3):
            noisy = K.salt_pepper(raw, 0.05, val, seed=0, sequence=i)
            preds.append(model.head_forward(model.backbone(noisy))[0].float())
    ref = otta.tta_aggregate(torch.stack(preds).cpu())
    assert out.shape == (1, 3, 44, 44)
    assert (out[0].cpu() - ref).abs().max().item() <= 1e-5 * max(ref.abs().max().item(), 1.0)


def test_infer_mode_cuda_graph_equals_eager_loop():
    """The test-time-augmentation loop replayed from a CUDA graph (the default) against the same loop driven
    eagerly: same noise streams in the same order, call after call, for two input shapes used alternately."""
    from cellulus_b200.models import get_model

    torch.manual_seed(1)
    model = get_model(1, 2, 4, 2, 8, [(2, 2)], 2).cuda().eval()
    eager = get_model(1, 2, 4, 2, 8, [(2, 2)], 2).cuda().eval()
    eager.load_state_dict(model.state_dict())
    model.set_infer(p_salt_pepper=0.05, num_infer_iterations=2, device=torch.device("cuda"))
    eager.set_infer(p_salt_pepper=0.05, num_infer_iterations=2, device=torch.device("cuda"), cuda_graph=False)
    inputs = [torch.rand(1, 1, 60, 60, device="cuda"), torch.rand(2, 1, 52, 68, device="cuda"),
              torch.rand(1, 1, 60, 60, device="cuda"), torch.rand(2, 1, 52, 68, device="cuda")]
    outs = []
    with torch.no_grad():
        for raw in inputs:
            a, b = model(raw), eager(raw)
            # same kernels on the same data; a convolution algorithm may differ under capture: last-bit tolerance
            assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-5, atol=1e-6)
            outs.append(a)
        assert len(model._tta_graphs) == 2 and len(eager._tta_graphs) == 0
        again = model(inputs[0])  # a later call on the same data draws other noise
    assert not torch.equal(again, outs[0]) and (again[:, 2] > 0).any()
    # with gradients enabled the loop is driven eagerly (nothing is captured under autograd)
    n_graphs = len(model._tta_graphs)
    model(torch.rand(1, 1, 44, 44, device="cuda"))
    assert len(model._tta_graphs) == n_graphs


def _toml(tmp, crop):
    return f"""
experiment_name = "e2e"
object_size = 12

[model_config]
num_fmaps = 8
fmap_inc_factor = 2
checkpoint = "{tmp}/models/000002.pth"

[train_config]
batch_size = 4
crop_size = [{crop}, {crop}]
max_iterations = 3
num_workers = 0
elastic_deform = false
save_model_every = 1000
save_snapshot_every = 2
device = "cuda:0"
[train_config.train_data_config]
container_path = "{tmp}/data.zarr"
dataset_name = "train"

[inference_config]
crop_size = [{crop}, {crop}]
num_infer_iterations = 2
num_bandwidths = 2
threshold = 0.02
reduction_probability = 1.0
grow_distance = 1
shrink_distance = 2
device = "cuda:0"
[inference_config.dataset_config]
container_path = "{tmp}/data.zarr"
dataset_name = "test"
[inference_config.prediction_dataset_config]
container_path = "{tmp}/out.zarr"
dataset_name = "embeddings"
[inference_config.detection_dataset_config]
container_path = "{tmp}/out.zarr"
dataset_name = "detection"
secondary_dataset_name = "embeddings"
[inference_config.segmentation_dataset_config]
container_path = "{tmp}/out.zarr"
dataset_name = "segmentation"
secondary_dataset_name = "detection"
"""


def test_train_then_infer_end_to_end(tmp_path, monkeypatch):
    from cellulus_b200.infer import infer
    from cellulus_b200.train import train

    # The network is barely trained here (3 iterations on random crops), so the foreground it predicts is
    # arbitrary -- possibly a handful of pixels.  reduction_probability = 1.0 keeps the run free of the
    # ValueError scikit-learn (and this build) raises when a random fit subset comes out empty.
    torch.manual_seed(0)
    monkeypatch.chdir(tmp_path)
    g = zarr_lite.open(tmp_path / "data.zarr")
    rng = np.random.default_rng(0)
    truth = {}
    for name, n in [("train", 3), ("test", 2)]:
        a = g.create_dataset(name, shape=(n, 1, 120, 130), dtype=np.uint8)
        img = np.zeros((n, 1, 120, 130), np.uint8)
        truth[name] = np.zeros((n, 1, 120, 130), np.uint16)
        for s in range(n):
            _, _, ids = synthetic.blob_scene((120, 130), 12, radius=7.0, seed=int(rng.integers(1000)))
            img[s, 0] = np.where(ids > 0, 200, 20) + rng.integers(0, 20, size=ids.shape)
            truth[name][s, 0] = ids
        a[...] = img
        a.attrs["axis_names"] = ["s", "c", "y", "x"]

    cfg = ExperimentConfig(**tomllib.loads(_toml(tmp_path, 76)))
    ckpt = cfg.model_config.checkpoint
    cfg.model_config.checkpoint = None
    train(cfg)
    state = torch.load(tmp_path / "models" / "000002.pth", map_location="cpu")
    assert set(state) == {"iteration", "lowest_loss", "model_state_dict", "optim_state_dict", "logger_data"}
    assert state["iteration"] == 2 and len(state["logger_data"]["loss"]) == 3
    assert all(np.isfinite(v) for v in state["logger_data"]["loss"])
    assert os.path.exists(tmp_path / "loss.csv")
    snap = zarr_lite.open(tmp_path / "snapshots.zarr", "r")
    assert snap["2"]["prediction"].shape == (4, 2, 60, 60) and snap["2"]["raw"].attrs["axis_names"] == ["s", "c", "y", "x"]

    cfg.model_config.checkpoint = ckpt
    np.random.seed(0)
    infer(cfg)
    assert cfg.inference_config.bandwidth == 6.0 and cfg.inference_config.min_size == int(0.1 * np.pi * 144 / 4)
    out = zarr_lite.open(tmp_path / "out.zarr", "r")
    emb, det, seg = out["embeddings"], out["detection"], out["segmentation"]
    assert emb.shape == (2, 3, 120, 130) and emb.dtype == np.float64 and emb.attrs["axis_names"] == ["s", "c", "y", "x"]
    assert det.shape == (2, 2, 120, 130) and det.dtype == np.uint16 and seg.shape == det.shape and seg.dtype == np.uint16
    assert out["binary-segmentation"].shape == (2, 1, 120, 130) and out["centered-embeddings"].shape == (2, 3, 120, 130)
    e = emb[...]
    assert np.isfinite(e).all() and (e[:, 2] >= 0).all()  # the std channel is a sum of standard deviations
    mask = out["binary-segmentation"][...][:, 0].astype(bool)
    assert np.array_equal(mask, e[:, 2] < 0.02)  # foreground mask: exact
    d = det[...]
    assert np.array_equal(d[:, 0] > 0, mask) and np.array_equal(d[:, 1] > 0, mask)

    # segment(): "cell" post-processing + size filter, against the oracle on the stored detection
    from oracle import evaluate as oeval
    from oracle import post_process as opost
    from oracle import size_filter as osize

    min_size = cfg.inference_config.min_size
    for sample in range(2):
        for k in range(2):
            ref = opost.grow_shrink(d[sample, k].copy(), 1, 2)
            assert np.array_equal(seg[...][sample, k], osize.size_filter(ref, min_size).astype(np.uint16))

    # second pass over the stored detection: "nucleus" post-processing and evaluate() against ground truth
    from cellulus_b200.configs import DatasetConfig

    gt = zarr_lite.open(tmp_path / "out.zarr").create_dataset("groundtruth", shape=(2, 1, 120, 130), dtype=np.uint16)
    gt[...] = truth["test"]
    ic = cfg.inference_config
    ic.prediction_dataset_config = None
    ic.detection_dataset_config = None
    ic.post_processing = "nucleus"
    ic.segmentation_dataset_config = DatasetConfig(container_path=tmp_path / "out.zarr", dataset_name="segmentation-nucleus",
                                                   secondary_dataset_name="detection")
    ic.evaluation_dataset_config = DatasetConfig(container_path=tmp_path / "out.zarr", dataset_name="groundtruth",
                                                 secondary_dataset_name="segmentation-nucleus")
    infer(cfg)
    out = zarr_lite.open(tmp_path / "out.zarr", "r")
    raw_test = zarr_lite.open(tmp_path / "data.zarr", "r")["test"][...]
    nuc = out["segmentation-nucleus"][...]
    for sample in range(2):
        for k in range(2):
            ref = opost.nucleus(d[sample, k].copy(), raw_test[sample, 0])
            assert np.array_equal(nuc[sample, k], osize.size_filter(ref, min_size).astype(np.uint16))
    for k in range(2):
        lines = open(tmp_path / f"results_bandwidth-{k}.txt").read().splitlines()
        assert lines[0] == "file index, F1, SEG, TP, FP, FN " and lines[1].startswith("+++")
        tp = fp = fn = 0
        seg_sum, n_ids = 0.0, 0
        for sample in range(2):
            IoU, SEG, n = oeval.compute_pairwise_IoU(nuc[sample, k], truth["test"][sample, 0])
            F1, TP, FP, FN = oeval.compute_F1(IoU)
            assert lines[2 + sample] == f"{sample}, {F1:.05f}, {SEG / n:.05f}, {TP}, {FP}, {FN}"
            tp, fp, fn, seg_sum, n_ids = tp + TP, fp + FP, fn + FN, seg_sum + SEG, n_ids + n
        assert lines[5] == f"F1 for complete dataset is {2 * tp / (2 * tp + fp + fn):.05f} "
        assert lines[6] == f"SEG for complete dataset is {seg_sum / n_ids:.05f} "


@pytest.mark.parametrize("case", ["2d", "3d"])
def test_greedy_clustering_matches_reference_golden(golden, case):
    """`clustering="greedy"` (utils/greedy_cluster.py) as one cooperative kernel vs the reference's output.
    Bar: ARI >= 0.999 and identical foreground; the reference's own CPU and CUDA runs differ in the last ulp
    of `exp`, so a handful of borderline pixels may legitimately flip."""
    from cellulus_b200.utils.greedy_cluster import Cluster2d, Cluster3d
    from test_gpu_parity import ari

    g = golden("greedy")
    emb = g[f"{case}_emb"].astype(np.float64)
    bw, min_size = g[f"{case}_cfg"]
    D = emb.shape[0] - 1
    fg = emb[D] < 0.5
    if D == 2:
        cl = Cluster2d(width=emb.shape[2], height=emb.shape[1], fg_mask=fg, device="cuda:0")
    else:
        cl = Cluster3d(width=emb.shape[3], height=emb.shape[2], depth=emb.shape[1], fg_mask=fg, device="cuda:0")
    seg = cl.cluster(prediction=emb, bandwidth=bw, min_object_size=int(min_size)).numpy()
    ref = g[f"{case}_labels"]
    assert seg.dtype == np.int16 and seg.shape == ref.shape
    assert (seg[~fg] == 0).all()
    assert ari(seg, ref) >= 0.999
    assert (seg == ref).mean() >= 0.9995
    assert seg.max() == ref.max()
