"""Generate the golden fixtures under tests/golden/ by running THE REFERENCE.

Run in the authoring container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference is pure Python; its hot-path functions are imported unmodified.
Third-party I/O libraries that are not installed here (gunpowder, zarr,
funlib.learn.torch, matplotlib, skimage) are replaced by empty stub modules in
`sys.modules` purely so the reference modules import -- none of the stubbed
symbols is executed by the functions we call:

  * `cellulus.criterions.get_loss` / `OCELoss.forward`        (criterions/oce_loss.py)
  * `UNetModel.select_and_add_coordinates` (static)            (models/unet.py:108-124)
  * `UNetModel.forward` in infer mode (TTA loop + std_mean)    (models/unet.py:73-100)
        with a tiny conv as the stubbed backbone class
  * `ZarrDataset.sample_coordinates` / `sample_offsets_within_radius`
        on an instance built without `__init__`               (datasets/zarr_dataset.py:177-251)
  * `mean_shift_segmentation`, `AnchorMeanshift`               (utils/mean_shift.py)
        loaded by file path (the package __init__ pulls matplotlib)
  * scikit-learn 1.9.0 `_mean_shift_single_seed` for per-seed (mode, count, iters)
  * `Cluster2d` / `Cluster3d`                                  (utils/greedy_cluster.py), loaded by file path
  * `segment(inference_config)`                                (segment.py:13-108) over an IN-MEMORY stand-in
        for zarr (`_MemoryStore`): the "cell" branch runs unmodified on scipy's distance transform; the
        "nucleus" branch runs with `skimage.filters.threshold_otsu` (not installed) replaced by the
        restatement in oracle/otsu.py -- bounding boxes, hole filling and write order are the reference's
  * `compute_pairwise_IoU`, `compute_F1`                       (evaluate.py:72-105)

The fixtures cannot be regenerated on the GPU box (no /root/reference there);
tests only read the committed .npz files.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from cellulus_b200 import synthetic  # noqa: E402


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _TinyBackbone(torch.nn.Module):
    """Stand-in for funlib's UNet *class* so `UNetModel.__init__` runs; a
    single same-padded conv.  The golden stores the exact predictions it makes."""

    def __init__(self, in_channels, num_fmaps_out, **kwargs):
        super().__init__()
        nd = len(kwargs["downsample_factors"][0])
        conv = torch.nn.Conv2d if nd == 2 else torch.nn.Conv3d
        self.conv = conv(in_channels, num_fmaps_out, 3, padding=1)

    def forward(self, x):
        return torch.relu(self.conv(x))


def install_stubs():
    _stub("gunpowder")
    _stub("zarr")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("funlib")
    _stub("funlib.learn")
    _stub("funlib.learn.torch")
    _stub("funlib.learn.torch.models", UNet=_TinyBackbone)


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_sampler(ZarrDataset, num_spatial_dims, crop, kappa, density):
    ds = object.__new__(ZarrDataset)
    ds.num_spatial_dims = num_spatial_dims
    ds.kappa = kappa
    ds.density = density
    ds.output_shape = tuple(int(c - 16) for c in crop)  # zarr_dataset.py:94
    ds.unbiased_shape = tuple(int(o - 2 * kappa) for o in ds.output_shape)  # :96-98
    return ds


def golden_loss(ZarrDataset, UNetModel, get_loss):
    out = {}
    cases = {
        "2d": dict(nd=2, crop=(60, 60), kappa=10.0, density=0.1, B=2, T=10.0, w=1e-5),
        "3d": dict(nd=3, crop=(40, 40, 40), kappa=5.0, density=0.5, B=2, T=10.0, w=1e-5),
        "2d_hot": dict(nd=2, crop=(48, 48), kappa=4.0, density=0.6, B=3, T=2.5, w=1e-2),
    }
    for name, c in cases.items():
        ds = make_sampler(ZarrDataset, c["nd"], c["crop"], c["kappa"], c["density"])
        np.random.seed(7)
        anchors, refs = [], []
        for _ in range(c["B"]):
            a, r = ds.sample_coordinates()
            anchors.append(a)
            refs.append(r)
        anchors = np.stack(anchors).astype(np.int64)
        refs = np.stack(refs).astype(np.int64)
        offsets = synthetic.loss_offsets(c["B"], c["nd"], ds.output_shape, seed=3)
        t_off = torch.from_numpy(offsets).clone().requires_grad_(True)
        crit = get_loss(
            temperature=c["T"], regularizer_weight=c["w"], density=c["density"],
            num_spatial_dims=c["nd"], device=torch.device("cpu"),
        )
        ea = UNetModel.select_and_add_coordinates(t_off, torch.from_numpy(anchors))
        er = UNetModel.select_and_add_coordinates(t_off, torch.from_numpy(refs))
        loss, oce, reg = crit(ea, er)
        loss.backward()
        out.update({
            f"{name}_offsets": offsets,
            f"{name}_anchors": anchors.astype(np.int16),
            f"{name}_refs": refs.astype(np.int16),
            f"{name}_params": np.array([c["T"], c["w"], c["kappa"], c["density"]], dtype=np.float64),
            f"{name}_ea": ea.detach().numpy(),
            f"{name}_er": er.detach().numpy(),
            f"{name}_loss": np.array([loss.item(), oce.item(), reg.item()], dtype=np.float64),
            f"{name}_loss_f32": np.array(
                [loss.detach().numpy(), oce.detach().numpy(), reg.detach().numpy()], dtype=np.float32),
            f"{name}_grad": t_off.grad.numpy(),
        })
        print(name, "P/sample", anchors.shape[1], "loss", loss.item())
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


def golden_sampler(ZarrDataset):
    out = {}
    for name, nd, crop, kappa, density, seed in [
        ("2d", 2, (76, 76), 10.0, 0.1, 0),
        ("3d", 3, (48, 48, 48), 6.0, 0.2, 1),
    ]:
        ds = make_sampler(ZarrDataset, nd, crop, kappa, density)
        np.random.seed(seed)
        a, r = ds.sample_coordinates()
        out[f"{name}_anchors"] = a.astype(np.int16)
        out[f"{name}_refs"] = r.astype(np.int16)
        out[f"{name}_cfg"] = np.array([nd, kappa, density, seed, *crop], dtype=np.float64)
        out[f"{name}_counts"] = np.array([ds.get_num_anchors(), ds.get_num_references()])
        print("sampler", name, a.shape, ds.get_num_anchors(), ds.get_num_references())
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)


def golden_tta(UNetModel):
    out = {}
    for name, nd, shape in [("2d", 2, (24, 28)), ("3d", 3, (6, 10, 12))]:
        torch.manual_seed(11)
        model = UNetModel(
            in_channels=1, out_channels=nd, num_fmaps=4, fmap_inc_factor=2,
            features_in_last_layer=4, downsampling_factors=[(2,) * nd], num_spatial_dims=nd,
        ).eval()
        n_iter, p = 4, 0.05
        model.set_infer(p_salt_pepper=p, num_infer_iterations=n_iter, device=torch.device("cpu"))
        raw = torch.rand(1, 1, *shape)
        with torch.no_grad():
            torch.manual_seed(5)
            ref_out = model(raw)[0]
            # replay the reference's TTA loop (models/unet.py:78-89) with the same
            # RNG stream to capture the T predictions it aggregated
            torch.manual_seed(5)
            preds = []
            for val in [0.5, 1.0]:
                for _ in range(n_iter):
                    noisy = raw.detach().clone()
                    rnd = torch.rand(*noisy.shape)
                    noisy[rnd <= p] = val
                    preds.append(model.head_forward(model.backbone(noisy))[0].detach())
            stack = torch.stack(preds, dim=0)
        out[f"{name}_stack"] = stack.numpy()
        out[f"{name}_out"] = ref_out.numpy()
        print("tta", name, stack.shape, ref_out.shape)
    np.savez_compressed(os.path.join(HERE, "tta.npz"), **out)


def golden_mean_shift(ms):
    from sklearn.cluster._mean_shift import _mean_shift_single_seed
    from sklearn.neighbors import NearestNeighbors
    import sklearn

    out = {"sklearn_version": np.array(sklearn.__version__)}
    cases = [
        # name, shape, objects, radius, bandwidth, reduction_probability, threshold, int seeds?
        ("2d_all", (64, 72), 7, 7.0, 4.0, 1.0, 0.5, False),
        ("2d_red", (96, 96), 12, 8.0, 5.0, 0.3, 0.5, False),
        ("3d_red", (20, 40, 44), 6, 6.0, 4.0, 0.25, 0.5, False),
        ("2d_seeds", (64, 64), 6, 7.0, 5.0, 0.5, 0.5, True),
    ]
    for name, shape, K, radius, bw, rp, thr, use_seeds in cases:
        emb, centres, ids = synthetic.blob_scene(shape, K, radius=radius, seed=len(name) + K)
        D = len(shape)
        emb64 = emb.astype(np.float64)
        seeds = None
        if use_seeds:
            # integer (x, y) seeds like detect.py:131-132 produces, some far from any object
            seeds = np.concatenate(
                [np.round(centres).astype(np.int64), np.array([[1, 1], [shape[1] - 2, 2]])], axis=0)
        mean_in = emb64[np.newaxis, :D].copy()
        np.random.seed(123)
        labels = ms.mean_shift_segmentation(
            mean_in, emb64[D], bandwidth=bw, min_size=10, reduction_probability=rp,
            threshold=thr, seeds=seeds)
        # same call again through AnchorMeanshift to capture the fitted centres
        mean_in2 = torch.from_numpy(emb64[np.newaxis, :D].copy())
        if D == 2:
            mean_in2[:, 1] += torch.arange(shape[0])[None, :, None]
            mean_in2[:, 0] += torch.arange(shape[1])[None, None, :]
        else:
            mean_in2[:, 2] += torch.arange(shape[0])[None, :, None, None]
            mean_in2[:, 1] += torch.arange(shape[1])[None, None, :, None]
            mean_in2[:, 0] += torch.arange(shape[2])[None, None, None, :]
        assert np.array_equal(mean_in2.numpy(), mean_in), "coordinate add mismatch"
        mask = (emb64[D] < thr)[None]
        ams = ms.AnchorMeanshift(bw, reduction_probability=rp, cluster_all=False, seeds=seeds)
        np.random.seed(123)
        labels2 = ams(mean_in2, mask=mask)[0] + 1
        assert np.array_equal(labels, labels2)
        centres_fit = ams.mean_shift.cluster_centers_
        # per-seed triples straight from scikit-learn's hill climb
        X = mean_in2[0].permute(*range(1, D + 1), 0)[torch.from_numpy(mask[0])].numpy()
        np.random.seed(123)
        fit_mask = np.random.rand(len(X)) < rp if rp < 1.0 else np.ones(len(X), bool)
        Xr = X[fit_mask]
        sd = Xr if seeds is None else np.asarray(seeds)
        nbrs = NearestNeighbors(radius=bw, n_jobs=1).fit(Xr)
        res = [_mean_shift_single_seed(s, Xr, nbrs, 300) for s in sd]
        out.update({
            f"{name}_emb": emb,
            f"{name}_cfg": np.array([bw, rp, thr], dtype=np.float64),
            f"{name}_labels": labels.astype(np.int32),
            f"{name}_centres": centres_fit,
            f"{name}_fit_mask": fit_mask,
            f"{name}_modes": np.array([r[0] for r in res], dtype=np.float64),
            f"{name}_counts": np.array([r[1] for r in res], dtype=np.int64),
            f"{name}_iters": np.array([r[2] for r in res], dtype=np.int64),
        })
        if seeds is not None:
            out[f"{name}_seeds"] = seeds
        print("ms", name, "N", len(X), "fit", len(Xr), "K", len(centres_fit), "labels", labels.max())
    np.savez_compressed(os.path.join(HERE, "mean_shift.npz"), **out)


def golden_bin_seeding():
    """BASELINE configs[3] "grid-binned" seeds: scikit-learn's own `get_bin_seeds` and
    `MeanShift(bandwidth, bin_seeding=True)` (what `AnchorMeanshift` would run with that option) on the
    foreground points of a blob scene: seeds, centres, labels of all points."""
    import sklearn
    from sklearn.cluster import MeanShift, get_bin_seeds

    out = {"sklearn_version": np.array(sklearn.__version__)}
    for name, shape, K, radius, bw in [("2d", (90, 110), 12, 8.0, 5.0), ("3d", (14, 34, 36), 5, 5.0, 3.5),
                                       ("2d_fine", (40, 44), 3, 6.0, 0.4), ("2d_fail", (30, 32), 2, 5.0, 0.01)]:
        emb, _, _ = synthetic.blob_scene(shape, K, radius=radius, seed=len(name) + K)
        D = len(shape)
        emb64 = emb.astype(np.float64)
        grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
        mask = emb64[D] < 0.5
        X = np.stack([(emb64[ch] + grids[D - 1 - ch])[mask] for ch in range(D)], axis=1)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # "Binning data failed": the 2d_fine case returns the points themselves
            seeds = get_bin_seeds(X, bw, 1)
            ms = MeanShift(bandwidth=bw, bin_seeding=True).fit(X)
        out.update({f"{name}_emb": emb, f"{name}_bw": np.array(bw), f"{name}_seeds": np.asarray(seeds, dtype=np.float64),
                    f"{name}_centres": ms.cluster_centers_, f"{name}_labels": ms.predict(X).astype(np.int32)})
        print("bin seeding", name, "N", len(X), "seeds", len(seeds), "K", len(ms.cluster_centers_))
    np.savez_compressed(os.path.join(HERE, "bin_seeding.npz"), **out)


def golden_greedy(gc):
    """`Cluster2d/3d.cluster` of the reference (utils/greedy_cluster.py, loaded by path) on small scenes."""
    out = {}
    for name, shape, K, radius, bw, min_size in [("2d", (72, 80), 9, 7.0, 3.0, 10), ("3d", (16, 40, 44), 5, 6.0, 3.0, 20)]:
        emb, _, _ = synthetic.blob_scene(shape, K, radius=radius, seed=21 + len(shape), offset_sigma=0.3)
        D = len(shape)
        pred = emb.astype(np.float64)  # what detect.py reads from the float64 `embeddings` dataset
        fg = pred[D] < 0.5
        if D == 2:
            cl = gc.Cluster2d(width=shape[1], height=shape[0], fg_mask=fg, device="cpu")
        else:
            cl = gc.Cluster3d(width=shape[2], height=shape[1], depth=shape[0], fg_mask=fg, device="cpu")
        seg = cl.cluster(prediction=pred, bandwidth=bw, min_object_size=min_size)
        out[f"{name}_emb"] = emb
        out[f"{name}_cfg"] = np.array([bw, min_size], dtype=np.float64)
        out[f"{name}_labels"] = seg.numpy().astype(np.int16)
        print("greedy", name, "fg", int(fg.sum()), "instances", int(seg.max()))
    np.savez_compressed(os.path.join(HERE, "greedy.npz"), **out)


class _MemoryArray:
    """What the reference touches of a zarr array: shape, attrs, reads that COPY, slice writes."""

    def __init__(self, data):
        self.data = data
        self.attrs = {}

    @property
    def shape(self):
        return self.data.shape

    def __getitem__(self, key):
        return np.array(self.data[key])

    def __setitem__(self, key, value):
        self.data[key] = value


class _MemoryStore:
    containers = {}

    def __init__(self, path):
        self.arrays = _MemoryStore.containers.setdefault(str(path), {})

    def __getitem__(self, name):
        return self.arrays[name]

    def create_dataset(self, name, shape, dtype, **kwargs):
        self.arrays[name] = _MemoryArray(np.zeros(shape, dtype))
        return self.arrays[name]


def _label_blobs(shape, n_blobs, radius, seed):
    rng = np.random.default_rng(seed)
    seg = np.zeros(shape, np.uint16)
    grids = np.indices(shape)
    for k in range(n_blobs):
        c = [rng.uniform(0, s) for s in shape]
        r = rng.uniform(0.5 * radius, radius)
        seg[sum((g - ci) ** 2 for g, ci in zip(grids, c)) < r * r] = k + 1
    return seg


def _shell_intensities(seg, dtype, seed):
    """Bright shell, dim noisy core under every instance: the thresholded masks have holes."""
    rng = np.random.default_rng(seed)
    grids = np.indices(seg.shape)
    raw = rng.random(seg.shape) * 0.3
    for k in np.unique(seg)[1:]:
        m = seg == k
        c = [g[m].mean() for g in grids]
        d = np.sqrt(sum((g - ci) ** 2 for g, ci in zip(grids, c)))
        raw[m] += np.where(d[m] > 0.45 * max(d[m].max(), 1.0), 0.6, 0.0) + 0.2 * rng.random(int(m.sum()))
    if np.issubdtype(np.dtype(dtype), np.integer):
        return (raw / raw.max() * (200 if dtype == np.uint8 else 40000)).astype(dtype)
    return raw.astype(dtype)


def golden_post_process():
    """The reference's `segment()` itself, fed from memory."""
    import zarr  # the stub

    zarr.open = lambda path, mode=None: _MemoryStore(path)
    from oracle import otsu as ootsu
    from oracle import size_filter as osize

    _stub("skimage")
    _stub("skimage.filters", threshold_otsu=ootsu.threshold_otsu)
    _stub("skimage.measure", label=osize.label_equal_regions)
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    from cellulus.configs.inference_config import InferenceConfig
    from cellulus.segment import segment

    out = {}
    cases = [("cell2d", (90, 110), 14, 12, "cell", np.uint8, (3, 6)), ("cell2d_b", (60, 64), 9, 9, "cell", np.uint8, (5, 2)),
             ("cell3d", (18, 44, 40), 8, 9, "cell", np.uint8, (3, 6)),
             ("nuc2d_u8", (90, 110), 14, 14, "nucleus", np.uint8, (3, 6)),
             ("nuc2d_u16", (80, 70), 10, 13, "nucleus", np.uint16, (3, 6)),
             ("nuc2d_f32", (80, 70), 10, 13, "nucleus", np.float32, (3, 6)),
             ("nuc3d_f32", (16, 40, 36), 6, 9, "nucleus", np.float32, (3, 6))]
    for name, shape, blobs, radius, mode, raw_dtype, (grow, shrink) in cases:
        nd = len(shape)
        axes = ["s", "c"] + ["z", "y", "x"][-nd:]
        seg = _label_blobs(shape, blobs, radius, seed=len(name) * 7 + blobs)
        raw = _shell_intensities(seg, raw_dtype, seed=blobs)
        store = _MemoryStore(name + ".zarr")
        store.arrays["raw"] = _MemoryArray(raw[None, None])
        store.arrays["raw"].attrs["axis_names"] = axes
        store.arrays["detection"] = _MemoryArray(seg[None, None].copy())
        cfg = InferenceConfig(
            dataset_config={"container_path": name + ".zarr", "dataset_name": "raw"},
            segmentation_dataset_config={"container_path": name + ".zarr", "dataset_name": "segmentation",
                                         "secondary_dataset_name": "detection"},
            post_processing=mode, grow_distance=grow, shrink_distance=shrink, min_size=0, num_bandwidths=1)
        segment(cfg)
        result = store.arrays["segmentation"].data[0, 0]
        out[f"{name}_detection"] = seg
        out[f"{name}_raw"] = raw
        out[f"{name}_cfg"] = np.array([grow, shrink], dtype=np.int64)
        out[f"{name}_segmentation"] = result.copy()
        print("segment", name, mode, "instances in", int(seg.max()), "labelled px in/out", int((seg > 0).sum()),
              int((result > 0).sum()))
    np.savez_compressed(os.path.join(HERE, "post_process.npz"), **out)


def golden_evaluate():
    """`compute_pairwise_IoU` + `compute_F1` of the reference (evaluate.py:72-105)."""
    from cellulus.evaluate import compute_F1, compute_pairwise_IoU

    out = {}
    for name, shape, blobs, radius in [("2d", (96, 120), 12, 13), ("3d", (14, 40, 44), 7, 9), ("empty_gt", (20, 20), 0, 5)]:
        gt = _label_blobs(shape, blobs, radius, seed=31 + blobs)
        rng = np.random.default_rng(5 + blobs)
        pred = np.zeros_like(gt)
        for k in np.unique(gt)[1:]:  # shifted / partly merged / dropped copies of the ground truth
            if rng.random() < 0.15:
                continue
            shift = tuple(int(v) for v in rng.integers(-3, 4, size=gt.ndim))
            m = np.roll(gt == k, shift, axis=tuple(range(gt.ndim)))
            pred[m] = k + 100 if rng.random() < 0.8 else 100
        out[f"{name}_groundtruth"] = gt
        out[f"{name}_prediction"] = pred
        returned = compute_pairwise_IoU(pred, gt)
        out[f"{name}_none"] = np.array(returned is None)
        if returned is not None:
            IoU, SEG, n = returned
            F1, TP, FP, FN = compute_F1(IoU)
            out[f"{name}_IoU"] = IoU
            out[f"{name}_scalars"] = np.array([SEG, n, F1, TP, FP, FN], dtype=np.float64)
            print("evaluate", name, "IoU table", IoU.shape, "SEG", SEG / n, "F1", F1, TP, FP, FN)
    np.savez_compressed(os.path.join(HERE, "evaluate.npz"), **out)


def main():
    install_stubs()
    from cellulus.criterions import get_loss
    from cellulus.models.unet import UNetModel
    from cellulus.datasets.zarr_dataset import ZarrDataset

    ms = load_by_path("ref_mean_shift", os.path.join(REF, "cellulus/utils/mean_shift.py"))
    golden_loss(ZarrDataset, UNetModel, get_loss)
    golden_sampler(ZarrDataset)
    golden_tta(UNetModel)
    golden_mean_shift(ms)
    golden_bin_seeding()
    golden_greedy(load_by_path("ref_greedy_cluster", os.path.join(REF, "cellulus/utils/greedy_cluster.py")))
    golden_post_process()
    golden_evaluate()


if __name__ == "__main__":
    main()
