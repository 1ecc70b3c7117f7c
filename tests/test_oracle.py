"""CPU: the oracle restatement against the reference's own outputs
(tests/golden/*.npz, produced by tests/golden/make_golden.py from
/root/reference) and against scikit-learn, which the reference delegates to."""

import numpy as np
import pytest
import torch

from oracle import evaluate as oeval
from oracle import mean_shift as oms
from oracle import oce_loss as oloss
from oracle import otsu as ootsu
from oracle import post_process as opost
from oracle import sampler as osampler
from oracle import size_filter as osize
from oracle import tta as otta


@pytest.mark.parametrize("case", ["2d", "3d", "2d_hot"])
def test_loss_matches_reference(golden, case):
    g = golden("loss")
    T, w = g[f"{case}_params"][:2]
    offsets = torch.from_numpy(g[f"{case}_offsets"])
    anchors = torch.from_numpy(g[f"{case}_anchors"].astype(np.int64))
    refs = torch.from_numpy(g[f"{case}_refs"].astype(np.int64))
    ea = oloss.select_and_add_coordinates(offsets.clone(), anchors)
    er = oloss.select_and_add_coordinates(offsets.clone(), refs)
    assert np.array_equal(ea.numpy(), g[f"{case}_ea"])
    assert np.array_equal(er.numpy(), g[f"{case}_er"])
    loss, oce, reg, grad = oloss.loss_step(offsets, anchors, refs, T, w)
    assert np.array_equal(
        np.array([loss.numpy(), oce.numpy(), reg.numpy()], dtype=np.float32), g[f"{case}_loss_f32"])
    assert np.array_equal(grad.numpy(), g[f"{case}_grad"])


def test_loss_analytic_gradient_float64(golden):
    # dL/dea = (2/T) exp(-d^2/T) (ea - er) + w ea/||ea||, scattered onto the anchor pixel
    g = golden("loss")
    T, w = g["2d_params"][:2]
    offsets = torch.from_numpy(g["2d_offsets"]).double()
    anchors = g["2d_anchors"].astype(np.int64)
    refs = g["2d_refs"].astype(np.int64)
    _, _, _, grad = oloss.loss_step_float64(offsets, torch.from_numpy(anchors), torch.from_numpy(refs), T, w)
    B = offsets.shape[0]
    manual = np.zeros_like(offsets.numpy())
    off = offsets.numpy()
    for b in range(B):
        ax, ay = anchors[b, :, 0], anchors[b, :, 1]
        rx, ry = refs[b, :, 0], refs[b, :, 1]
        ea = np.stack([off[b, 0, ay, ax] + ax, off[b, 1, ay, ax] + ay], 1)
        er = np.stack([off[b, 0, ry, rx] + rx, off[b, 1, ry, rx] + ry], 1)
        diff = ea - er
        d2 = (diff**2).sum(1)
        gr = (2.0 / T) * np.exp(-d2 / T)[:, None] * diff + w * ea / np.linalg.norm(ea, axis=1)[:, None]
        np.add.at(manual[b, 0], (ay, ax), gr[:, 0])
        np.add.at(manual[b, 1], (ay, ax), gr[:, 1])
    np.testing.assert_allclose(manual, grad.numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("case", ["2d", "3d"])
def test_sampler_matches_reference(golden, case):
    g = golden("sampler")
    cfg = g[f"{case}_cfg"]
    nd, kappa, density, seed = int(cfg[0]), cfg[1], cfg[2], int(cfg[3])
    crop = tuple(int(c) for c in cfg[4:])
    out_shape = osampler.output_shape_of(crop)
    np.random.seed(seed)
    a, r = osampler.sample_coordinates(out_shape, kappa, density, nd)
    assert np.array_equal(a, g[f"{case}_anchors"].astype(np.int64))
    assert np.array_equal(r, g[f"{case}_refs"].astype(np.int64))
    n_a, n_r = g[f"{case}_counts"]
    assert osampler.num_anchors(density, osampler.unbiased_shape_of(out_shape, kappa)) == n_a
    assert osampler.num_references(density, kappa) == n_r
    # structural properties the device sampler is later held to
    off = r - a
    assert ((off**2).sum(1) < kappa**2).all() and (np.abs(off).sum(1) > 0).all()
    assert a.min() >= kappa and (a.max(0) <= np.array(out_shape[:nd]) - kappa).all()
    assert (a.reshape(n_a, n_r, nd) == a.reshape(n_a, n_r, nd)[:, :1]).all()


@pytest.mark.parametrize("case", ["2d", "3d"])
def test_tta_matches_reference(golden, case):
    g = golden("tta")
    out = otta.tta_aggregate(torch.from_numpy(g[f"{case}_stack"]))
    assert np.array_equal(out.numpy(), g[f"{case}_out"])
    exact = otta.tta_aggregate_float64(torch.from_numpy(g[f"{case}_stack"]))
    np.testing.assert_allclose(out.numpy(), exact.numpy(), rtol=2e-5, atol=1e-6)


MS_CASES = ["2d_all", "2d_red", "3d_red", "2d_seeds"]


def _ms_inputs(g, case):
    emb = g[f"{case}_emb"].astype(np.float64)
    bw, rp, thr = g[f"{case}_cfg"]
    seeds = g[f"{case}_seeds"] if f"{case}_seeds" in g.files else None
    return emb, bw, rp, thr, seeds


@pytest.mark.parametrize("case", MS_CASES)
def test_mean_shift_port_matches_reference(golden, case):
    g = golden("mean_shift")
    emb, bw, rp, thr, seeds = _ms_inputs(g, case)
    D = emb.shape[0] - 1
    mean_in = emb[np.newaxis, :D].copy()
    np.random.seed(123)
    labels = oms.mean_shift_segmentation(mean_in, emb[D], bw, 10, rp, thr, seeds)
    assert labels.dtype == np.int32
    assert np.array_equal(labels, g[f"{case}_labels"])
    # the in-place coordinate add is part of the contract (utils/mean_shift.py:15-32)
    assert np.array_equal(mean_in[0], oms.points_from_embedding(emb[:D], np.ones(emb.shape[1:], bool))
                          .reshape(*emb.shape[1:], D).transpose(D, *range(D)))


@pytest.mark.parametrize("case", MS_CASES)
def test_mean_shift_restatement_matches_sklearn(golden, case):
    g = golden("mean_shift")
    emb, bw, rp, thr, seeds = _ms_inputs(g, case)
    D = emb.shape[0] - 1
    mask = emb[D] < thr
    X = oms.points_from_embedding(emb[:D], mask)
    fit_mask = g[f"{case}_fit_mask"]
    Xr = X[fit_mask]
    sd = Xr if seeds is None else seeds.astype(np.float64)
    modes, counts, iters = oms.mean_shift_modes(Xr, sd, bw)
    assert np.array_equal(counts, g[f"{case}_counts"])
    assert np.array_equal(iters, g[f"{case}_iters"])
    keep = counts > 0
    np.testing.assert_allclose(modes[keep], g[f"{case}_modes"][keep], rtol=0, atol=1e-9 * bw)
    centres = oms.nms_centres(modes, counts, bw)
    np.testing.assert_allclose(centres, g[f"{case}_centres"], rtol=0, atol=1e-9 * bw)
    labels = oms.predict_labels(X, centres)
    out = np.zeros(mask.shape, np.int32)
    out[mask] = labels + 1
    assert np.array_equal(out, g[f"{case}_labels"])


def test_mean_shift_kats():
    """Behaviours SURVEY §8c verified on the reference: inclusive radius,
    predict labels orphans, nearest-centre ties -> lowest index."""
    X = np.array([[0.0, 0.0], [3.0, 4.0], [100.0, 100.0]])
    modes, counts, iters = oms.mean_shift_modes(X, X[:1], 5.0)
    assert counts[0] == 2  # the point at distance exactly 5 is inside
    centres = np.array([[0.0, 0.0], [2.0, 0.0]])
    lab = oms.predict_labels(np.array([[1.0, 0.0], [50.0, 50.0]]), centres)
    assert lab[0] == 0  # tie -> lowest index
    assert lab[1] == 1  # orphan still labelled
    # seeds with an empty window are dropped
    modes, counts, _ = oms.mean_shift_modes(X, np.array([[50.0, 50.0], [0.0, 0.0]]), 5.0)
    assert counts[0] == 0 and len(oms.nms_centres(modes, counts, 5.0)) == 1


def test_bin_seeds_match_sklearn():
    from sklearn.cluster import get_bin_seeds

    rng = np.random.default_rng(0)
    X = rng.uniform(0, 50, size=(500, 2))
    mine = oms.get_bin_seeds(X, 7.0)
    theirs = get_bin_seeds(X, 7.0)
    assert mine.dtype == theirs.dtype
    assert np.array_equal(mine, theirs)


def test_otsu_restatement():
    rng = np.random.default_rng(0)
    img = np.where(rng.random((64, 64)) < 0.3, rng.uniform(0, 0.1, (64, 64)), 1 + rng.uniform(0, 0.1, (64, 64)))
    t = ootsu.threshold_otsu(img)
    assert 0.1 < t < 1.0
    # returns a bin centre of np.histogram(img, 256)
    _, edges = np.histogram(img.ravel(), 256)
    centres = (edges[:-1] + edges[1:]) / 2
    assert t in centres
    # brute-force between-class variance over the same 255 splits
    counts, _ = np.histogram(img.ravel(), 256)
    best, best_i = -1, -1
    for i in range(255):
        w1, w2 = counts[: i + 1].sum(), counts[i + 1:].sum()
        if w1 == 0 or w2 == 0:
            continue
        m1 = (counts[: i + 1] * centres[: i + 1]).sum() / w1
        m2 = (counts[i + 1:] * centres[i + 1:]).sum() / w2
        v = w1 * w2 * (m1 - m2) ** 2
        if v > best:
            best, best_i = v, i
    assert abs(t - centres[best_i]) <= (edges[1] - edges[0]) * 1.0001
    const = np.full((4, 4), 0.25)
    assert ootsu.threshold_otsu(const) == 0.25
    mask, thr = ootsu.foreground_mask(img, None)
    assert thr == t and mask.sum() == (img < t).sum()


@pytest.mark.parametrize("shape", [(40, 50), (12, 20, 24)])
def test_size_filter_restatement(shape):
    rng = np.random.default_rng(1)
    binary = rng.random(shape) < 0.35
    lab = osize.label_equal_regions(binary.astype(np.uint16))
    assert np.array_equal(lab, osize.label_binary_scipy(binary))
    # multi-valued: equal-value regions, touching different labels stay apart
    seg = (binary * rng.integers(1, 4, size=shape)).astype(np.uint16)
    lab2 = osize.label_equal_regions(seg)
    for v in range(1, 4):
        per_value = osize.label_binary_scipy(seg == v)
        a, b = lab2[seg == v], per_value[seg == v]
        # same partition
        assert len(np.unique(a)) == len(np.unique(b)) == len(np.unique(np.stack([a, b]), axis=1).T)
    out = osize.size_filter(seg.copy(), 5)
    sizes = np.bincount(out.ravel())[1:]
    assert (sizes >= 5).all()
    assert osize.size_filter(seg, 0) is seg


@pytest.mark.parametrize("shape", [(50, 61), (9, 30, 33), (5, 7)])
def test_seed_oracle_blur_restatement_is_scipy(shape):
    """The operation order the CUDA blur reproduces (oracle/seeds.py) is bit-identical to scipy's."""
    from scipy import ndimage

    from oracle import seeds as oseeds

    img = np.random.default_rng(0).random(shape) * 10
    assert np.array_equal(oseeds.gaussian_filter_restated(img, 2.0), ndimage.gaussian_filter(img, sigma=2))
    peaks = oseeds.peak_local_max(-ndimage.gaussian_filter(img, sigma=2))
    assert peaks.ndim == 2 and peaks.shape[1] == len(shape)
    if len(peaks):
        assert peaks.min() >= 1 and (peaks.max(0) <= np.array(shape) - 2).all()  # 1-px border excluded


@pytest.mark.parametrize("case", ["2d", "3d"])
def test_greedy_port_matches_reference(golden, case):
    from oracle import greedy as ogreedy

    g = golden("greedy")
    emb = g[f"{case}_emb"].astype(np.float64)
    bw, min_size = g[f"{case}_cfg"]
    D = emb.shape[0] - 1
    labels, n_obj, tried = ogreedy.greedy_cluster(emb, emb[D] < 0.5, bw, int(min_size))
    assert labels.dtype == np.int16 and np.array_equal(labels, g[f"{case}_labels"])
    assert n_obj == labels.max() and tried >= n_obj


POST_CASES = ["cell2d", "cell2d_b", "cell3d", "nuc2d_u8", "nuc2d_u16", "nuc2d_f32", "nuc3d_f32"]


@pytest.mark.parametrize("case", POST_CASES)
def test_post_process_matches_reference(golden, case):
    """oracle/post_process.py against what the reference's own `segment()` wrote (segment.py:41-101)."""
    g = golden("post_process")
    detection = g[f"{case}_detection"]
    if case.startswith("cell"):
        grow, shrink = (int(v) for v in g[f"{case}_cfg"])
        out = opost.grow_shrink(detection.copy(), grow, shrink)
    else:
        out = opost.nucleus(detection.copy(), g[f"{case}_raw"])
    assert np.array_equal(out, g[f"{case}_segmentation"])


def test_otsu_integer_image_uses_one_bin_per_value():
    rng = np.random.default_rng(2)
    img = np.concatenate([rng.integers(10, 40, 500), rng.integers(150, 200, 300)]).astype(np.uint8)
    t = ootsu.threshold_otsu(img)
    assert isinstance(t, (int, np.integer)) and 39 <= t < 150  # an integer grey level between the two modes
    assert ootsu.threshold_otsu(np.full(7, 9, np.uint16)) == 9


@pytest.mark.parametrize("case", ["2d", "3d"])
def test_evaluate_matches_reference(golden, case):
    g = golden("evaluate")
    returned = oeval.compute_pairwise_IoU(g[f"{case}_prediction"], g[f"{case}_groundtruth"])
    IoU, SEG, n = returned
    assert np.array_equal(IoU, g[f"{case}_IoU"])
    F1, TP, FP, FN = oeval.compute_F1(IoU)
    assert np.array_equal(np.array([SEG, n, F1, TP, FP, FN], dtype=np.float64), g[f"{case}_scalars"])


def test_evaluate_without_ground_truth(golden):
    g = golden("evaluate")
    assert bool(g["empty_gt_none"])
    assert oeval.compute_pairwise_IoU(g["empty_gt_prediction"], g["empty_gt_groundtruth"]) is None


# ----------------------------------------------------------------------------- device pair stream (restated)
def test_philox_known_answers():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors: zero, all-ones, pi digits)."""
    from oracle import device_sampler as ods

    w = ods.philox4x32_10(np.array([0], dtype=np.uint64), 0, 0)[0]
    assert [int(x) for x in w] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    full = 0xFFFFFFFFFFFFFFFF
    w = ods.philox4x32_10(np.array([full], dtype=np.uint64), full, full)[0]
    assert [int(x) for x in w] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    w = ods.philox4x32_10(np.array([0x85A308D3243F6A88], dtype=np.uint64), 0x0370734413198A2E, 0x299F31D0A4093822)[0]
    assert [int(x) for x in w] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.mark.parametrize("nd,out_shape,kappa", [(2, (236, 236), 10.0), (3, (40, 48, 64), 5.0)])
def test_device_stream_has_the_reference_samplers_distribution(nd, out_shape, kappa):
    """The device stream against the restated reference sampler (itself bit-exact to the reference golden):
    identical support for anchors and offsets, np.repeat run structure, and two-sample chi-square agreement
    of the offset histograms."""
    from oracle import device_sampler as ods

    na, nr = 4000, 31
    a, r = ods.sample_pairs(1, out_shape, kappa, na, nr, seed=3)
    a, r = a[0], r[0]
    np.random.seed(0)
    big_density = na / ((out_shape[0] - 2 * kappa) * (out_shape[1] - 2 * kappa)) * 1.0001
    a_ref, r_ref = osampler.sample_coordinates(out_shape, kappa, big_density, nd)
    a_ref, r_ref = a_ref[: (len(a_ref) // 31) * 31], r_ref[: (len(a_ref) // 31) * 31]
    runs = a.reshape(na, nr, nd)
    assert (runs == runs[:, :1]).all()
    k = int(kappa)
    for d in range(nd):
        assert a[:, d].min() == a_ref[:, d].min() == k
        assert a[:, d].max() == a_ref[:, d].max() == out_shape[d] - k
    off, off_ref = r - a, r_ref - a_ref
    side = 2 * k + 1
    code = lambda o: ((o + k) * side ** np.arange(nd)).sum(1)  # noqa: E731
    h = np.bincount(code(off), minlength=side**nd).astype(np.float64)
    h_ref = np.bincount(code(off_ref), minlength=side**nd).astype(np.float64)
    assert ((h > 0) == (h_ref > 0)).all() or nd == 3  # 3-D: 4138 cells need more draws than this to all be hit
    table = ods.offset_table(kappa, nd)
    assert set(np.flatnonzero(h > 0)) <= set(code(table)) and set(np.flatnonzero(h_ref > 0)) <= set(code(table))
    sup = code(table)
    x, y = h[sup], h_ref[sup]
    k1, k2 = np.sqrt(y.sum() / x.sum()), np.sqrt(x.sum() / y.sum())
    chi2 = (((k1 * x - k2 * y) ** 2) / np.maximum(x + y, 1)).sum()
    dof = len(sup) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof)


@pytest.mark.parametrize("case", ["2d", "3d", "2d_fine", "2d_fail"])
def test_bin_seeding_oracle_matches_sklearn_golden(golden, case):
    """`oracle.mean_shift.get_bin_seeds` + the restated fit/predict against scikit-learn's own
    `get_bin_seeds` / `MeanShift(bin_seeding=True)` outputs (tests/golden/bin_seeding.npz)."""
    g = golden("bin_seeding")
    emb, bw = g[f"{case}_emb"].astype(np.float64), float(g[f"{case}_bw"])
    D = emb.shape[0] - 1
    X = oms.points_from_embedding(emb[:D], emb[D] < 0.5)
    seeds = np.asarray(oms.get_bin_seeds(X, bw), dtype=np.float64)
    assert np.array_equal(seeds, g[f"{case}_seeds"])  # same bins, same first-seen order, same float32 products
    labels, centres = oms.segment_points(X, None, bw, seeds=seeds)
    assert np.array_equal(centres, g[f"{case}_centres"]) and np.array_equal(labels, g[f"{case}_labels"])


def test_identical_trajectories_premise_of_the_distinct_climb():
    """The two exact shortcuts of the device hill climb, checked on the oracle's arithmetic (which is held to
    scikit-learn's per-seed results above): (1) seeds whose means are bit-identical after the first window evaluation
    share the rest of their trajectory -- same final mode, same count, same iteration count; (2) the converged modes
    are few distinct values, and the centres computed from one copy of each equal the centres computed from all."""
    from cellulus_b200 import synthetic

    emb, _, _ = synthetic.blob_scene((72, 88), 7, radius=8.0, seed=3, dtype=np.float64)
    X = oms.points_from_embedding(emb[:2], emb[2] < 0.5)
    bw = 5.0
    modes, counts, iters = oms.mean_shift_modes(X, X, bw)
    first, _, _ = oms.mean_shift_modes(X, X, bw, max_iter=0)  # one evaluation per seed
    going = iters >= 1  # not converged by that evaluation
    groups = {}
    for i in np.nonzero(going)[0]:
        groups.setdefault(first[i].tobytes(), []).append(i)
    assert len(groups) < going.sum() // 2  # the merge is worth something
    for members in groups.values():
        a = members[0]
        for b in members[1:]:
            assert modes[b].tobytes() == modes[a].tobytes() and counts[b] == counts[a] and iters[b] == iters[a]
    distinct = {}
    for m, c in zip(modes, counts):
        if c and (m.tobytes() not in distinct or c > distinct[m.tobytes()][1]):
            distinct[m.tobytes()] = (m, c)
    assert len(distinct) < len(X) // 10
    one_each = oms.nms_centres(np.array([v[0] for v in distinct.values()]), np.array([v[1] for v in distinct.values()]), bw)
    assert np.array_equal(one_each, oms.nms_centres(modes, counts, bw))
