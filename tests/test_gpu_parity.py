"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs, and against the committed reference outputs (tests/golden/*.npz).

Tolerances (BASELINE.json north_star): loss and gradients within 1e-5 relative (fp32);
converged mean-shift modes within 1e-3 x bandwidth; labels equal up to permutation with
ARI >= 0.999; foreground masks, histograms and size filtering bit-exact.
"""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cellulus_b200 import kernels as K  # noqa: E402
from cellulus_b200 import synthetic  # noqa: E402
from oracle import evaluate as oeval  # noqa: E402
from oracle import mean_shift as oms  # noqa: E402
from oracle import oce_loss as oloss  # noqa: E402
from oracle import otsu as ootsu  # noqa: E402
from oracle import sampler as osampler  # noqa: E402
from oracle import post_process as opost  # noqa: E402
from oracle import size_filter as osize  # noqa: E402
from oracle import tta as otta  # noqa: E402

LOSS_RTOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def ari(a, b):
    """Adjusted Rand index of two labelings (numpy, contingency-table form)."""
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    _, ai = np.unique(a, return_inverse=True)
    _, bi = np.unique(b, return_inverse=True)
    table = np.zeros((ai.max() + 1, bi.max() + 1), dtype=np.int64)
    np.add.at(table, (ai, bi), 1)
    comb = lambda x: x * (x - 1) / 2.0  # noqa: E731
    s = comb(table).sum()
    sa, sb = comb(table.sum(1)).sum(), comb(table.sum(0)).sum()
    n = comb(len(a))
    expected = sa * sb / n
    denom = 0.5 * (sa + sb) - expected
    return 1.0 if denom == 0 else (s - expected) / denom


# ----------------------------------------------------------------------------- loss slice
@pytest.mark.parametrize("case", ["2d", "3d", "2d_hot"])
def test_fused_loss_matches_reference_golden(golden, case):
    from cellulus_b200.criterions import oce_loss_fused

    g = golden("loss")
    T, w = g[f"{case}_params"][:2]
    offsets = torch.from_numpy(g[f"{case}_offsets"]).to(_dev()).requires_grad_(True)
    anchors = torch.from_numpy(g[f"{case}_anchors"].astype(np.int64)).to(_dev())
    refs = torch.from_numpy(g[f"{case}_refs"].astype(np.int64)).to(_dev())
    loss, oce, reg, raw = oce_loss_fused(offsets, anchors, refs, T, w, return_raw=True)
    loss.backward()
    ref = g[f"{case}_loss"]
    assert abs(loss.item() - ref[0]) <= LOSS_RTOL * abs(ref[0])
    assert abs(oce.item() - ref[1]) <= LOSS_RTOL * abs(ref[1])
    assert abs(reg.item() - ref[2]) <= LOSS_RTOL * abs(ref[2])
    assert raw[3].item() == 0
    assert _rel(offsets.grad.cpu().numpy(), g[f"{case}_grad"]) <= LOSS_RTOL
    # and at least as close to exact arithmetic as the fp32 reference itself
    exact = oloss.loss_step_float64(torch.from_numpy(g[f"{case}_offsets"]), anchors.cpu(), refs.cpu(), T, w)
    assert abs(loss.item() - exact[0].item()) <= LOSS_RTOL * abs(exact[0].item())
    assert _rel(offsets.grad.cpu().numpy(), exact[3].numpy()) <= LOSS_RTOL


@pytest.mark.parametrize("nd,coord_dtype", [(2, torch.int64), (2, torch.int32), (2, torch.int16), (3, torch.int64),
                                            (3, torch.int16)])
def test_fused_loss_matches_oracle_seeded(nd, coord_dtype):
    from cellulus_b200.criterions import oce_loss_fused

    crop = (140, 140) if nd == 2 else (48, 48, 48)
    kappa, density, B = (10.0, 0.1, 3) if nd == 2 else (6.0, 0.4, 2)
    out_shape = osampler.output_shape_of(crop)
    np.random.seed(0)
    pairs = [osampler.sample_coordinates(out_shape, kappa, density, nd) for _ in range(B)]
    anchors = torch.from_numpy(np.stack([p[0] for p in pairs])).long()
    refs = torch.from_numpy(np.stack([p[1] for p in pairs])).long()
    offsets = torch.from_numpy(synthetic.loss_offsets(B, nd, out_shape, seed=1))
    l_ref, o_ref, r_ref, g_ref = oloss.loss_step(offsets, anchors, refs, 10.0, 1e-5)
    off_d = offsets.to(_dev()).requires_grad_(True)
    loss, oce, reg = oce_loss_fused(off_d, anchors.to(_dev()).to(coord_dtype), refs.to(_dev()).to(coord_dtype),
                                    10.0, 1e-5)
    (2.0 * loss).backward()  # a non-unit upstream gradient exercises the device-side scale
    assert abs(loss.item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
    assert abs(oce.item() - o_ref.item()) <= LOSS_RTOL * abs(o_ref.item())
    assert abs(reg.item() - r_ref.item()) <= LOSS_RTOL * abs(r_ref.item())
    assert _rel(off_d.grad.cpu().numpy(), 2.0 * g_ref.numpy()) <= LOSS_RTOL


@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_fused_loss_planar_staged_and_in_place_agree(odt):
    """Planar 2-D offsets take one of three routes through the same entry: gathered from the kernel's own
    channels-last copy (staging scratch passed, the default), gathered in place (no scratch), and -- when the grid
    does not fit one resident wave (here: more samples than resident thread blocks) -- in place behind a separate
    zero-fill grid.  All three against the oracle; odd extents exercise the scalar transposition / zero-fill."""
    for B, out_shape, P in ((3, (61, 75), 3000), (2, (64, 96), 5000), (900, (12, 14), 64)):
        rng = np.random.default_rng(B)
        ext_xyz = np.array(out_shape[::-1])
        anchors = np.stack([rng.integers(0, ext_xyz[k], size=(B, P)) for k in range(2)], -1)
        anchors = np.repeat(anchors[:, ::16], 16, axis=1)[:, :P]
        refs = np.clip(anchors + rng.integers(-5, 6, size=anchors.shape), 0, ext_xyz - 1)
        anchors, refs = torch.from_numpy(anchors).long(), torch.from_numpy(refs).long()
        offsets = torch.from_numpy(synthetic.loss_offsets(B, 2, out_shape, seed=3)).to(odt)
        l_ref, o_ref, r_ref, g_ref = oloss.loss_step(offsets.float(), anchors, refs, 10.0, 1e-3)
        off_d, a_d, r_d = offsets.to(_dev()), anchors.to(_dev()), refs.to(_dev())
        for staged in (True, False):
            for want_grad in (True, False):
                out, grad = K.oce_loss_fwd_bwd(off_d, a_d, r_d, 10.0, 1e-3, want_grad=want_grad, staged=staged)
                out = out.cpu().numpy()
                assert abs(out[0] - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item()), (B, staged, want_grad)
                assert abs(out[2] - r_ref.item()) <= LOSS_RTOL * abs(r_ref.item())
                assert out[3] == 0
                if want_grad:
                    assert grad.is_contiguous()
                    assert _rel(grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL, (B, staged)
        # the step is repeatable on the same workspace (arrival counters are left zeroed)
        out2, grad2 = K.oce_loss_fwd_bwd(off_d, a_d, r_d, 10.0, 1e-3)
        assert _rel(grad2.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL


@pytest.mark.parametrize("nd", [2, 3])
def test_fused_loss_channels_last_layout(nd):
    """The same op on a channels-last offsets tensor (zero-copy, vector gathers): identical results, and the
    gradient comes back in the layout of its input."""
    from cellulus_b200.criterions import oce_loss_fused

    out_shape = (104, 120) if nd == 2 else (24, 28, 32)
    ext_xyz = np.array(out_shape[::-1])
    rng = np.random.default_rng(nd)
    B, P, kap = 2, 4000, 6
    anchors = np.stack([rng.integers(kap, ext_xyz[k] - kap, size=(B, P)) for k in range(nd)], -1)
    anchors = np.repeat(anchors[:, ::25], 25, axis=1)  # runs of equal anchors, like np.repeat in the sampler
    refs = anchors + rng.integers(-kap, kap + 1, size=anchors.shape)
    anchors, refs = torch.from_numpy(anchors).long(), torch.from_numpy(refs).long()
    offsets = torch.from_numpy(synthetic.loss_offsets(B, nd, out_shape, seed=2))
    l_ref, o_ref, r_ref, g_ref = oloss.loss_step(offsets, anchors, refs, 10.0, 1e-3)
    fmt = torch.channels_last if nd == 2 else torch.channels_last_3d
    off_d = offsets.to(_dev()).contiguous(memory_format=fmt).requires_grad_(True)
    loss, oce, reg = oce_loss_fused(off_d, anchors.to(_dev()), refs.to(_dev()), 10.0, 1e-3)
    loss.backward()
    assert abs(loss.item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
    assert abs(reg.item() - r_ref.item()) <= LOSS_RTOL * abs(r_ref.item())
    assert off_d.grad.is_contiguous(memory_format=fmt)
    assert _rel(off_d.grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL


def test_unfused_drop_in_matches_reference_golden(golden):
    """The reference's three-call shape: gather, gather, criterion (train.py:169-176)."""
    from cellulus_b200.criterions import get_loss
    from cellulus_b200.models import UNetModel

    g = golden("loss")
    for case in ["2d", "3d"]:
        T, w = g[f"{case}_params"][:2]
        offsets = torch.from_numpy(g[f"{case}_offsets"]).to(_dev()).requires_grad_(True)
        anchors = torch.from_numpy(g[f"{case}_anchors"].astype(np.int64)).to(_dev())
        refs = torch.from_numpy(g[f"{case}_refs"].astype(np.int64)).to(_dev())
        crit = get_loss(temperature=T, regularizer_weight=w, density=0.1, num_spatial_dims=offsets.ndim - 2,
                        device=_dev())
        ea = UNetModel.select_and_add_coordinates(offsets, anchors)
        er = UNetModel.select_and_add_coordinates(offsets, refs)
        assert np.array_equal(ea.detach().cpu().numpy(), g[f"{case}_ea"])  # gather + add is bit-exact
        assert np.array_equal(er.detach().cpu().numpy(), g[f"{case}_er"])
        loss, oce, reg = crit(ea, er)
        loss.backward()
        ref = g[f"{case}_loss"]
        assert abs(loss.item() - ref[0]) <= LOSS_RTOL * abs(ref[0])
        assert abs(oce.item() - ref[1]) <= LOSS_RTOL * abs(ref[1])
        assert _rel(offsets.grad.cpu().numpy(), g[f"{case}_grad"]) <= LOSS_RTOL


def test_fused_loss_edge_cases():
    from cellulus_b200.criterions import oce_loss_fused

    dev = _dev()
    offsets = torch.randn(2, 2, 16, 20, device=dev, requires_grad=True)
    # empty pair list -> zero loss, zero gradient
    empty = torch.zeros((2, 0, 2), dtype=torch.int64, device=dev)
    loss, oce, reg = oce_loss_fused(offsets, empty, empty, 10.0, 1e-5)
    loss.backward()
    assert loss.item() == 0.0 and offsets.grad.abs().max().item() == 0.0
    # ragged tail (P not a multiple of 32), duplicate anchors, anchor == reference (d = 0, norm grad 0)
    P = 45
    a = torch.stack([torch.randint(0, 20, (2, P)), torch.randint(0, 16, (2, P))], -1).to(dev)
    r = a.clone()
    r[:, ::2, 0] = (r[:, ::2, 0] + 3) % 20
    a[:, 5:20] = a[:, 5:6]
    r[:, 5:20] = r[:, 5:6]
    off = offsets.detach().clone().requires_grad_(True)
    loss, _, _ = oce_loss_fused(off, a, r, 3.0, 1e-2)
    loss.backward()
    l_ref, _, _, g_ref = oloss.loss_step(offsets.detach().cpu(), a.cpu(), r.cpu(), 3.0, 1e-2)
    assert abs(loss.item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
    assert _rel(off.grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL
    # out-of-range coordinates are counted, not dereferenced
    bad = a.clone()
    bad[0, 0, 0] = 20
    bad[1, 3, 1] = -17
    _, _, _, raw = oce_loss_fused(offsets.detach(), bad, r, 3.0, 1e-2, return_raw=True)
    assert raw[3].item() == 2.0
    # negative indices wrap once, as torch advanced indexing does
    neg = a.clone()
    neg[0, 1, 0] -= 20
    l_neg, _, _ = oce_loss_fused(offsets.detach(), neg, r, 3.0, 1e-2)
    l_ref2, _, _, _ = oloss.loss_step(offsets.detach().cpu(), neg.cpu(), r.cpu(), 3.0, 1e-2)
    assert abs(l_neg.item() - l_ref2.item()) <= LOSS_RTOL * abs(l_ref2.item())


def test_fused_loss_autograd_generality():
    """Differentiating `oce_loss` or `regularization_loss` on their own, a weighted mix, and a second backward
    through a retained graph all agree with the oracle (the common `loss.backward()` path is the fast one)."""
    from cellulus_b200.criterions import get_loss, oce_loss_fused
    from cellulus_b200.models import UNetModel

    dev = _dev()
    out_shape = (60, 60)
    np.random.seed(5)
    a, r = osampler.sample_coordinates(out_shape, 10.0, 0.1, 2)
    anchors, refs = torch.from_numpy(a)[None], torch.from_numpy(r)[None]
    offsets = torch.from_numpy(synthetic.loss_offsets(1, 2, out_shape, seed=4))

    def oracle_grad(wl, wo, wr):
        o = offsets.clone().requires_grad_(True)
        ea = oloss.select_and_add_coordinates(o, anchors)
        er = oloss.select_and_add_coordinates(o, refs)
        loss, oce, reg = oloss.oce_loss(ea, er, 10.0, 1e-2)
        (wl * loss + wo * oce + wr * reg).backward()
        return o.grad.numpy()

    for wl, wo, wr in [(0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.5, 2.0, -1.0), (3.0, 0.0, 0.0)]:
        o = offsets.to(dev).requires_grad_(True)
        loss, oce, reg = oce_loss_fused(o, anchors.to(dev), refs.to(dev), 10.0, 1e-2)
        (wl * loss + wo * oce + wr * reg).backward()
        assert _rel(o.grad.cpu().numpy(), oracle_grad(wl, wo, wr)) <= 2e-5, (wl, wo, wr)
    # retained graph: two backward passes accumulate twice the gradient
    o = offsets.to(dev).requires_grad_(True)
    loss, _, _ = oce_loss_fused(o, anchors.to(dev), refs.to(dev), 10.0, 1e-2)
    loss.backward(retain_graph=True)
    loss.backward()
    assert _rel(o.grad.cpu().numpy(), 2.0 * oracle_grad(1.0, 0.0, 0.0)) <= LOSS_RTOL
    # the unfused three-call shape with a weighted mix
    o = offsets.to(dev).requires_grad_(True)
    crit = get_loss(temperature=10.0, regularizer_weight=1e-2, density=0.1, num_spatial_dims=2, device=dev)
    loss, oce, reg = crit(UNetModel.select_and_add_coordinates(o, anchors.to(dev)),
                          UNetModel.select_and_add_coordinates(o, refs.to(dev)))
    (0.5 * loss + 2.0 * oce - reg).backward()
    assert _rel(o.grad.cpu().numpy(), oracle_grad(0.5, 2.0, -1.0)) <= 2e-5


def test_graphed_loss_steps_and_cycle_match_oracle():
    """`GraphedLossStep` (one step per CUDA graph) and `GraphedLossCycle` (several steps, each with its own static
    buffers, in ONE graph -- how bench.py replays the loss step) against the oracle, for explicit lists in both
    layouts and for the sampled kernel; replays are repeatable and see new data written into the static buffers."""
    from cellulus_b200.criterions import GraphedLossCycle, GraphedLossStep

    dev = _dev()
    out_shape = (60, 76)
    steps, expected = [], []
    for i, fmt in enumerate((torch.contiguous_format, torch.channels_last, torch.channels_last)):
        np.random.seed(20 + i)
        pairs = [osampler.sample_coordinates((60, 60), 10.0, 0.1, 2) for _ in range(2)]
        anchors = torch.from_numpy(np.stack([p[0] for p in pairs])).long()
        refs = torch.from_numpy(np.stack([p[1] for p in pairs])).long()
        offsets = torch.from_numpy(synthetic.loss_offsets(2, 2, out_shape, seed=30 + i))
        expected.append(oloss.loss_step(offsets, anchors, refs, 10.0, 1e-4))
        steps.append(GraphedLossStep(offsets.to(dev).contiguous(memory_format=fmt), anchors.to(dev), refs.to(dev), 10.0, 1e-4))
    cycle = GraphedLossCycle(steps)
    assert len(cycle) == 3
    for _ in range(2):
        cycle.replay()
    for (raw, grad), (l_ref, _, r_ref, g_ref) in zip(cycle.results, expected):
        assert abs(raw[0].item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
        assert abs(raw[2].item() - r_ref.item()) <= LOSS_RTOL * abs(r_ref.item())
        assert raw[3].item() == 0
        assert _rel(grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL
    for s, (l_ref, _, _, g_ref) in zip(steps, expected):
        s.replay()
        assert abs(s.loss.item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
        assert _rel(s.grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL
    # new data in a static buffer is what the next replay computes on
    new_off = torch.from_numpy(synthetic.loss_offsets(2, 2, out_shape, seed=99))
    steps[1].offsets.copy_(new_off.to(dev))
    l_new, _, _, g_new = oloss.loss_step(new_off, steps[1].anchors.cpu(), steps[1].refs.cpu(), 10.0, 1e-4)
    cycle.replay()
    assert abs(cycle.results[1][0][0].item() - l_new.item()) <= LOSS_RTOL * abs(l_new.item())
    assert _rel(cycle.results[1][1].cpu().numpy(), g_new.numpy()) <= LOSS_RTOL
    # the sampled kernel in a graph: equal to the explicit-list kernel on the lists of the same stream
    sp = dict(kappa=10.0, num_anchors=160, num_references=31, seed=7, sequence=3, extent_xyz=(60, 60))
    off = torch.from_numpy(synthetic.loss_offsets(2, 2, out_shape, seed=5)).to(dev).contiguous(memory_format=torch.channels_last)
    g = GraphedLossStep(off, None, None, 10.0, 1e-4, sampled=sp)
    g.replay()
    a, r = K.sample_pairs(2, (60, 60), 10.0, 160, 31, seed=7, sequence=3, device=dev)
    l_ref, _, _, g_ref = oloss.loss_step(off.cpu().contiguous(), a.cpu(), r.cpu(), 10.0, 1e-4)
    assert abs(g.loss.item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
    assert _rel(g.grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL


def test_fused_loss_bf16_offsets():
    from cellulus_b200.criterions import oce_loss_fused

    dev = _dev()
    np.random.seed(3)
    out_shape = (60, 60)
    a, r = osampler.sample_coordinates(out_shape, 10.0, 0.1, 2)
    anchors = torch.from_numpy(a)[None].to(dev)
    refs = torch.from_numpy(r)[None].to(dev)
    off_bf16 = torch.randn(1, 2, *out_shape, device=dev).to(torch.bfloat16).requires_grad_(True)
    loss, _, _ = oce_loss_fused(off_bf16, anchors, refs, 10.0, 1e-5)
    loss.backward()
    # oracle on the SAME bf16-rounded values, computed in fp32
    l_ref, _, _, g_ref = oloss.loss_step(off_bf16.detach().float().cpu(), anchors.cpu(), refs.cpu(), 10.0, 1e-5)
    assert abs(loss.item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
    assert off_bf16.grad.dtype == torch.bfloat16
    assert _rel(off_bf16.grad.float().cpu().numpy(), g_ref.numpy()) <= 2 ** -8  # bf16 rounding of the result


def test_loss_linearity_property_full_size():
    """Size-independent property at BASELINE config #2 size: with w = 0 the loss of a pair list is the sum of
    the losses of its halves, and the gradient likewise (no oracle needed at 5.6 M pairs)."""
    from cellulus_b200 import kernels as K

    dev = _dev()
    B, S = 8, (496, 496)
    unb = (S[0] - 20, S[1] - 20)
    na, nr = int(0.1 * unb[0] * unb[1]), 31
    anchors, refs = K.sample_pairs(B, (S[1], S[0]), 10.0, na, nr, seed=1, device=dev)
    offsets = torch.randn(B, 2, *S, device=dev)
    out_all, g_all = K.oce_loss_fwd_bwd(offsets, anchors, refs, 10.0, 0.0)
    h = anchors.shape[1] // 2
    out_a, g_a = K.oce_loss_fwd_bwd(offsets, anchors[:, :h].contiguous(), refs[:, :h].contiguous(), 10.0, 0.0)
    out_b, g_b = K.oce_loss_fwd_bwd(offsets, anchors[:, h:].contiguous(), refs[:, h:].contiguous(), 10.0, 0.0)
    assert abs(out_all[0].item() - (out_a[0].item() + out_b[0].item())) <= 1e-5 * abs(out_all[0].item())
    assert _rel(g_all.cpu().numpy(), (g_a + g_b).cpu().numpy()) <= 1e-5
    assert out_all[3].item() == 0


def test_device_sampler_distribution():
    """Same distribution as zarr_dataset.py:177-251 (different RNG stream): bounds, run structure,
    strict open ball minus origin, uniformity over the admissible offsets."""
    from cellulus_b200 import kernels as K

    for nd, ext, kappa in [(2, (236, 236), 10.0), (3, (64, 48, 40), 5.0)]:
        na, nr = 5000, 31
        a, r = K.sample_pairs(2, ext, kappa, na, nr, seed=7, device=_dev(), dtype=torch.int64)
        a, r = a.cpu().numpy(), r.cpu().numpy()
        assert a.shape == (2, na * nr, nd)
        runs = a.reshape(2, na, nr, nd)
        assert (runs == runs[:, :, :1]).all()  # np.repeat structure
        for k in range(nd):
            col = runs[:, :, 0, k]
            assert col.min() >= int(kappa) and col.max() <= ext[k] - int(kappa)
            assert col.min() == int(kappa) and col.max() == ext[k] - int(kappa)  # inclusive range is reached
        off = (r - a).reshape(-1, nd)
        assert ((off**2).sum(1) < kappa**2).all() and (np.abs(off).sum(1) > 0).all()
        # uniform over the admissible offsets: chi-square against the flat distribution
        k_int = int(kappa)
        grid = np.stack(np.meshgrid(*[np.arange(-k_int, k_int + 1)] * nd, indexing="ij"), -1).reshape(-1, nd)
        ok = ((grid**2).sum(1) < kappa**2) & (np.abs(grid).sum(1) > 0)
        n_adm = ok.sum()
        codes = ((off + k_int) * (2 * k_int + 1) ** np.arange(nd)).sum(1)
        counts = np.bincount(codes, minlength=(2 * k_int + 1) ** nd)
        counts = counts[counts > 0]
        assert len(counts) == n_adm
        expected = len(off) / n_adm
        chi2 = ((counts - expected) ** 2 / expected).sum()
        assert chi2 < n_adm + 6 * np.sqrt(2 * n_adm)
        # different seeds / sequences give different streams; same seed reproduces
        a2, _ = K.sample_pairs(2, ext, kappa, na, nr, seed=7, device=_dev())
        a3, _ = K.sample_pairs(2, ext, kappa, na, nr, seed=8, device=_dev())
        assert np.array_equal(a2.cpu().numpy(), a) and not np.array_equal(a3.cpu().numpy(), a)


@pytest.mark.parametrize("nd,ext,kappa,na,nr,dtype,seq", [
    (2, (236, 236), 10.0, 700, 31, torch.int64, 0), (2, (90, 120), 6.5, 333, 7, torch.int16, 3),
    (3, (64, 48, 40), 5.0, 500, 31, torch.int32, 1), (3, (40, 40, 40), 3.0, 64, 1, torch.int64, 2 ** 33 + 5)])
def test_device_sampler_bit_exact_to_stream_oracle(nd, ext, kappa, na, nr, dtype, seq):
    """`cb200_sample_pairs` writes exactly the pairs the numpy restatement of the device stream names
    (oracle/device_sampler.py: Philox4x32-10 checked against the Random123 known answers on the CPU side)."""
    from oracle import device_sampler as ods

    B, seed = 3, 0x1234_5678_9ABC_DEF0
    a, r = K.sample_pairs(B, ext, kappa, na, nr, seed=seed, sequence=seq, device=_dev(), dtype=dtype)
    a_ref, r_ref = ods.sample_pairs(B, ext, kappa, na, nr, seed, seq)
    assert np.array_equal(a.cpu().numpy().astype(np.int64), a_ref)
    assert np.array_equal(r.cpu().numpy().astype(np.int64), r_ref)


@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_sampled_loss_planar_staged_and_in_place_agree(odt):
    """Planar 2-D offsets through the sampled kernel: gathered from the kernel's own channels-last copy (staging
    scratch, the default) and gathered in place, against the oracle on the pairs of the same stream; odd extents
    exercise the scalar transposition, a repeat call the reset of the arrival counter."""
    dev = _dev()
    for B, out_shape, na in ((3, (61, 75), 120), (2, (64, 96), 300)):
        offsets = torch.from_numpy(synthetic.loss_offsets(B, 2, out_shape, seed=8)).to(odt)
        ext = (out_shape[1], out_shape[0])
        a, r = K.sample_pairs(B, ext, 6.0, na, 11, seed=3, sequence=2, device=dev)
        l_ref, _, r_ref, g_ref = oloss.loss_step(offsets.float(), a.cpu(), r.cpu(), 10.0, 1e-3)
        off_d = offsets.to(dev)
        for staged in (True, False, True):
            for want_grad in (True, False):
                out, grad, _ = K.oce_loss_sampled(off_d, 6.0, na, 11, 3, 2, 10.0, 1e-3, extent_xyz=ext, want_grad=want_grad,
                                                  staged=staged)
                out = out.cpu().numpy()
                assert abs(out[0] - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item()), (B, staged, want_grad)
                assert abs(out[2] - r_ref.item()) <= LOSS_RTOL * abs(r_ref.item()) and out[3] == 0
                if want_grad:
                    assert grad.is_contiguous() and _rel(grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL, (B, staged)


@pytest.mark.parametrize("nd,layout,odt", [(2, "planar", torch.float32), (2, "cl", torch.float32),
                                           (2, "cl", torch.bfloat16), (3, "planar", torch.float32),
                                           (3, "cl", torch.float32)])
def test_sampled_loss_matches_oracle_on_its_own_pairs(nd, layout, odt):
    """Fused-sampling mode: the pairs the kernel drew (dump mode) are the stream oracle's, and loss + gradient
    equal the reference arithmetic (oracle.oce_loss) on exactly those pairs, and the explicit-list kernel."""
    from cellulus_b200.criterions import oce_loss_fused, oce_loss_fused_sampled
    from oracle import device_sampler as ods

    dev = _dev()
    if nd == 2:
        S, kappa, na, nr, B = (70, 90), 10.0, 1003, 31, 3  # na not a multiple of 32, nr not of 4
    else:
        S, kappa, na, nr, B = (24, 30, 36), 4.0, 777, 13, 2
    ext = S[::-1]
    seed, seq, T, w = 99, 4, 10.0, 1e-3
    offsets = torch.from_numpy(synthetic.loss_offsets(B, nd, S, seed=3)).to(odt)
    fmt = torch.contiguous_format if layout == "planar" else (torch.channels_last if nd == 2 else torch.channels_last_3d)
    off_d = offsets.to(dev).contiguous(memory_format=fmt)
    out, grad, lists = K.oce_loss_sampled(off_d, kappa, na, nr, seed, seq, T, w, dump_dtype=torch.int64)
    a_ref, r_ref = ods.sample_pairs(B, ext, kappa, na, nr, seed, seq)
    assert np.array_equal(lists[0].cpu().numpy(), a_ref) and np.array_equal(lists[1].cpu().numpy(), r_ref)
    l_ref, o_ref, g_reg, g_ref = oloss.loss_step(offsets.float(), torch.from_numpy(a_ref), torch.from_numpy(r_ref), T, w)
    assert abs(out[0].item() - l_ref.item()) <= LOSS_RTOL * abs(l_ref.item())
    assert abs(out[1].item() - o_ref.item()) <= LOSS_RTOL * abs(o_ref.item())
    assert abs(out[2].item() - g_reg.item()) <= LOSS_RTOL * abs(g_reg.item())
    assert out[3].item() == 0
    assert _rel(grad.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL
    # without the dump (the production variant), through autograd, vs the explicit-list kernel on the same lists
    o1 = off_d.clone().requires_grad_(True)
    loss, oce, reg = oce_loss_fused_sampled(o1, kappa, na, nr, seed, seq, T, w)
    (3.0 * loss).backward()
    o2 = off_d.clone().requires_grad_(True)
    loss2, _, _ = oce_loss_fused(o2, lists[0], lists[1], T, w)
    (3.0 * loss2).backward()
    assert abs(loss.item() - loss2.item()) <= LOSS_RTOL * abs(loss2.item())
    tol = LOSS_RTOL if odt == torch.float32 else 2 ** -7
    assert _rel(o1.grad.float().cpu().numpy(), o2.grad.float().cpu().numpy()) <= tol
    assert o1.grad.dtype == odt and o1.grad.is_contiguous(memory_format=fmt)


def test_sampled_loss_edges_and_autograd():
    from cellulus_b200.criterions import oce_loss_fused_sampled
    from oracle import device_sampler as ods

    dev = _dev()
    S = (40, 52)
    offsets = torch.from_numpy(synthetic.loss_offsets(2, 2, S, seed=8))
    # nothing to draw -> zero loss, zero gradient
    for na, nr in [(0, 31), (5, 0)]:
        o = offsets.to(dev).requires_grad_(True)
        loss, _, _ = oce_loss_fused_sampled(o, 10.0, na, nr, 1, 0, 10.0, 1e-5)
        loss.backward()
        assert loss.item() == 0.0 and o.grad.abs().max().item() == 0.0
    # sampling extents that exceed the tensor (the reference's quirk Q4 on a non-square crop): pairs outside
    # the tensor are skipped and counted, the rest equals the oracle on the in-range pairs
    ext = (S[0] + 12, S[1])  # x drawn from a wider range than the tensor has
    out, grad, lists = K.oce_loss_sampled(offsets.to(dev), 10.0, 400, 31, 5, 0, 10.0, 1e-2, extent_xyz=ext,
                                          dump_dtype=torch.int32)
    a, r = lists[0].cpu().numpy().astype(np.int64), lists[1].cpu().numpy().astype(np.int64)
    a_ref, r_ref = ods.sample_pairs(2, ext, 10.0, 400, 31, 5, 0)
    assert np.array_equal(a, a_ref) and np.array_equal(r, r_ref)
    lim = np.array(S[::-1])
    inside = ((a >= 0) & (a < lim) & (r >= 0) & (r < lim)).all(-1)
    assert out[3].item() == float((~inside).sum()) and (~inside).sum() > 0
    l_tot, g_tot = 0.0, torch.zeros_like(offsets)
    for b in range(2):
        sel = inside[b]
        l, _, _, g = oloss.loss_step(offsets[b:b + 1], torch.from_numpy(a[b][sel])[None], torch.from_numpy(r[b][sel])[None],
                                     10.0, 1e-2)
        l_tot += l.item()
        g_tot[b] = g[0]
    assert abs(out[0].item() - l_tot) <= LOSS_RTOL * abs(l_tot)
    assert _rel(grad.cpu().numpy(), g_tot.numpy()) <= LOSS_RTOL
    # autograd: separate terms, a weighted mix, a second backward through a retained graph
    a_ref, r_ref = ods.sample_pairs(2, S[::-1], 10.0, 300, 31, 11, 2)

    def oracle_grad(wl, wo, wr):
        o = offsets.clone().requires_grad_(True)
        ea = oloss.select_and_add_coordinates(o, torch.from_numpy(a_ref))
        er = oloss.select_and_add_coordinates(o, torch.from_numpy(r_ref))
        loss, oce, reg = oloss.oce_loss(ea, er, 10.0, 1e-2)
        (wl * loss + wo * oce + wr * reg).backward()
        return o.grad.numpy()

    for wl, wo, wr in [(0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.5, 2.0, -1.0)]:
        o = offsets.to(dev).requires_grad_(True)
        loss, oce, reg = oce_loss_fused_sampled(o, 10.0, 300, 31, 11, 2, 10.0, 1e-2)
        (wl * loss + wo * oce + wr * reg).backward()
        assert _rel(o.grad.cpu().numpy(), oracle_grad(wl, wo, wr)) <= 2e-5, (wl, wo, wr)
    o = offsets.to(dev).requires_grad_(True)
    loss, _, _ = oce_loss_fused_sampled(o, 10.0, 300, 31, 11, 2, 10.0, 1e-2)
    loss.backward(retain_graph=True)
    loss.backward()
    assert _rel(o.grad.cpu().numpy(), 2.0 * oracle_grad(1.0, 0.0, 0.0)) <= LOSS_RTOL
    # kappa whose offset table does not fit in shared memory is refused, not mis-sampled
    with pytest.raises(Exception):
        K.oce_loss_sampled(torch.zeros(1, 3, 80, 80, 80, device=dev), 30.0, 10, 31, 1, 0, 10.0, 1e-5)


def test_loss_full_size_config1_against_oracle():
    """BASELINE configs[1] at FULL size (8 x 702 367 pairs on (8, 2, 496, 496)): both kernels against the
    float64 oracle arithmetic on the same pairs -- the explicit-list kernel on the reference-format int64
    lists, the fused-sampling kernel on the pairs it draws itself (stream oracle)."""
    from oracle import device_sampler as ods

    dev = _dev()
    B, S, kappa, T, w = 8, (496, 496), 10.0, 10.0, 1e-5
    na, nr = int(0.1 * (S[0] - 20) * (S[1] - 20)), 31
    assert (na, nr) == (22657, 31)
    a_ref, r_ref = ods.sample_pairs(B, S[::-1], kappa, na, nr, seed=2024, sequence=1)
    offsets = torch.from_numpy(synthetic.loss_offsets(B, 2, S, seed=0))
    exact = oloss.loss_step_float64(offsets, torch.from_numpy(a_ref), torch.from_numpy(r_ref), T, w)
    for layout in (torch.contiguous_format, torch.channels_last):
        off_d = offsets.to(dev).contiguous(memory_format=layout)
        out_l, g_l = K.oce_loss_fwd_bwd(off_d, torch.from_numpy(a_ref).to(dev), torch.from_numpy(r_ref).to(dev), T, w)
        out_s, g_s, _ = K.oce_loss_sampled(off_d, kappa, na, nr, 2024, 1, T, w)
        for out, g in ((out_l, g_l), (out_s, g_s)):
            for i in range(3):
                assert abs(out[i].item() - exact[i].item()) <= LOSS_RTOL * abs(exact[i].item())
            assert out[3].item() == 0
            assert _rel(g.cpu().numpy(), exact[3].numpy()) <= LOSS_RTOL


def test_torch_library_ops_opcheck_and_values():
    """The dispatcher registration (`torch.ops.cellulus_b200.*`): `torch.library.opcheck` (schema, fake tensor,
    autograd registration, AOT dispatch) and the same numbers as the eager entry points / the oracle."""
    import cellulus_b200.ops  # noqa: F401
    from cellulus_b200.criterions import oce_loss_fused
    from oracle import device_sampler as ods

    dev = _dev()
    ns = torch.ops.cellulus_b200
    S = (44, 44)  # square: the reference sampler draws column d from output_shape[d] (quirk Q4)
    np.random.seed(2)
    a, r = osampler.sample_coordinates(S, 6.0, 0.2, 2)
    anchors, refs = torch.from_numpy(a)[None].to(dev), torch.from_numpy(r)[None].to(dev)
    offsets = torch.from_numpy(synthetic.loss_offsets(1, 2, S, seed=9)).to(dev)
    o = offsets.clone().requires_grad_(True)
    torch.library.opcheck(ns.oce_loss_fused.default, (o, anchors, refs, 10.0, 1e-2))
    torch.library.opcheck(ns.oce_loss_sampled.default, (o, 6.0, 200, 11, 3, 1, 10.0, 1e-2))
    stack = torch.from_numpy(synthetic.tta_stack(8, 2, (20, 24), seed=0)).to(dev)
    torch.library.opcheck(ns.tta_aggregate.default, (stack,))
    emb, _, _ = synthetic.blob_scene((48, 56), 5, radius=6.0, seed=2)
    emb_d = torch.from_numpy(emb).to(dev)
    torch.library.opcheck(ns.detect_volume.default, (emb_d, 3.0, 0.5, 1.0, 0))
    # values and gradients: a weighted mix of all three results, against the oracle
    res, _ = ns.oce_loss_fused(o, anchors, refs, 10.0, 1e-2)
    (0.5 * res[0] + 2.0 * res[1] - res[2]).backward()
    oo = offsets.cpu().clone().requires_grad_(True)
    ea = oloss.select_and_add_coordinates(oo, anchors.cpu())
    er = oloss.select_and_add_coordinates(oo, refs.cpu())
    loss, oce, reg = oloss.oce_loss(ea, er, 10.0, 1e-2)
    (0.5 * loss + 2.0 * oce - reg).backward()
    assert abs(res[0].item() - loss.item()) <= LOSS_RTOL * abs(loss.item())
    assert _rel(o.grad.cpu().numpy(), oo.grad.numpy()) <= 2e-5
    # the dispatcher op and the eager entry point launch the same kernel
    l2, _, _ = oce_loss_fused(offsets, anchors, refs, 10.0, 1e-2)
    assert abs(l2.item() - res[0].item()) <= 1e-6 * abs(l2.item())
    res_s, g_s = ns.oce_loss_sampled(offsets, 6.0, 200, 11, 3, 1, 10.0, 1e-2)
    a_s, r_s = ods.sample_pairs(1, S[::-1], 6.0, 200, 11, 3, 1)
    l_s, _, _, g_ref = oloss.loss_step(offsets.cpu(), torch.from_numpy(a_s), torch.from_numpy(r_s), 10.0, 1e-2)
    assert abs(res_s[0].item() - l_s.item()) <= LOSS_RTOL * abs(l_s.item())
    assert _rel(g_s.cpu().numpy(), g_ref.numpy()) <= LOSS_RTOL
    assert torch.equal(ns.tta_aggregate(stack), K.tta_aggregate(stack))
    assert torch.equal(ns.detect_volume(emb_d, 3.0, 0.5, 1.0, 0), K.detect_volume(emb_d, 3.0, 0.5, 1.0, 0)[0])


# ----------------------------------------------------------------------------- TTA
@pytest.mark.parametrize("case", ["2d", "3d"])
def test_tta_matches_reference_golden(golden, case):
    from cellulus_b200.models import TTAAccumulator, tta_aggregate

    g = golden("tta")
    stack = torch.from_numpy(g[f"{case}_stack"]).to(_dev())
    out = tta_aggregate(stack).cpu().numpy()
    ref = g[f"{case}_out"]
    exact = otta.tta_aggregate_float64(torch.from_numpy(g[f"{case}_stack"])).numpy()
    scale = np.abs(exact).max()
    assert np.abs(out - ref).max() <= 1e-5 * scale
    assert np.abs(out - exact).max() <= max(np.abs(ref - exact).max() * 4, 1e-6 * scale)
    acc = TTAAccumulator(stack.shape[1], stack.shape[2:], _dev())
    for t in range(stack.shape[0]):
        acc.add(stack[t])
    assert np.abs(acc.result().cpu().numpy() - ref).max() <= 1e-5 * scale


@pytest.mark.parametrize("shape", [(32, 2, 100, 104), (6, 3, 7, 9, 11), (5, 1, 33)])
def test_tta_matches_oracle_seeded(shape):
    from cellulus_b200.models import tta_aggregate

    stack = torch.from_numpy(synthetic.tta_stack(shape[0], shape[1], shape[2:], seed=2))
    ref = otta.tta_aggregate(stack).numpy()
    out = tta_aggregate(stack.to(_dev())).cpu().numpy()
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


# ----------------------------------------------------------------------------- detect preamble
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_histogram_and_otsu_exact(dtype):
    from cellulus_b200 import kernels as K
    from cellulus_b200.detect import threshold_otsu

    emb, _, _ = synthetic.blob_scene((150, 170), 25, radius=9.0, seed=4)
    std = emb[2].astype(dtype)
    std.ravel()[::97] = std.max()  # values exactly on the last edge
    std.ravel()[5::89] = std.min()
    d = torch.from_numpy(std).to(_dev())
    mm = K.minmax(d).cpu().numpy()
    assert mm[0] == std.min() and mm[1] == std.max()
    counts_ref, edges = np.histogram(std.astype(np.float64).ravel(), 256)
    counts = K.histogram(d, torch.from_numpy(edges).to(_dev())).cpu().numpy()
    assert np.array_equal(counts, counts_ref)
    t = threshold_otsu(d)
    assert t == float(ootsu.threshold_otsu(std.astype(np.float64)))
    const = torch.full((8, 8), 0.25, device=_dev())
    assert threshold_otsu(const) == 0.25


@pytest.mark.parametrize("shape,dtype", [((70, 90), np.float32), ((9, 40, 50), np.float64), ((5, 7), np.float32)])
def test_foreground_compaction_exact(shape, dtype):
    from cellulus_b200 import kernels as K

    radius = 1.5 if min(shape) < 8 else (4.0 if min(shape) < 20 else 7.0)
    emb, _, _ = synthetic.blob_scene(shape, 6, radius=radius, seed=5, dtype=dtype)
    D = len(shape)
    thr = 0.5
    pts, pix, n, mask = K.fg_compact(torch.from_numpy(emb).to(_dev()), thr, mask_dtype=torch.uint16)
    ref_mask = emb[D].astype(np.float64) < thr
    X = oms.points_from_embedding(emb[:D], ref_mask)
    assert n == ref_mask.sum()
    assert np.array_equal(mask.cpu().numpy().astype(bool), ref_mask)  # integer foreground mask: exact
    assert np.array_equal(pts[:, :n].cpu().numpy().T, X)  # float64 points: exact
    assert np.array_equal(pix[:n].cpu().numpy(), np.flatnonzero(ref_mask.ravel()))
    # nothing / everything foreground
    _, _, n0, _ = K.fg_compact(torch.from_numpy(emb).to(_dev()), -1.0)
    _, _, n1, _ = K.fg_compact(torch.from_numpy(emb).to(_dev()), 10.0)
    assert n0 == 0 and n1 == ref_mask.size


# ----------------------------------------------------------------------------- mean-shift
MS_CASES = ["2d_all", "2d_red", "3d_red", "2d_seeds"]


def _golden_points(g, case):
    emb = g[f"{case}_emb"]
    bw, rp, thr = g[f"{case}_cfg"]
    D = emb.shape[0] - 1
    mask = emb[D].astype(np.float64) < thr
    X = oms.points_from_embedding(emb[:D], mask)
    seeds = g[f"{case}_seeds"] if f"{case}_seeds" in g.files else None
    return emb, X, mask, bw, rp, thr, seeds


def _soa(X, dev):
    n, D = X.shape
    cap = max(2, (n + 1) & ~1)
    t = torch.zeros((D, cap), dtype=torch.float64, device=dev)
    t[:, :n] = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)
    return t


@pytest.mark.parametrize("method", ["brute", "grid"])
@pytest.mark.parametrize("case", MS_CASES)
def test_mean_shift_modes_match_sklearn_golden(golden, case, method):
    """Per-seed (mode, count, iterations) against scikit-learn's own hill climb."""
    from cellulus_b200 import kernels as K

    g = golden("mean_shift")
    _, X, _, bw, _, _, seeds = _golden_points(g, case)
    Xr = X[g[f"{case}_fit_mask"]]
    sd = Xr if seeds is None else seeds.astype(np.float64)
    pts = _soa(Xr, _dev())
    modes = _soa(sd, _dev())
    if method == "brute":
        counts, iters = K.ms_brute_modes(pts, len(Xr), modes, len(sd), bw)
    else:
        lo, hi = K.bounding_box(pts, len(Xr))
        grid = K.plan_grid(lo, hi, bw)
        sorted_pts, cell_start, _ = K.grid_build(pts, len(Xr), grid)
        counts, iters = K.ms_grid_modes(sorted_pts, len(Xr), grid, cell_start, modes, len(sd), bw)
    counts = counts[: len(sd)].cpu().numpy()
    iters = iters[: len(sd)].cpu().numpy()
    m = modes[:, : len(sd)].cpu().numpy().T
    assert np.array_equal(counts, g[f"{case}_counts"])  # identical neighbour sets at the last step
    assert np.array_equal(iters, g[f"{case}_iters"])  # identical trajectories
    keep = counts > 0
    assert np.abs(m[keep] - g[f"{case}_modes"][keep]).max() <= 1e-3 * bw  # the stated bar ...
    assert np.abs(m[keep] - g[f"{case}_modes"][keep]).max() <= 1e-9 * bw  # ... met with 6 orders to spare


@pytest.mark.parametrize("method", ["brute", "grid"])
@pytest.mark.parametrize("case", MS_CASES)
def test_mean_shift_segmentation_drop_in_matches_reference(golden, case, method):
    from cellulus_b200.utils.mean_shift import mean_shift_segmentation

    g = golden("mean_shift")
    emb, _, _, bw, rp, thr, seeds = _golden_points(g, case)
    D = emb.shape[0] - 1
    emb64 = emb.astype(np.float64)
    mean_in = emb64[np.newaxis, :D].copy()
    ref_in = emb64[np.newaxis, :D].copy()
    np.random.seed(123)
    labels = mean_shift_segmentation(mean_in, emb64[D], bw, 10, rp, thr, seeds, method=method)
    assert labels.dtype == np.int32 and labels.shape == emb.shape[1:]
    ref = g[f"{case}_labels"]
    assert np.array_equal(labels > 0, ref > 0)  # foreground mask: exact
    assert ari(labels, ref) >= 0.999
    assert np.array_equal(labels, ref)  # in fact identical, numbering included
    # side effect of the reference: coordinates added in place (utils/mean_shift.py:15-32)
    np.random.seed(123)
    oms.mean_shift_segmentation(ref_in, emb64[D], bw, 10, rp, thr, seeds)
    assert np.array_equal(mean_in, ref_in)


@pytest.mark.parametrize("case", ["2d_red", "3d_red"])
def test_centres_match_sklearn_golden(golden, case):
    from cellulus_b200.utils.mean_shift import cluster_points_device

    g = golden("mean_shift")
    _, X, _, bw, _, _, seeds = _golden_points(g, case)
    Xr = X[g[f"{case}_fit_mask"]]
    pts = _soa(X, _dev())
    fit = _soa(Xr, _dev())
    for method in ["brute", "grid"]:
        centres, k, _ = cluster_points_device(pts, len(X), fit, len(Xr), bw, seeds=seeds, method=method)
        ref = g[f"{case}_centres"]
        assert k == len(ref)
        assert np.abs(centres[:, :k].cpu().numpy().T - ref).max() <= 1e-9 * bw  # same centres, same ORDER


def test_mean_shift_matches_oracle_larger_scene():
    """A scene the numpy oracle still finishes in seconds; grid and brute kernels agree with it exactly."""
    from cellulus_b200.utils.mean_shift import segment_embeddings_device

    emb, _, _ = synthetic.blob_scene((160, 160), 40, radius=8.0, seed=9)
    bw, thr, rp = 5.0, 0.5, 0.2
    D = 2
    mask = emb[D].astype(np.float64) < thr
    X = oms.points_from_embedding(emb[:D], mask)
    np.random.seed(5)
    fit_mask = np.random.rand(len(X)) < rp
    ref_labels, ref_centres = oms.segment_points(X, fit_mask, bw)
    ref = np.zeros(mask.shape, np.int32)
    ref[mask] = ref_labels + 1
    for method in ["brute", "grid"]:
        np.random.seed(5)
        labels, info = segment_embeddings_device(torch.from_numpy(emb).to(_dev()), bw, thr, rp, method=method)
        assert info["k"] == len(ref_centres)
        assert np.abs(info["centres"].cpu().numpy().T - ref_centres).max() <= 1e-3 * bw
        assert ari(labels.cpu().numpy(), ref) >= 0.999
        assert np.array_equal(labels.cpu().numpy() > 0, ref > 0)


@pytest.mark.parametrize("shape,objects,radius,bw,rp", [((160, 200), 40, 8.0, 5.0, 0.3), ((40, 96, 96), 30, 8.0, 6.0, 1.0),
                                                        ((24, 40, 40), 3, 9.0, 2.5, 1.0), ((30, 30), 1, 4.0, 40.0, 1.0)])
def test_distinct_trajectory_climb_equals_full_climb(shape, objects, radius, bw, rp):
    """`cb200_ms_grid_modes_distinct` (one evaluation per seed, then one representative per distinct unfinished mean)
    against `cb200_ms_grid_modes` (every seed to convergence): every seed it did not merge holds bit for bit the same
    mode, count and iteration count; the merged copies are exactly seeds whose full climb ends in a mode another seed
    reports; the distinct (mode, count) sets are equal, and so are the centres."""
    dev = _dev()
    emb_np, _, _ = synthetic.blob_scene(shape, objects, radius=radius, seed=13)
    emb = torch.from_numpy(emb_np).to(dev)
    pts, _, n, _ = K.fg_compact(emb, 0.5)
    fit, n_fit = K.select_points(pts, n, K.bernoulli_flags(n, rp, 3, dev)) if rp < 1.0 else (pts, n)
    lo, hi = K.bounding_box(fit, n_fit)
    grid = K.plan_grid(lo, hi, bw)
    sorted_pts, cell_start, _ = K.grid_build(fit, n_fit, grid)
    full = fit.clone()
    c_full, i_full = K.ms_grid_modes(sorted_pts, n_fit, grid, cell_start, full, n_fit, bw)
    tests_full = K.grid_modes_distance_tests()
    def distinct(m, c):
        u, cu, k = K.unique_modes(m, c, n_fit)
        rows = torch.cat([u[:, :k].view(torch.int64), cu[:k].long()[None]], 0).T.cpu().numpy()
        return np.unique(rows, axis=0)

    a, ka = K.nms_centres(full, c_full, n_fit, bw, grid)
    tests_prev = tests_full
    for merge_rounds in (1, 2, 6):  # merged once, or again after every further evaluation
        dist = fit.clone()
        c_dist, i_dist = K.ms_grid_modes_distinct(sorted_pts, n_fit, grid, cell_start, dist, n_fit, bw,
                                                  merge_rounds=merge_rounds)
        tests_dist = K.grid_modes_distance_tests()
        kept = (c_dist[:n_fit] > 0) | (c_full[:n_fit] == 0)
        assert torch.equal(c_dist[:n_fit][kept], c_full[:n_fit][kept])
        assert torch.equal(i_dist[:n_fit][kept], i_full[:n_fit][kept])
        assert torch.equal(dist[:, :n_fit][:, kept].view(torch.int64), full[:, :n_fit][:, kept].view(torch.int64))
        merged = ~kept
        assert (i_dist[:n_fit][merged] < 0).all()
        assert np.array_equal(distinct(full, c_full), distinct(dist, c_dist))  # same distinct (mode, count) pairs
        b, kb = K.nms_centres(dist, c_dist, n_fit, bw, grid)
        assert ka == kb and torch.equal(a[:, :ka], b[:, :kb])
        assert tests_dist <= tests_prev  # more merges never add work
        tests_prev = tests_dist
        if n_fit > 5000:
            assert int(merged.sum()) > n_fit // 4 and tests_dist < 0.8 * tests_full  # it actually saves work


def test_unique_modes_keeps_the_copy_the_suppression_keeps():
    """`cb200_unique_modes`: one copy per bit-identical mode with count > 0 -- the one with the highest count, then
    the lowest index -- in input order; against numpy on crafted duplicates, and the suppression gives the same
    centres in the same order with and without the pass on a real scene."""
    dev = _dev()
    rng = np.random.default_rng(0)
    base = rng.normal(size=(3, 37)) * 50
    pick = rng.integers(0, 37, size=5000)
    modes = base[:, pick].copy()
    modes[0, 100] = -0.0  # a different bit pattern than +0.0: its own mode (the suppression merges the two anyway)
    modes[0, 101] = 0.0
    counts = rng.integers(0, 4, size=5000).astype(np.int32)  # zeros are dropped
    m = torch.zeros((3, 5000), dtype=torch.float64, device=dev)
    m[:] = torch.from_numpy(modes).to(dev)
    out, c_out, n_u = K.unique_modes(m, torch.from_numpy(counts).to(dev), 5000)
    best = {}
    for i in range(5000):
        if counts[i] <= 0:
            continue
        key = modes[:, i].tobytes()
        if key not in best or counts[i] > counts[best[key]]:
            best[key] = i
    keep = sorted(best.values())
    assert n_u == len(keep)
    assert np.array_equal(out[:, :n_u].cpu().numpy().view(np.int64), modes[:, keep].view(np.int64))
    assert np.array_equal(c_out[:n_u].cpu().numpy(), counts[keep])
    # all-empty and tiny inputs
    _, _, n0 = K.unique_modes(m, torch.zeros(5000, dtype=torch.int32, device=dev), 5000)
    assert n0 == 0
    # same centres, same order, with and without the pass
    emb_np, _, _ = synthetic.blob_scene((40, 96, 96), 30, radius=8.0, seed=4)
    emb = torch.from_numpy(emb_np).to(dev)
    pts, _, n, _ = K.fg_compact(emb, 0.5)
    lo, hi = K.bounding_box(pts, n)
    grid = K.plan_grid(lo, hi, 6.0)
    sorted_pts, cell_start, _ = K.grid_build(pts, n, grid)
    seeds = pts.clone()
    cnt, _ = K.ms_grid_modes(sorted_pts, n, grid, cell_start, seeds, n, 6.0)
    a, ka = K.nms_centres(seeds, cnt, n, 6.0, grid, dedupe=True)
    b, kb = K.nms_centres(seeds, cnt, n, 6.0, grid, dedupe=False)
    assert ka == kb and ka >= 20 and torch.equal(a[:, :ka], b[:, :kb])
    _, _, n_u = K.unique_modes(seeds, cnt, n)
    assert ka <= n_u < n // 10  # orders of magnitude fewer candidates than seeds


def test_mean_shift_kats_and_edges():
    """Known answers verified on the reference: inclusive radius, orphans labelled by predict,
    ties -> lowest index, empty-window seeds dropped, empty mask -> all background."""
    from cellulus_b200 import kernels as K
    from cellulus_b200.utils.mean_shift import cluster_points_device, segment_embeddings_device

    dev = _dev()
    X = np.array([[0.0, 0.0], [3.0, 4.0], [100.0, 100.0]])
    for method in ["brute", "grid"]:
        pts = _soa(X, dev)
        seeds = np.array([[0.0, 0.0], [50.0, 50.0]])
        centres, k, info = cluster_points_device(pts, 3, pts, 3, 5.0, seeds=seeds, method=method)
        counts = info["counts"].cpu().numpy()
        assert counts[0] == 2 and counts[1] == 0  # distance exactly 5 is inside; far seed has an empty window
        assert k == 1
        assert np.allclose(centres[:, 0].cpu().numpy(), [1.5, 2.0])
        labels = torch.zeros(3, dtype=torch.int32, device=dev)
        K.assign_labels(pts, 3, centres, k, None, labels)
        assert labels.cpu().tolist() == [1, 1, 1]  # the orphan at (100, 100) is labelled too
    cen = _soa(np.array([[0.0, 0.0], [2.0, 0.0]]), dev)
    q = _soa(np.array([[1.0, 0.0], [1.5, 0.0]]), dev)
    labels = torch.zeros(2, dtype=torch.int32, device=dev)
    K.assign_labels(q, 2, cen, 2, None, labels)
    assert labels.cpu().tolist() == [1, 2]  # tie -> lowest index
    emb = np.ones((3, 12, 12), np.float32)
    labels, info = segment_embeddings_device(torch.from_numpy(emb).to(dev), 3.0, 0.5, 1.0)
    assert info["n_fg"] == 0 and labels.abs().max().item() == 0
    with pytest.raises(ValueError, match="No point was within bandwidth"):
        pts = _soa(X, dev)
        cluster_points_device(pts, 3, pts, 3, 5.0, seeds=np.array([[50.0, 50.0]]), method="grid")


@pytest.mark.parametrize("case", ["2d", "3d", "2d_fine", "2d_fail"])
def test_bin_seeding_matches_sklearn_golden(golden, case):
    """Grid-binned seeding (BASELINE configs[3]): `cb200_bin_seeds` yields scikit-learn's `get_bin_seeds` set
    (bit-identical float32 products; sklearn's order is first-seen, ours key order) and clustering with
    `bin_seeding=True` reproduces `MeanShift(bandwidth, bin_seeding=True)`: same centres in the same order, same
    labels for every foreground pixel."""
    from cellulus_b200.utils.mean_shift import segment_embeddings_device

    g = golden("bin_seeding")
    emb, bw = g[f"{case}_emb"], float(g[f"{case}_bw"])
    D = emb.shape[0] - 1
    X = oms.points_from_embedding(emb[:D].astype(np.float64), emb[D].astype(np.float64) < 0.5)
    seeds, k = K.bin_seeds(_soa(X, _dev()), len(X), bw)
    got = seeds[:, :k].cpu().numpy().T
    ref = g[f"{case}_seeds"]
    assert got.shape == ref.shape
    order = lambda a: a[np.lexsort(a.T[::-1])]  # noqa: E731
    assert np.array_equal(order(got), order(ref))
    labels, info = segment_embeddings_device(torch.from_numpy(emb).to(_dev()), bw, 0.5, 1.0, bin_seeding=True)
    centres = info["centres"].cpu().numpy().T
    assert centres.shape == g[f"{case}_centres"].shape
    assert np.abs(centres - g[f"{case}_centres"]).max() <= 1e-9 * bw
    fg = labels[torch.from_numpy(emb[D] < 0.5).to(_dev())].cpu().numpy()
    assert np.array_equal(fg - 1, g[f"{case}_labels"])


@pytest.mark.parametrize("nd", [2, 3])
def test_pruned_label_assignment_equals_brute_force(nd):
    """The grid-pruned nearest-centre search (+ brute force for orphans) is exactly `predict`:
    nearest centre, ties -> lowest index, points far from every centre still labelled."""
    from cellulus_b200 import kernels as K

    dev = _dev()
    rng = np.random.default_rng(nd)
    bw = 3.0
    centres = rng.uniform(0, 60, size=(200, nd))
    centres[10] = centres[11]  # exact duplicate: tie -> lowest index
    pts = np.concatenate([
        centres[rng.integers(0, 200, 20000)] + rng.normal(0, 1.5, size=(20000, nd)),
        rng.uniform(-500, 500, size=(500, nd)),           # orphans, some far outside the grid
        (centres[3] + centres[4])[None] / 2,               # equidistant point
        np.round(centres[:50]),                            # integer-valued points
    ])
    ref = oms.predict_labels(pts, centres) + 1
    P, Cn = _soa(pts, dev), _soa(centres, dev)
    lo, hi = K.bounding_box(Cn, len(centres))
    grid = K.plan_grid(lo, hi, bw)
    for dtype in [torch.int32, torch.uint16]:
        brute = torch.zeros(len(pts), dtype=dtype, device=dev)
        K.assign_labels(P, len(pts), Cn, len(centres), None, brute)
        pruned = torch.zeros(len(pts), dtype=dtype, device=dev)
        K.assign_labels(P, len(pts), Cn, len(centres), None, pruned, grid=grid)
        assert torch.equal(brute, pruned)
        assert np.array_equal(brute.cpu().numpy().astype(np.int64), ref)


def test_detect_full_pipeline_3d_properties():
    """BASELINE config #3 shape (128 x 256 x 256, 3-D embeddings): no oracle at this size; size-independent
    properties instead -- idempotent, every object recovered, labels constant inside an object."""
    from cellulus_b200.detect import detect_embeddings

    shape = (128, 256, 256)
    emb, centres, ids = synthetic.blob_scene(shape, 400, radius=10.0, seed=0)
    d = torch.from_numpy(emb).to(_dev())
    np.random.seed(0)
    labels, thr, mask, infos = detect_embeddings(d, bandwidth=7.0, threshold=0.5, reduction_probability=0.1,
                                                 return_info=True)
    np.random.seed(0)
    labels2, _, _ = detect_embeddings(d, bandwidth=7.0, threshold=0.5, reduction_probability=0.1)
    assert torch.equal(labels, labels2)  # deterministic
    lab = labels[0].cpu().numpy()
    assert lab.dtype == np.uint16
    assert np.array_equal(lab > 0, ids > 0)
    assert np.array_equal(mask.cpu().numpy().astype(bool), ids > 0)
    # vs the generating ground truth: balls that touch or overlap legitimately merge, so this is a sanity
    # bound, not the parity bar (parity is against the oracle / reference at oracle-sized scenes above)
    assert ari(lab[ids > 0], ids[ids > 0]) >= 0.9
    assert abs(infos[0]["k"] - len(np.unique(ids[ids > 0]))) <= 0.15 * len(centres)
    # two independent kernels (grid-hash and brute-force n-body) agree exactly at full size
    np.random.seed(0)
    labels_b, _, _, infos_b = detect_embeddings(d, bandwidth=7.0, threshold=0.5, reduction_probability=0.1,
                                                method="brute", return_info=True)
    assert infos_b[0]["method"] == "brute" and infos[0]["method"] == "grid"
    assert torch.equal(infos_b[0]["counts"], infos[0]["counts"]) and torch.equal(infos_b[0]["iters"], infos[0]["iters"])
    assert infos_b[0]["k"] == infos[0]["k"]
    assert (infos_b[0]["centres"] - infos[0]["centres"]).abs().max().item() <= 1e-9 * 7.0
    assert torch.equal(labels_b, labels)


# ----------------------------------------------------------------------------- size filter
def test_detect_full_size_config2_equals_reference_golden():
    """BASELINE configs[2] at FULL size (128 x 256 x 256, 1.48 M foreground voxels, 148 k seeds): the drop-in
    `mean_shift_segmentation` returns, label for label, what the REFERENCE returned on the same volume under the
    same numpy seed (tests/golden/config2_labels.npz, written by tests/golden/make_golden_config2.py from
    /root/reference: 220 s of scikit-learn on the authoring host)."""
    import os

    from cellulus_b200.utils.mean_shift import mean_shift_segmentation

    path = os.path.join(os.path.dirname(__file__), "golden", "config2_labels.npz")
    g = np.load(path)
    bw, rp, thr, radius, objects = (float(v) for v in g["cfg"])
    shape = tuple(int(v) for v in g["shape"])
    emb, _, _ = synthetic.blob_scene(shape, int(objects), radius=radius, seed=0)
    emb64 = emb.astype(np.float64)
    mean_in = emb64[np.newaxis, :3].copy()
    np.random.seed(0)
    labels = mean_shift_segmentation(mean_in, emb64[3], bandwidth=bw, min_size=0, reduction_probability=rp,
                                     threshold=thr, seeds=None)
    gold = g["labels"].astype(np.int32)
    assert labels.dtype == np.int32 and labels.shape == gold.shape
    assert int((labels > 0).sum()) == int(g["foreground"])
    assert np.array_equal(labels, gold)  # identical incl. numbering: ARI = 1 >= 0.999
    # the reference's in-place side effect on its first argument (utils/mean_shift.py:15-32)
    assert np.array_equal(mean_in[0, 0], emb64[0] + np.arange(shape[2])[None, None, :])


@pytest.mark.parametrize("shape", [(64, 80), (10, 30, 34), (1, 9), (300, 300)])
def test_size_filter_exact(shape):
    from cellulus_b200 import kernels as K
    from cellulus_b200.utils.misc import size_filter

    rng = np.random.default_rng(3)
    binary = rng.random(shape) < 0.4
    seg = (binary * rng.integers(1, 5, size=shape)).astype(np.uint16)
    lab, n = K.label_components(torch.from_numpy(seg.astype(np.int32)).to(_dev()))
    ref_lab = osize.label_equal_regions(seg)
    assert np.array_equal(lab.cpu().numpy(), ref_lab)  # same regions AND same raster-order numbering
    assert n.item() == ref_lab.max()
    for min_size in [0, 3, 12]:
        a, b = seg.copy(), seg.copy()
        out = size_filter(a, min_size)
        ref = osize.size_filter(b, min_size)
        assert np.array_equal(a, b)  # in-place removal set: exact
        assert np.array_equal(np.asarray(out), np.asarray(ref))


# ---------------------------------------------------------------------------------------------------
# "cell" post-processing (segment.py:41-51): thresholded Euclidean distance transforms, exact
def _label_blobs(shape, n_blobs, radius, seed):
    rng = np.random.default_rng(seed)
    seg = np.zeros(shape, np.int32)
    grids = np.indices(shape)
    for k in range(n_blobs):
        c = [rng.uniform(0, s) for s in shape]
        r = rng.uniform(0.5 * radius, radius)
        d2 = sum((g - ci) ** 2 for g, ci in zip(grids, c))
        seg[d2 < r * r] = k + 1
    return seg


@pytest.mark.parametrize("shape", [(97, 131), (5, 7), (1, 40), (23, 37, 41), (3, 50, 2)])
@pytest.mark.parametrize("radius", [0, 1, 1.5, 3, 6, 9.99, 25])
def test_edt_within_exact(shape, radius):
    rng = np.random.default_rng(hash((shape, radius)) % 2**32)
    mask = rng.random(shape) < 0.97  # sparse zeros: long distances
    mask[tuple(s // 2 for s in shape)] = False
    ref = opost.edt_within(mask, radius)
    got = K.edt_within(torch.from_numpy(mask).to(_dev()), radius).cpu().numpy().astype(bool)
    assert np.array_equal(got, ref)
    dense = rng.random(shape) < 0.5
    assert np.array_equal(K.edt_within(torch.from_numpy(dense).to(_dev()), radius).cpu().numpy().astype(bool),
                          opost.edt_within(dense, radius))


@pytest.mark.parametrize("shape", [(6, 9), (4, 5, 6)])
def test_edt_within_no_zero_element(shape):
    """scipy measures distances to a virtual element at (-1, 0, ..., 0) when the input has no zero."""
    mask = np.ones(shape, bool)
    for radius in [1, 2.5, 4, 7, 100]:
        got = K.edt_within(torch.from_numpy(mask).to(_dev()), radius).cpu().numpy().astype(bool)
        assert np.array_equal(got, opost.edt_within(mask, radius)), radius


@pytest.mark.parametrize("shape,blobs,radius", [((200, 260), 40, 14), ((40, 90, 70), 30, 10), ((31, 17), 0, 5),
                                                ((12, 12), 3, 30)])
@pytest.mark.parametrize("grow,shrink", [(3, 6), (1, 1), (5, 2), (0, 3), (4, 0)])
def test_grow_shrink_exact(shape, blobs, radius, grow, shrink):
    seg = _label_blobs(shape, blobs, radius, seed=len(shape) * 100 + blobs)
    ref = opost.grow_shrink(seg.copy(), grow, shrink)
    got = K.grow_shrink_(torch.from_numpy(seg.copy()).to(_dev()), grow, shrink).cpu().numpy()
    assert np.array_equal(got, ref)


# ---------------------------------------------------------------------------------------------------
# "nucleus" post-processing (segment.py:52-101): per-instance Otsu + hole filling, exact
def _nucleus_scene(shape, n_blobs, radius, dtype, seed):
    rng = np.random.default_rng(seed)
    seg = _label_blobs(shape, n_blobs, radius, seed)
    grids = np.indices(shape)
    raw = rng.random(shape) * 0.3
    for k in range(1, n_blobs + 1):  # a bright shell with a dim, noisy core: thresholded masks get holes
        m = seg == k
        if not m.any():
            continue
        c = [g[m].mean() for g in grids]
        d = np.sqrt(sum((g - ci) ** 2 for g, ci in zip(grids, c)))
        r = max(d[m].max(), 1.0)
        raw[m] += np.where(d[m] > 0.45 * r, 0.6, 0.0) + 0.2 * rng.random(int(m.sum()))
    if n_blobs >= 3:  # one instance of constant intensity: skimage returns that value, the mask is empty
        raw[seg == 2] = 0.5
    if np.issubdtype(np.dtype(dtype), np.integer):
        scale = 200 if dtype == np.uint8 else 40000
        return seg, (raw / raw.max() * scale).astype(dtype)
    return seg, raw.astype(dtype)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32, np.float64])
@pytest.mark.parametrize("shape,blobs,radius", [((150, 170), 25, 16), ((24, 60, 50), 14, 11), ((40, 40), 1, 30),
                                                ((30, 30), 0, 5), ((1, 64, 64), 6, 9)])
def test_nucleus_post_processing_exact(shape, blobs, radius, dtype):
    from cellulus_b200.segment import nucleus

    seg, raw = _nucleus_scene(shape, blobs, radius, dtype, seed=7 + len(shape) + blobs)
    ref = opost.nucleus(seg.copy(), raw)
    got = nucleus(seg.copy(), raw)
    assert got.dtype == seg.dtype
    assert np.array_equal(got, ref)
    if blobs > 3 and min(shape) > 1:  # the scene does exercise hole filling (a 1-voxel-thick 3-D box has no interior)
        thresholded = np.zeros_like(seg)
        for k in np.unique(seg)[1:]:
            m = seg == k
            thresholded[m & (raw > ootsu.threshold_otsu(raw[m]))] = k
        assert (ref != thresholded).sum() > 0


# ---------------------------------------------------------------------------------------------------
# golden fixtures written by the reference's own segment() / compute_pairwise_IoU (tests/golden/make_golden.py)
@pytest.mark.parametrize("case", ["cell2d", "cell2d_b", "cell3d", "nuc2d_u8", "nuc2d_u16", "nuc2d_f32", "nuc3d_f32"])
def test_post_process_golden(golden, case):
    from cellulus_b200.segment import grow_shrink, nucleus

    g = golden("post_process")
    detection = g[f"{case}_detection"]
    if case.startswith("cell"):
        grow, shrink = (int(v) for v in g[f"{case}_cfg"])
        out = grow_shrink(detection.copy(), grow, shrink)
    else:
        out = nucleus(detection.copy(), g[f"{case}_raw"])
    assert out.dtype == detection.dtype
    assert np.array_equal(out, g[f"{case}_segmentation"])


@pytest.mark.parametrize("case", ["2d", "3d"])
def test_evaluate_golden(golden, case):
    from cellulus_b200.evaluate import compute_F1, compute_pairwise_IoU

    g = golden("evaluate")
    IoU, SEG, n = compute_pairwise_IoU(g[f"{case}_prediction"], g[f"{case}_groundtruth"])
    assert np.array_equal(IoU, g[f"{case}_IoU"])  # bit-identical tables
    F1, TP, FP, FN = compute_F1(IoU)
    assert np.array_equal(np.array([SEG, n, F1, TP, FP, FN], dtype=np.float64), g[f"{case}_scalars"])
    assert compute_pairwise_IoU(g["empty_gt_prediction"], g["empty_gt_groundtruth"]) is None


def test_evaluate_against_oracle_many_ids():
    from cellulus_b200.evaluate import compute_pairwise_IoU

    rng = np.random.default_rng(11)
    gt = _label_blobs((220, 260), 60, 14, seed=3).astype(np.uint16)
    pred = np.roll(gt, (2, -3), axis=(0, 1)).copy()
    pred[pred > 0] += 1000  # ids are arbitrary 16-bit values
    pred[rng.random(pred.shape) < 0.02] = 0
    IoU, SEG, n = compute_pairwise_IoU(pred, gt)
    rIoU, rSEG, rn = oeval.compute_pairwise_IoU(pred, gt)
    assert np.array_equal(IoU, rIoU) and SEG == rSEG and n == rn


@pytest.mark.parametrize("shape", [(33, 47), (5, 9, 11), (1, 7)])
def test_evaluate_odd_sizes(shape):
    """Sizes that are not a multiple of the 16-byte vector width exercise the scalar tails."""
    from cellulus_b200.evaluate import compute_pairwise_IoU

    rng = np.random.default_rng(sum(shape))
    gt = rng.integers(0, 6, size=shape).astype(np.uint16)
    pred = rng.integers(0, 7, size=shape).astype(np.uint16) * 9
    IoU, SEG, n = compute_pairwise_IoU(pred, gt)
    rIoU, rSEG, rn = oeval.compute_pairwise_IoU(pred, gt)
    assert np.array_equal(IoU, rIoU) and SEG == rSEG and n == rn


# ---------------------------------------------------------------------------------------------------
# cb200_detect_volume: the step-by-step sequence behind one call -- identical labels
@pytest.mark.parametrize("shape,objects,radius,bw,rp", [((96, 120), 14, 8.0, 4.0, 1.0), ((96, 120), 14, 8.0, 4.0, 0.3),
                                                      ((40, 64, 72), 12, 7.0, 5.0, 0.2), ((40, 64, 72), 12, 7.0, 2.5, 1.0)])
def test_detect_volume_matches_step_sequence(shape, objects, radius, bw, rp):
    from cellulus_b200.utils.mean_shift import segment_embeddings_device

    emb, _, _ = synthetic.blob_scene(shape, objects, radius=radius, seed=5)
    d = torch.from_numpy(emb).to(_dev())
    kw = dict(bandwidth=bw, threshold=0.5, reduction_probability=rp, rng="philox", philox_seed=11, want_mask=True)
    for label_dtype in (torch.int32, torch.uint16):
        ref, ref_info = segment_embeddings_device(d, label_dtype=label_dtype, **kw)
        got, info = segment_embeddings_device(d, label_dtype=label_dtype, one_call=True, **kw)
        assert torch.equal(got, ref) and got.dtype == label_dtype
        assert torch.equal(info["mask"], ref_info["mask"])
        assert (info["n_fg"], info["n_fit"], info["k"]) == (ref_info["n_fg"], ref_info["n_fit"], ref_info["k"])
    labels, mask, centres, info = K.detect_volume(d, bw, 0.5, rp, philox_seed=11, centre_capacity=ref_info["k"] + 3)
    assert torch.equal(centres[:, : info["k"]], ref_info["centres"])


def test_detect_volume_edge_cases():
    emb, _, _ = synthetic.blob_scene((48, 56), 5, radius=6.0, seed=2)
    d = torch.from_numpy(emb).to(_dev())
    labels, mask, _, info = K.detect_volume(d, 3.0, -1.0, 1.0, want_mask=True)  # nothing below the threshold
    assert info["n_fg"] == 0 and int(labels.abs().sum()) == 0 and int(mask.sum()) == 0
    with pytest.raises(ValueError, match="0 sample"):  # a fit subset that selects nothing
        K.detect_volume(d, 3.0, 0.5, 0.0)
    d64 = d.double()
    a, _, _, _ = K.detect_volume(d64, 3.0, 0.5, 1.0)
    K.release_scratch()  # the workspace is rebuilt on demand
    b, _, _, _ = K.detect_volume(d, 3.0, 0.5, 1.0)
    assert torch.equal(a, b)  # float64 storage of float32 values: same result
    # the workspace protocol: a buffer sized for too little foreground is refused with the size to come back
    # with (CB200_ENOSPACE), nothing is allocated inside the library; the wrapper grows its buffer and retries
    import ctypes as C

    from cellulus_b200 import _cabi

    lib = _cabi.load()
    spatial = _cabi.spatial_array(d.shape[1:])
    small = lib.cb200_detect_volume_workspace_bytes(2, spatial, 16, 1.0)
    ws = torch.empty(small, dtype=torch.uint8, device=d.device)
    labels = torch.empty(d.shape[1:], dtype=torch.int32, device=d.device)
    info = _cabi.DetectInfo()
    rc = lib.cb200_detect_volume(d.data_ptr(), _cabi.F32, 2, spatial, 0.5, 3.0, 1.0, 0, 300, labels.data_ptr(), _cabi.I32,
                                 None, 0, None, 0, ws.data_ptr(), ws.numel(), 16, C.byref(info),
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == _cabi.ENOSPACE and info.n_foreground == int((d[2] < 0.5).sum()) and info.workspace_needed > small
    ws = torch.empty(info.workspace_needed, dtype=torch.uint8, device=d.device)
    rc = lib.cb200_detect_volume(d.data_ptr(), _cabi.F32, 2, spatial, 0.5, 3.0, 1.0, 0, 300, labels.data_ptr(), _cabi.I32,
                                 None, 0, None, 0, ws.data_ptr(), ws.numel(), info.n_foreground, C.byref(info),
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and torch.equal(labels, a.to(torch.int32))
    K.release_scratch()
    big, _, _ = synthetic.blob_scene((300, 300), 120, radius=9.0, seed=3)  # 40 % foreground: beyond the first guess
    c, _, _, info2 = K.detect_volume(torch.from_numpy(big).to(_dev()), 4.0, 0.5, 1.0)
    assert info2["n_fg"] > 300 * 300 // 4 and int((c > 0).sum()) == info2["n_fg"]
