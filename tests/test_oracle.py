"""Pin the oracle (CPU restatement) -- runs without a GPU.

Anchors, strongest first: (1) the reference's own known-answer NMS vectors
(tests/golden/nms_reference_tests.npz, captured from reference tests/test_nms.py by make_golden.py);
(2) outputs of the reference's own CPU kernels compiled in place (oracle/_ref) -- committed fixtures
plus, when oracle/_ref is present, live comparison; (3) torchvision CPU ops for the ops the reference
never implemented on CPU.
"""
import os

import numpy as np
import pytest
import torch
import torchvision

import _inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_nms_reference_known_answers(oracle_mod):
    g = np.load(os.path.join(GOLD, "nms_reference_tests.npz"))
    assert int(g["n"]) == 6
    for i in range(6):
        keep = oracle_mod.nms(torch.from_numpy(g["boxes%d" % i]), torch.from_numpy(g["scores%d" % i]), float(g["thr%d" % i]))
        np.testing.assert_array_equal(keep.numpy(), g["keep%d" % i])
    # the literal vector of reference tests/test_nms.py:52-53
    exp = [[1, 3], [1, 3], [1, 3], [1, 2, 3, 4], [0, 1, 2, 3, 4]]
    for i, e in enumerate(exp):
        assert g["keep%d" % i].tolist() == e
    assert len(g["keep5"]) == 26


def test_nms_vs_reference_cpu_fixture(oracle_mod):
    g = np.load(os.path.join(GOLD, "nms_ref_random.npz"))
    for i, (n, thr, seed) in enumerate(g["cases"]):
        boxes, scores = _inputs.nms_boxes(int(n), int(seed))
        keep = oracle_mod.nms(boxes, scores, float(thr))
        np.testing.assert_array_equal(keep.numpy(), g["keep%d" % i])
        frac = len(keep) / n
        assert 0.05 < frac < 0.95  # the inputs really exercise suppression


def test_nms_vs_ref_live(oracle_mod):
    ref = oracle_mod.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    for seed in range(5):
        boxes, scores = _inputs.nms_boxes(700 + 100 * seed, 50 + seed)
        for thr in (0.3, 0.5, 0.7):
            assert torch.equal(oracle_mod.nms(boxes, scores, thr), ref.nms(boxes, scores, thr))


def test_nms_edge_cases(oracle_mod):
    assert oracle_mod.nms(torch.zeros(0, 4), torch.zeros(0), 0.5).numel() == 0
    b = torch.tensor([[0., 0., 10., 10.]])
    assert oracle_mod.nms(b, torch.tensor([0.3]), 0.5).tolist() == [0]
    # identical boxes: only the best survives, even at thr == 1.0 (IoU == 1 >= 1)
    b = torch.tensor([[0., 0., 10., 10.]] * 4)
    assert oracle_mod.nms(b, torch.tensor([0.1, 0.9, 0.5, 0.2]), 1.0).tolist() == [1]
    # explicit order == stable descending order
    boxes, scores = _inputs.nms_boxes(300, 9, distinct_scores=False)
    order = torch.sort(scores, stable=True, descending=True)[1]
    assert torch.equal(oracle_mod.nms(boxes, scores, 0.5), oracle_mod.nms(boxes, scores, 0.5, order=order))


def test_roi_align_fwd_vs_reference_fixture(oracle_mod):
    g = np.load(os.path.join(GOLD, "roi_align_ref.npz"))
    feat, rois = _inputs.roi_align_small()
    for tag, (ph, pw, s) in {"7x7s2": (7, 7, 2), "14x14s2": (14, 14, 2), "7x7s0": (7, 7, 0), "3x5s1": (3, 5, 1)}.items():
        y = oracle_mod.roi_align_forward(feat, rois, 0.25, ph, pw, s)
        np.testing.assert_array_equal(y.numpy(), g["small_" + tag])  # bit-exact


@pytest.mark.timeout(120)
def test_roi_align_fwd_config1_fixture(oracle_mod):
    g = np.load(os.path.join(GOLD, "roi_align_ref.npz"))
    feat, rois = _inputs.roi_align_config1()
    y = oracle_mod.roi_align_forward(feat, rois, 0.25, 7, 7, 2).numpy().reshape(-1)
    np.testing.assert_array_equal(y[::int(g["config1_stride"])], g["config1_samples"])
    assert abs(float(y.astype(np.float64).sum()) - float(g["config1_sum"])) == 0.0


def test_roi_align_vs_torchvision(oracle_mod):
    feat, rois = _inputs.roi_align_small()
    for (p, s) in ((7, 2), (14, 2), (7, 0)):
        x = feat.clone().requires_grad_(True)
        tv = torchvision.ops.roi_align(x, rois, (p, p), 0.25, s, aligned=False)
        y = oracle_mod.roi_align_forward(feat, rois, 0.25, p, p, s)
        assert torch.equal(y, tv.detach())
        g = torch.randn(tv.shape, generator=torch.Generator().manual_seed(1))
        tv.backward(g)
        gi = oracle_mod.roi_align_backward(g, rois, 0.25, p, p, *feat.shape, s)
        torch.testing.assert_close(gi, x.grad, rtol=1e-5, atol=1e-6)


def test_roi_pool_vs_torchvision(oracle_mod):
    feat, rois = _inputs.roi_align_small()
    x = feat.clone().requires_grad_(True)
    tv = torchvision.ops.roi_pool(x, rois, (7, 7), 0.25)
    y, am = oracle_mod.roi_pool_forward(feat, rois, 0.25, 7, 7)
    assert torch.equal(y, tv.detach())
    g = torch.randn(tv.shape, generator=torch.Generator().manual_seed(2))
    tv.backward(g)
    gi = oracle_mod.roi_pool_backward(g, rois, am, *feat.shape)
    torch.testing.assert_close(gi, x.grad, rtol=1e-6, atol=1e-6)
    assert am.dtype == torch.int32 and int(am.min()) >= -1


def _focal_formula(logits, targets, gamma, alpha):
    """The reference's python CPU path, layers/sigmoid_focal_loss.py:40-50, in float64."""
    logits = logits.double()
    c = torch.arange(1, logits.shape[1] + 1)[None, :]
    t = targets.long()[:, None]
    p = torch.sigmoid(logits)
    return -(t == c).double() * (1 - p) ** gamma * torch.log(p) * alpha \
        - ((t != c) & (t >= 0)).double() * p ** gamma * torch.log(1 - p) * (1 - alpha)


def test_focal_vs_reference_python_formula(oracle_mod):
    logits, targets = _inputs.focal_inputs(2000, 80, 0)
    logits = logits + torch.randn(logits.shape, generator=torch.Generator().manual_seed(3)) * 3
    for gamma, alpha in ((2.0, 0.25), (1.5, 0.5), (0.0, 0.75)):
        x = logits.double().requires_grad_(True)
        want = _focal_formula(x, targets, gamma, alpha)
        got = oracle_mod.sigmoid_focalloss_forward(logits, targets, 80, gamma, alpha)
        torch.testing.assert_close(got.double(), want.detach(), rtol=2e-5, atol=1e-6)
        d = torch.rand(logits.shape, generator=torch.Generator().manual_seed(4))
        want.backward(d.double())
        gb = oracle_mod.sigmoid_focalloss_backward(logits, targets, d, 80, gamma, alpha)
        torch.testing.assert_close(gb.double(), x.grad, rtol=2e-4, atol=2e-6)
    # ignore label: no loss, no gradient
    got = oracle_mod.sigmoid_focalloss_forward(logits, torch.full_like(targets, -1), 80, 2.0, 0.25)
    assert float(got.abs().max()) == 0.0


@pytest.mark.parametrize("groups,dg,stride,pad,dil,use_mask,bias", [
    (1, 1, 1, 1, 1, False, False), (1, 1, 1, 1, 1, True, True), (2, 2, 2, 1, 1, True, False), (1, 4, 1, 2, 2, False, False)])
def test_deform_conv_vs_torchvision(oracle_mod, groups, dg, stride, pad, dil, use_mask, bias):
    g = torch.Generator().manual_seed(11)
    n, c, h, w, co = 2, 8, 11, 13, 12
    x = torch.randn(n, c, h, w, generator=g).requires_grad_(True)
    wt = torch.randn(co, c // groups, 3, 3, generator=g).requires_grad_(True)
    ho, wo = (h + 2 * pad - (dil * 2 + 1)) // stride + 1, (w + 2 * pad - (dil * 2 + 1)) // stride + 1
    off = (torch.randn(n, dg * 18, ho, wo, generator=g) * 2).requires_grad_(True)
    m = torch.rand(n, dg * 9, ho, wo, generator=g).requires_grad_(True) if use_mask else None
    b = torch.randn(co, generator=g).requires_grad_(True) if bias else None
    y = torchvision.ops.deform_conv2d(x, off, wt, b, stride=stride, padding=pad, dilation=dil, mask=m)
    yo = oracle_mod.deform_conv_forward(x.detach(), off.detach(), m.detach() if use_mask else None, wt.detach(),
                                        b.detach() if bias else None, stride, pad, dil, groups, dg)
    torch.testing.assert_close(yo, y.detach(), rtol=1e-5, atol=1e-5)
    go = torch.randn(y.shape, generator=g)
    y.backward(go)
    r = oracle_mod.deform_conv_backward(x.detach(), off.detach(), m.detach() if use_mask else None, wt.detach(), go,
                                        stride, pad, dil, groups, dg, with_bias=bias)
    torch.testing.assert_close(r["grad_input"], x.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(r["grad_offset"], off.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(r["grad_weight"], wt.grad, rtol=1e-4, atol=1e-4)
    if use_mask:
        torch.testing.assert_close(r["grad_mask"], m.grad, rtol=1e-4, atol=1e-5)
    if bias:
        torch.testing.assert_close(r["grad_bias"], b.grad, rtol=1e-4, atol=1e-4)


def test_deform_psroi_properties(oracle_mod):
    """No independent implementation exists in this image (parity unpinned): check invariants.
    With no_trans the op is a position-sensitive average: constant input -> constant output,
    and the backward is the adjoint of the forward (<out, g> == <data, backward(g)>)."""
    g = torch.Generator().manual_seed(5)
    out_dim, gs, pooled = 3, 2, 4
    data = torch.randn(2, out_dim * gs * gs, 20, 24, generator=g)
    rois = _inputs.rois_for_level(6, 2, 3, img=(96, 80), min_size=8, max_size=60)
    const = torch.full_like(data, 2.5)
    out, cnt = oracle_mod.deform_psroi_forward(const, rois, None, True, 0.25, out_dim, gs, pooled, pooled, 3, 0.0)
    assert torch.all((out == 2.5) | (cnt == 0))
    out, cnt = oracle_mod.deform_psroi_forward(data, rois, None, True, 0.25, out_dim, gs, pooled, pooled, 3, 0.0)
    go = torch.randn(out.shape, generator=g)
    gi, _ = oracle_mod.deform_psroi_backward(go, data, rois, None, cnt, True, 0.25, out_dim, gs, pooled, pooled, 3, 0.0)
    lhs, rhs = float((out * go).sum()), float((data * gi).sum())
    assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(lhs))
    # with trans: finite-difference check of d out / d trans
    trans = (torch.rand(6, 2, pooled, pooled, generator=g) - 0.5)
    o0, c0 = oracle_mod.deform_psroi_forward(data, rois, trans, False, 0.25, out_dim, gs, pooled, pooled, 3, 0.1)
    _, gt = oracle_mod.deform_psroi_backward(go, data, rois, trans, c0, False, 0.25, out_dim, gs, pooled, pooled, 3, 0.1)
    assert torch.isfinite(gt).all() and float(gt.abs().max()) > 0
