"""GPU parity: every kernel of libmrb_b200.so, called through the `_C` / `layers` boundary, against the
oracle on the same seeded inputs, plus the committed reference fixtures.  Bars: NMS indices and ROIPool
argmax bit-exact; ROIAlign forward bit-exact (stronger than the 1e-4 the north star asks); everything
that sums in a different order (atomics, split-K) within rtol 1e-4."""
import os

import numpy as np
import pytest
import torch

import _inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"


@pytest.fixture(scope="module")
def C(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from maskrcnn_benchmark import _C
    return _C


@pytest.fixture(scope="module")
def L(built_lib):
    from maskrcnn_benchmark import layers
    return layers


def close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(a.cpu(), b.cpu(), rtol=rtol, atol=atol)


# ------------------------------------------------------------------------------------------ NMS
def test_nms_reference_known_answers(C):
    g = np.load(os.path.join(GOLD, "nms_reference_tests.npz"))
    for i in range(int(g["n"])):
        keep = C.nms(torch.from_numpy(g["boxes%d" % i]).to(DEV), torch.from_numpy(g["scores%d" % i]).to(DEV),
                     float(g["thr%d" % i]))
        assert keep.dtype == torch.int64 and keep.is_cuda
        np.testing.assert_array_equal(keep.cpu().numpy(), g["keep%d" % i])


def test_nms_reference_cpu_fixture(C):
    g = np.load(os.path.join(GOLD, "nms_ref_random.npz"))
    for i, (n, thr, seed) in enumerate(g["cases"]):
        boxes, scores = _inputs.nms_boxes(int(n), int(seed))
        keep = C.nms(boxes.to(DEV), scores.to(DEV), float(thr))
        np.testing.assert_array_equal(keep.cpu().numpy(), g["keep%d" % i])


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 128, 129, 1000, 2000, 4097, 12000])
def test_nms_vs_oracle(C, oracle_mod, n):
    boxes, scores = _inputs.nms_boxes(n, 100 + n)
    for thr in (0.5, 0.7):
        keep = C.nms(boxes.to(DEV), scores.to(DEV), thr)
        assert torch.equal(keep.cpu(), oracle_mod.nms(boxes, scores, thr))
        assert torch.all(keep[1:] > keep[:-1])  # ascending by index (nms_cpu.cpp:64)


def test_nms_ties_and_duplicates(C, oracle_mod):
    boxes, scores = _inputs.nms_boxes(1500, 77, distinct_scores=False)  # 8 distinct score values
    boxes[100:200] = boxes[0:100]  # exact duplicate boxes
    order = torch.sort(scores, stable=True, descending=True)[1]
    want = oracle_mod.nms(boxes, scores, 0.5, order=order)  # tie rule: ascending index
    assert torch.equal(C.nms(boxes.to(DEV), scores.to(DEV), 0.5).cpu(), want)


def test_nms_edge_cases(C):
    e = C.nms(torch.zeros(0, 4, device=DEV), torch.zeros(0, device=DEV), 0.5)
    assert e.numel() == 0 and e.dtype == torch.int64 and e.device.type == "cpu"  # csrc/nms.h:17-18
    b = torch.tensor([[0., 0., 10., 10.]] * 4, device=DEV)
    assert C.nms(b, torch.tensor([0.1, 0.9, 0.5, 0.2], device=DEV), 1.0).tolist() == [1]
    far = torch.tensor([[0., 0., 1., 1.], [100., 100., 101., 101.]], device=DEV)
    assert C.nms(far, torch.tensor([0.5, 0.6], device=DEV), 0.0001).tolist() == [0, 1]
    with pytest.raises(RuntimeError):
        C.nms(far.half(), torch.tensor([0.5, 0.6], device=DEV).half(), 0.5)


def test_nms_layer_forces_fp32_under_autocast(L, oracle_mod):
    boxes, scores = _inputs.nms_boxes(500, 5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        keep = L.nms(boxes.to(DEV), scores.to(DEV), 0.7)
    assert torch.equal(keep.cpu(), oracle_mod.nms(boxes, scores, 0.7))


def test_nms_batched_matches_single(C, oracle_mod):
    import ctypes
    sizes = [2000, 819, 0, 1, 1333, 64]
    bs = [_inputs.nms_boxes(n, 300 + i) if n else (torch.zeros(0, 4), torch.zeros(0)) for i, n in enumerate(sizes)]
    boxes = torch.cat([b for b, _ in bs]).to(DEV)
    scores = torch.cat([s for _, s in bs]).to(DEV)
    offs = np.cumsum([0] + sizes).astype(np.int32)
    offs_c = (ctypes.c_int * len(offs))(*offs.tolist())
    lib = C.lib
    ws_bytes = lib.mrb_nms_batched_workspace_bytes(offs_c, len(sizes))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    keep = torch.full((int(offs[-1]),), -1, dtype=torch.int64, device=DEV)
    cnt = torch.empty(len(sizes), dtype=torch.int32, device=DEV)
    rc = lib.mrb_nms_batched(C._ptr(boxes), C._ptr(scores), offs_c, len(sizes), ctypes.c_float(0.7), C._ptr(keep),
                             C._ptr(cnt), C._ptr(ws), ctypes.c_size_t(ws_bytes), C._stream())
    assert rc == 0
    cnt = cnt.cpu().tolist()
    for i, (b, s) in enumerate(bs):
        want = oracle_mod.nms(b, s, 0.7) if sizes[i] else torch.zeros(0, dtype=torch.int64)
        assert cnt[i] == len(want)
        assert torch.equal(keep[offs[i]:offs[i] + cnt[i]].cpu(), want)


# ------------------------------------------------------------------------------------- ROIAlign
@pytest.mark.parametrize("ph,pw,s", [(7, 7, 2), (14, 14, 2), (7, 7, 0), (3, 5, 1), (1, 1, 4)])
def test_roi_align_forward_small_bit_exact(C, oracle_mod, ph, pw, s):
    feat, rois = _inputs.roi_align_small()
    want = oracle_mod.roi_align_forward(feat, rois, 0.25, ph, pw, s)
    got = C.roi_align_forward(feat.to(DEV), rois.to(DEV), 0.25, ph, pw, s)
    assert got.shape == want.shape and got.is_contiguous()
    assert torch.equal(got.cpu(), want)
    got_cl = C.roi_align_forward(feat.to(DEV).contiguous(memory_format=torch.channels_last), rois.to(DEV), 0.25, ph, pw, s)
    assert torch.equal(got_cl.cpu(), want)


def test_roi_align_forward_reference_fixture(C):
    g = np.load(os.path.join(GOLD, "roi_align_ref.npz"))
    feat, rois = _inputs.roi_align_small()
    for tag, (ph, pw, s) in {"7x7s2": (7, 7, 2), "14x14s2": (14, 14, 2), "7x7s0": (7, 7, 0), "3x5s1": (3, 5, 1)}.items():
        got = C.roi_align_forward(feat.to(DEV), rois.to(DEV), 0.25, ph, pw, s)
        np.testing.assert_array_equal(got.cpu().numpy(), g["small_" + tag])
    # BASELINE config 1 (1x256x200x336, 100 boxes, 7x7, S=2, scale 0.25) vs the reference CPU _C
    feat, rois = _inputs.roi_align_config1()
    for x in (feat.to(DEV), feat.to(DEV).contiguous(memory_format=torch.channels_last)):
        y = C.roi_align_forward(x, rois.to(DEV), 0.25, 7, 7, 2).cpu().numpy().reshape(-1)
        np.testing.assert_array_equal(y[::int(g["config1_stride"])], g["config1_samples"])
        assert float(y.astype(np.float64).sum()) == float(g["config1_sum"])


@pytest.mark.parametrize("level,scale", [(0, 0.25), (2, 0.0625), (3, 0.03125)])
@pytest.mark.parametrize("p", [7, 14])
def test_roi_align_fpn_shapes_vs_oracle(C, oracle_mod, level, scale, p):
    feats = _inputs.fpn_features(2, level, channels=64)
    feat = feats[level]
    rois = _inputs.rois_for_level(96, 2, 10 + level)
    want = oracle_mod.roi_align_forward(feat, rois, scale, p, p, 2)
    assert torch.equal(C.roi_align_forward(feat.to(DEV), rois.to(DEV), scale, p, p, 2).cpu(), want)
    cl = feat.to(DEV).contiguous(memory_format=torch.channels_last)
    assert torch.equal(C.roi_align_forward(cl, rois.to(DEV), scale, p, p, 2).cpu(), want)
    # backward (atomics: order differs) within 1e-4
    g = torch.randn(want.shape, generator=torch.Generator().manual_seed(3))
    wantb = oracle_mod.roi_align_backward(g, rois, scale, p, p, *feat.shape, 2)
    gotb = C.roi_align_backward(g.to(DEV), rois.to(DEV), scale, p, p, *feat.shape, 2)
    close(gotb, wantb, rtol=1e-4, atol=1e-5)


def test_roi_align_odd_channels_and_empty(C, oracle_mod):
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(2, 7, 20, 30, generator=g)
    rois = _inputs.rois_for_level(11, 2, 4, img=(120, 80), min_size=4, max_size=100)
    want = oracle_mod.roi_align_forward(feat, rois, 0.25, 7, 7, 2)
    assert torch.equal(C.roi_align_forward(feat.to(DEV), rois.to(DEV), 0.25, 7, 7, 2).cpu(), want)
    cl = feat.to(DEV).contiguous(memory_format=torch.channels_last)
    assert torch.equal(C.roi_align_forward(cl, rois.to(DEV), 0.25, 7, 7, 2).cpu(), want)
    out = C.roi_align_forward(feat.to(DEV), torch.zeros(0, 5, device=DEV), 0.25, 7, 7, 2)
    assert out.shape == (0, 7, 7, 7)
    gb = C.roi_align_backward(torch.zeros(0, 7, 7, 7, device=DEV), torch.zeros(0, 5, device=DEV), 0.25, 7, 7, 2, 7, 20, 30, 2)
    assert gb.shape == (2, 7, 20, 30) and float(gb.abs().sum()) == 0.0


def test_roi_align_module_autograd(L, oracle_mod):
    feat, rois = _inputs.roi_align_small()
    x = feat.to(DEV).requires_grad_(True)
    m = L.ROIAlign((7, 7), 0.25, 2)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x, rois.to(DEV))
    assert y.dtype == torch.float32  # fp32-forced like amp.float_function
    g = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    y.backward(g.to(DEV))
    close(x.grad, oracle_mod.roi_align_backward(g, rois, 0.25, 7, 7, *feat.shape, 2))


# -------------------------------------------------------------------------------------- ROIPool
def test_roi_pool_vs_oracle(C, L, oracle_mod):
    feat, rois = _inputs.roi_align_small()
    want, wam = oracle_mod.roi_pool_forward(feat, rois, 0.25, 7, 7)
    got, am = C.roi_pool_forward(feat.to(DEV), rois.to(DEV), 0.25, 7, 7)
    assert torch.equal(got.cpu(), want) and torch.equal(am.cpu(), wam) and am.dtype == torch.int32
    g = torch.randn(want.shape, generator=torch.Generator().manual_seed(2))
    wantb = oracle_mod.roi_pool_backward(g, rois, wam, *feat.shape)
    gotb = C.roi_pool_backward(g.to(DEV), feat.to(DEV), rois.to(DEV), am, 0.25, 7, 7, *feat.shape)
    close(gotb, wantb)
    x = feat.to(DEV).requires_grad_(True)
    y = L.ROIPool((7, 7), 0.25)(x, rois.to(DEV))
    y.backward(g.to(DEV))
    close(x.grad, wantb)


# ----------------------------------------------------------------------------- SigmoidFocalLoss
@pytest.mark.parametrize("gamma,alpha", [(2.0, 0.25), (1.5, 0.5), (0.0, 0.75)])
@pytest.mark.parametrize("nc", [80, 5])
def test_focal_vs_oracle(C, oracle_mod, gamma, alpha, nc):
    logits, targets = _inputs.focal_inputs(5000, nc, 1)
    logits = logits + torch.randn(logits.shape, generator=torch.Generator().manual_seed(3)) * 4
    logits[0, 0], logits[1, 1], logits[2, 2] = -95.0, 60.0, 0.0  # clamp / saturation corners
    targets[0], targets[1] = 1, 2
    want = oracle_mod.sigmoid_focalloss_forward(logits, targets, nc, gamma, alpha)
    got = C.sigmoid_focalloss_forward(logits.to(DEV), targets.to(DEV), nc, gamma, alpha)
    close(got, want, rtol=1e-4, atol=1e-6)
    d = torch.rand(logits.shape, generator=torch.Generator().manual_seed(4))
    wantb = oracle_mod.sigmoid_focalloss_backward(logits, targets, d, nc, gamma, alpha)
    gotb = C.sigmoid_focalloss_backward(logits.to(DEV), targets.to(DEV), d.to(DEV), nc, gamma, alpha)
    close(gotb, wantb, rtol=1e-4, atol=1e-6)


def test_focal_module(L, oracle_mod):
    logits, targets = _inputs.focal_inputs(3000, 80, 2)
    x = logits.to(DEV).requires_grad_(True)
    loss = L.SigmoidFocalLoss(2.0, 0.25)(x, targets.to(DEV))
    want = oracle_mod.sigmoid_focalloss_forward(logits, targets, 80, 2.0, 0.25)
    assert abs(float(loss) - float(want.double().sum())) <= 1e-4 * float(want.double().sum())
    loss.backward()
    close(x.grad, oracle_mod.sigmoid_focalloss_backward(logits, targets, torch.ones_like(logits), 80, 2.0, 0.25),
          rtol=1e-4, atol=1e-6)
    with pytest.raises(RuntimeError):
        C_targets64 = targets.long().to(DEV)
        L.SigmoidFocalLoss(2.0, 0.25)(x, C_targets64)


# ------------------------------------------------------------------------------ deformable conv
def _dcn_case(groups, dg, stride, pad, dil, use_mask, bias, seed=11, n=2, c=16, h=19, w=23, co=24):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g)
    wt = torch.randn(co, c // groups, 3, 3, generator=g) * 0.2
    ho, wo = (h + 2 * pad - (dil * 2 + 1)) // stride + 1, (w + 2 * pad - (dil * 2 + 1)) // stride + 1
    off = torch.randn(n, dg * 18, ho, wo, generator=g) * 2
    m = torch.rand(n, dg * 9, ho, wo, generator=g) if use_mask else None
    b = torch.randn(co, generator=g) if bias else None
    go = torch.randn(n, co, ho, wo, generator=g)
    return x, wt, off, m, b, go


@pytest.mark.parametrize("groups,dg,stride,pad,dil", [(1, 1, 1, 1, 1), (2, 2, 2, 1, 1), (1, 4, 1, 2, 2)])
def test_deform_conv_v1_layer(L, oracle_mod, groups, dg, stride, pad, dil):
    x, wt, off, _, _, go = _dcn_case(groups, dg, stride, pad, dil, False, False)
    xd, wd, od = x.to(DEV).requires_grad_(True), wt.to(DEV).requires_grad_(True), off.to(DEV).requires_grad_(True)
    y = L.deform_conv(xd, od, wd, stride, pad, dil, groups, dg)
    close(y, oracle_mod.deform_conv_forward(x, off, None, wt, None, stride, pad, dil, groups, dg), rtol=1e-4, atol=1e-4)
    y.backward(go.to(DEV))
    r = oracle_mod.deform_conv_backward(x, off, None, wt, go, stride, pad, dil, groups, dg)
    close(xd.grad, r["grad_input"], rtol=1e-4, atol=1e-4)
    close(od.grad, r["grad_offset"], rtol=1e-4, atol=1e-4)
    close(wd.grad, r["grad_weight"], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("groups,dg,stride,pad,dil,bias", [(1, 1, 1, 1, 1, True), (2, 2, 2, 1, 1, False), (1, 1, 1, 1, 1, False)])
def test_deform_conv_v2_layer(L, oracle_mod, groups, dg, stride, pad, dil, bias):
    x, wt, off, m, b, go = _dcn_case(groups, dg, stride, pad, dil, True, bias)
    xd, wd, od, md = (t.to(DEV).requires_grad_(True) for t in (x, wt, off, m))
    bd = b.to(DEV).requires_grad_(True) if bias else None
    y = L.modulated_deform_conv(xd, od, md, wd, bd, stride, pad, dil, groups, dg)
    close(y, oracle_mod.deform_conv_forward(x, off, m, wt, b, stride, pad, dil, groups, dg), rtol=1e-4, atol=1e-4)
    y.backward(go.to(DEV))
    r = oracle_mod.deform_conv_backward(x, off, m, wt, go, stride, pad, dil, groups, dg, with_bias=bias)
    close(xd.grad, r["grad_input"], rtol=1e-4, atol=1e-4)
    close(od.grad, r["grad_offset"], rtol=1e-4, atol=1e-4)
    close(md.grad, r["grad_mask"], rtol=1e-4, atol=1e-4)
    close(wd.grad, r["grad_weight"], rtol=1e-4, atol=1e-3)
    if bias:
        close(bd.grad, r["grad_bias"], rtol=1e-4, atol=1e-3)


def test_dfconv2d_block(L, oracle_mod):
    torch.manual_seed(0)
    for modulated in (True, False):
        blk = L.DFConv2d(16, 32, with_modulated_dcn=modulated).to(DEV)
        x = torch.randn(2, 16, 20, 28, device=DEV)
        y = blk(x)
        assert y.shape == (2, 32, 20, 28)
        om = blk.offset(x)
        if modulated:
            want = oracle_mod.deform_conv_forward(x.cpu(), om[:, :18].detach().cpu(), om[:, -9:].sigmoid().detach().cpu(),
                                                  blk.conv.weight.detach().cpu(), None, 1, 1, 1, 1, 1)
        else:
            want = oracle_mod.deform_conv_forward(x.cpu(), om.detach().cpu(), None, blk.conv.weight.detach().cpu(),
                                                  None, 1, 1, 1, 1, 1)
        close(y.detach(), want, rtol=1e-4, atol=1e-4)
        y.sum().backward()
        assert blk.conv.weight.grad is not None and blk.offset.weight.grad is not None


# --------------------------------------------------------------------- deformable PS-ROI pooling
@pytest.mark.parametrize("no_trans", [True, False])
def test_deform_psroi(L, oracle_mod, no_trans):
    g = torch.Generator().manual_seed(5)
    out_dim, gs, pooled, spp, ts = 4, 3, 6, 3, 0.1
    data = torch.randn(2, out_dim * gs * gs, 24, 30, generator=g)
    rois = _inputs.rois_for_level(9, 2, 3, img=(120, 96), min_size=8, max_size=80)
    trans = (torch.rand(9, 2, pooled, pooled, generator=g) - 0.5) if not no_trans else torch.zeros(0)
    d, tr = data.to(DEV).requires_grad_(True), trans.to(DEV).requires_grad_(not no_trans)
    y = L.deform_roi_pooling(d, rois.to(DEV), tr, 0.25, pooled, out_dim, no_trans, gs, pooled, spp, ts)
    want, cnt = oracle_mod.deform_psroi_forward(data, rois, None if no_trans else trans, no_trans, 0.25, out_dim, gs,
                                                pooled, pooled, spp, ts)
    close(y, want, rtol=1e-4, atol=1e-5)
    go = torch.randn(want.shape, generator=g)
    y.backward(go.to(DEV))
    gi, gt = oracle_mod.deform_psroi_backward(go, data, rois, None if no_trans else trans, cnt, no_trans, 0.25, out_dim,
                                              gs, pooled, pooled, spp, ts)
    close(d.grad, gi, rtol=1e-4, atol=1e-5)
    if not no_trans:
        close(tr.grad, gt, rtol=1e-3, atol=1e-4)
