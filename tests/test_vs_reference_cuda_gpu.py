"""Every `_C` op of libmrb_b200.so against the REFERENCE'S OWN CUDA KERNELS running on the same B200:
oracle/_ref/cuda/mrb_ref_cuda.so = the reference's csrc/cuda/*.cu (ROIAlign_cuda.cu, ROIPool_cuda.cu, nms.cu,
SigmoidFocalLoss_cuda.cu, deform_conv_cuda.cu + deform_conv_kernel_cuda.cu, deform_pool_cuda.cu + deform_pool_kernel_cuda.cu)
compiled unmodified for sm_100a through oracle/ref_cuda + oracle/thc_compat (oracle/build_ref.py::build_cuda).
This pins the ops for which the reference has no CPU implementation (ROIAlign backward, ROIPool, SigmoidFocalLoss,
deformable conv v1/v2, deformable PS-ROI pooling: csrc/ROIAlign.h:44 etc.) to the reference itself rather than to a
restatement.  Tolerances: the reference kernels are compiled with FMA contraction and use atomics, so fp32 results agree
to 1e-5 / 1e-4 (backward), not bit-exactly; integer results (NMS indices, ROIPool argmax) must be identical."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _inputs  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def R(oracle_mod):
    m = oracle_mod.ref_cuda()
    if m is None:
        pytest.skip("oracle/_ref/cuda/mrb_ref_cuda.so absent (built where /root/reference exists)")
    return m


@pytest.fixture(scope="module")
def C(built_lib):
    from maskrcnn_benchmark import _C
    return _C


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("case", ["small", "config1", "p3_1024"])
@pytest.mark.parametrize("P,S", [(7, 2), (14, 2), (7, 0)])
def test_roi_align_fwd_bwd(C, R, case, P, S):
    if case == "small":
        feat, rois = _inputs.roi_align_small()
        scale = 0.25
    elif case == "config1":
        feat, rois = _inputs.roi_align_config1()
        scale = 0.25
    else:
        feat = _inputs.fpn_features(2, 3)[1]
        rois, scale = _inputs.rois_for_level(1024, 2, 9), 0.125
    x, r = feat.to(DEV), rois.to(DEV)
    want = R.roi_align_forward(x, r, scale, P, P, S)
    got = C.roi_align_forward(x, r, scale, P, P, S)
    assert _rel(got, want) < 5e-5      # ours reproduces the CPU kernel bit for bit; the CUDA reference contracts to FMAs
    g = torch.randn_like(want)
    n, c, h, w = feat.shape
    wb = R.roi_align_backward(g, r, scale, P, P, n, c, h, w, S)
    gb = C.roi_align_backward(g, r, scale, P, P, n, c, h, w, S)
    assert _rel(gb, wb) < 1e-4


def test_roi_pool_fwd_bwd(C, R):
    feat = _inputs.fpn_features(2, 4)[2][:, :64].contiguous()
    rois = _inputs.rois_for_level(300, 2, 11)
    x, r = feat.to(DEV), rois.to(DEV)
    want, warg = R.roi_pool_forward(x, r, 1 / 16, 7, 7)
    got, garg = C.roi_pool_forward(x, r, 1 / 16, 7, 7)
    assert torch.equal(got, want) and torch.equal(garg, warg)
    g = torch.randn_like(want)
    n, c, h, w = feat.shape
    wb = R.roi_pool_backward(g, x, r, warg, 1 / 16, 7, 7, n, c, h, w)
    gb = C.roi_pool_backward(g, x, r, garg, 1 / 16, 7, 7, n, c, h, w)
    assert _rel(gb, wb) < 1e-5


@pytest.mark.parametrize("n", [819, 2000, 6000])
@pytest.mark.parametrize("thr", [0.7, 0.5])
def test_nms_indices(C, R, n, thr):
    """csrc/nms.h:10-28 as the reference dispatches CUDA tensors: cat(dets, scores) -> nms_cuda.  The CUDA kernel suppresses
    on IoU > thr, the CPU one (our parity target) on >=; with these inputs no pair sits exactly on the threshold."""
    boxes, scores = _inputs.nms_boxes(n, 40 + n)
    b, s = boxes.to(DEV), scores.to(DEV)
    want = R.nms(torch.cat([b, s[:, None]], 1), thr)
    got = C.nms(b, s, thr)
    assert torch.equal(got.cpu(), want.cpu())


def test_sigmoid_focal_loss(C, R):
    logits, targets = _inputs.focal_inputs(50000, 80, 3)
    l, t = logits.to(DEV), targets.to(DEV)
    want = R.sigmoid_focalloss_forward(l, t, 80, 2.0, 0.25)
    got = C.sigmoid_focalloss_forward(l, t, 80, 2.0, 0.25)
    assert _rel(got, want) < 1e-5
    d = torch.rand_like(want)
    wb = R.sigmoid_focalloss_backward(l, t, d, 80, 2.0, 0.25)
    gb = C.sigmoid_focalloss_backward(l, t, d, 80, 2.0, 0.25)
    assert _rel(gb, wb) < 1e-5


DCN_SHAPES = [(128, 100, 168), (256, 50, 84), (512, 25, 42)]      # BASELINE config 5: res3 / res4 / res5 layers, N = 2


@pytest.mark.parametrize("c,h,w", DCN_SHAPES)
def test_deform_conv_v1_baseline_shapes(C, R, c, h, w):
    g = torch.Generator().manual_seed(c)
    x = torch.randn(2, c, h, w, generator=g).to(DEV)
    wt = (torch.randn(c, c, 3, 3, generator=g) / (9 * c) ** 0.5).to(DEV)
    off = (torch.randn(2, 18, h, w, generator=g) * 2).to(DEV)
    go = torch.randn(2, c, h, w, generator=g).to(DEV)
    geom = (3, 3, 1, 1, 1, 1, 1, 1, 1, 1)          # kW kH dW dH padW padH dilW dilH group deformable_group
    # ours: the whole batch, im2col_step = batch (what layers/dcn/deform_conv_func.py passes)
    out, gi, goff, gw = x.new_empty(2, c, h, w), torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(wt)
    C.deform_conv_forward(x, wt, off, out, x.new_empty(0), x.new_empty(0), *geom, 2)
    C.deform_conv_backward_input(x, off, go, gi, goff, wt, x.new_empty(0), *geom, 2)
    C.deform_conv_backward_parameters(x, off, go, gw, x.new_empty(0), x.new_empty(0), *geom, 1.0, 2)
    # reference: ONE IMAGE PER CALL.  Its v1 host code re-views `columns` / `weight` inside the per-step loop
    # (deform_conv_cuda.cu:226-229) and views a transposed buffer (:240-246); with more than one step, or a step of more than
    # one image, those views throw on a current PyTorch.  The kernels are the same either way; dW accumulates across calls.
    r_out, r_gi, r_goff, r_gw = [], [], [], torch.zeros_like(wt)
    for i in range(2):
        xi, oi, gi_ = x[i:i + 1].contiguous(), off[i:i + 1].contiguous(), go[i:i + 1].contiguous()
        o = x.new_empty(1, c, h, w)
        R.deform_conv_forward(xi, wt, oi, o, x.new_empty(0), x.new_empty(0), *geom, 1)
        a, b = torch.zeros_like(xi), torch.zeros_like(oi)
        R.deform_conv_backward_input(xi, oi, gi_, a, b, wt, x.new_empty(0), *geom, 1)
        R.deform_conv_backward_parameters(xi, oi, gi_, r_gw, x.new_empty(0), x.new_empty(0), *geom, 1.0, 1)
        r_out.append(o); r_gi.append(a); r_goff.append(b)
    want = (torch.cat(r_out), torch.cat(r_gi), torch.cat(r_goff), r_gw)
    for got, ref, tol in zip((out, gi, goff, gw), want, (1e-4, 1e-4, 1e-4, 2e-4)):
        assert _rel(got, ref) < tol


@pytest.mark.parametrize("c,h,w", DCN_SHAPES)
def test_deform_conv_v2_baseline_shapes(C, R, c, h, w):
    g = torch.Generator().manual_seed(7 + c)
    x = torch.randn(2, c, h, w, generator=g).to(DEV)
    wt = (torch.randn(c, c, 3, 3, generator=g) / (9 * c) ** 0.5).to(DEV)
    bias = torch.randn(c, generator=g).to(DEV)
    off = (torch.randn(2, 18, h, w, generator=g) * 2).to(DEV)
    mask = torch.rand(2, 9, h, w, generator=g).to(DEV)
    go = torch.randn(2, c, h, w, generator=g).to(DEV)
    geom = (3, 3, 1, 1, 1, 1, 1, 1, 1, 1)          # kh kw sh sw ph pw dh dw group deformable_group
    outs = []
    for M in (R, C):
        out = x.new_empty(2, c, h, w)
        M.modulated_deform_conv_forward(x, wt, bias, x.new_empty(0), off, mask, out, x.new_empty(0), *geom, True)
        gi, gw, gb = torch.zeros_like(x), torch.zeros_like(wt), torch.zeros_like(bias)
        goff, gm = torch.zeros_like(off), torch.zeros_like(mask)
        M.modulated_deform_conv_backward(x, wt, bias, x.new_empty(0), off, mask, x.new_empty(0), gi, gw, gb, goff, gm, go, *geom, True)
        outs.append((out, gi, gw, gb, goff, gm))
    for a, b, tol in zip(outs[1], outs[0], (1e-4, 1e-4, 2e-4, 2e-4, 1e-4, 1e-4)):
        assert _rel(a, b) < tol


@pytest.mark.parametrize("no_trans", [1, 0])
@pytest.mark.parametrize("geom", [(4, 3, 6, 3), (8, 7, 7, 4)])     # (output_dim, group_size, pooled, sample_per_part)
def test_deform_psroi_pooling(C, R, no_trans, geom):
    """The pin the PS-ROI oracle lacked in round 1: csrc/cuda/deform_pool_kernel_cuda.cu itself."""
    out_dim, gs, pooled, spp = geom
    g = torch.Generator().manual_seed(5 + out_dim)
    data = torch.randn(2, out_dim * gs * gs, 38, 50, generator=g).to(DEV)
    rois = _inputs.rois_for_level(64, 2, 3, img=(200, 152), min_size=8, max_size=120).to(DEV)
    trans = ((torch.rand(64, 2, pooled, pooled, generator=g) - 0.5).to(DEV)) if not no_trans else data.new_zeros(0)
    go = torch.randn(64, out_dim, pooled, pooled, generator=g).to(DEV)
    res = []
    for M in (R, C):
        out, cnt = data.new_zeros(64, out_dim, pooled, pooled), data.new_zeros(64, out_dim, pooled, pooled)
        M.deform_psroi_pooling_forward(data, rois, trans, out, cnt, no_trans, 0.25, out_dim, gs, pooled, pooled, spp, 0.1)
        gi, gt = torch.zeros_like(data), torch.zeros_like(trans)
        M.deform_psroi_pooling_backward(go, data, rois, trans, cnt, gi, gt, no_trans, 0.25, out_dim, gs, pooled, pooled, spp, 0.1)
        res.append((out, cnt, gi, gt))
    assert _rel(res[1][0], res[0][0]) < 1e-5
    assert torch.equal(res[1][1], res[0][1])
    assert _rel(res[1][2], res[0][2]) < 1e-4
    if not no_trans:
        assert _rel(res[1][3], res[0][3]) < 1e-4
