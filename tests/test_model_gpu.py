"""GPU: the harness model on the B200 backend (tcgen05 convs, fused FPN ROIAlign, batched NMS)
against the CPU checker backend (fp32 PyTorch + oracle) with identical weights, and the fused
multi-level ROIAlign kernel against the oracle."""
import pytest
import torch

import _inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tiny_cfg(**kw):
    from mrb_b200.model import RCNNConfig
    base = dict(stem_out=16, width_per_group=16, res2_out=64, fpn_out=64, mlp_head_dim=128, mask_conv_layers=(64, 64),
                roi_batch_size=64, rpn_batch_size=64, pre_nms_top_n_train=300, post_nms_top_n_train=300,
                fpn_post_nms_top_n_train=300, pre_nms_top_n_test=200, post_nms_top_n_test=200, fpn_post_nms_top_n_test=200)
    base.update(kw)
    return RCNNConfig(**base)


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-6)


@pytest.mark.parametrize("nhwc", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [7, 14])
def test_roi_align_fpn_vs_oracle(built_lib, oracle_mod, nhwc, dtype, p):
    from mrb_b200 import ops
    from oracle.cpu_backend import CpuCheckerBackend
    feats = [f.to(dtype) for f in _inputs.fpn_features(2, 5, channels=64)]
    rois = _inputs.rois_for_level(300, 2, 21)
    scales = (0.25, 0.125, 0.0625, 0.03125)
    cpu_feats = [f.float().clone().requires_grad_(True) for f in feats]
    want = CpuCheckerBackend().roi_align_fpn(cpu_feats, rois, scales, p, 2, False)
    dfe = [f.detach().to(DEV).contiguous(memory_format=torch.channels_last).clone().requires_grad_(True) for f in feats]
    got = ops.roi_align_fpn(dfe, rois.to(DEV), scales, p, 2, out_nhwc=nhwc)
    assert got.shape == want.shape
    if dtype == torch.float32:
        assert torch.equal(got.detach().cpu(), want.detach())        # same arithmetic as the reference kernel
    else:
        assert _rel(got.detach(), want.detach()) < 1e-2              # one bf16 rounding of the output
    g = torch.randn(want.shape, generator=torch.Generator().manual_seed(1))
    want.backward(g)
    got.backward(g.to(DEV).to(dtype))
    for a, b in zip(dfe, cpu_feats):
        assert _rel(a.grad, b.grad) < (1e-4 if dtype == torch.float32 else 2e-2)


def test_backbone_fpn_matches_fp32_reference(built_lib):
    from mrb_b200.model import GeneralizedRCNN
    from mrb_b200.model.backend import B200Backend
    from oracle.cpu_backend import CpuCheckerBackend
    torch.manual_seed(0)
    cfg = _tiny_cfg()
    ref = GeneralizedRCNN(cfg, CpuCheckerBackend()).eval()
    g = torch.Generator().manual_seed(1)
    for n, b in ref.named_buffers():
        if n.endswith("running_var"):
            b.copy_(torch.rand(b.shape, generator=g) + 0.5)
        elif n.endswith("running_mean"):
            b.copy_(torch.randn(b.shape, generator=g) * 0.1)
    mine = GeneralizedRCNN(cfg, B200Backend()).to(DEV).eval()
    mine.load_state_dict(ref.state_dict())
    imgs = torch.randn(2, 3, 192, 256, generator=g) * 50
    with torch.no_grad():
        want = ref.backbone.run(ref.be, imgs)
        got = mine.backbone.run(mine.be, imgs.to(DEV))
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert _rel(a, b) < 4e-2, _rel(a, b)   # bf16 activations through ~50 layers


def test_train_step_runs_and_grads_match_direction(built_lib):
    """Full train step on the B200 backend: finite losses, a gradient for every trainable parameter,
    and loss values close to the fp32 checker given identical weights and inputs (the random sampling
    differs, so only the deterministic RPN/backbone statistics are compared loosely)."""
    from mrb_b200.model import GeneralizedRCNN
    from mrb_b200.model.backend import B200Backend
    torch.manual_seed(0)
    cfg = _tiny_cfg()
    model = GeneralizedRCNN(cfg, B200Backend()).to(DEV).train()
    g = torch.Generator().manual_seed(2)
    imgs = (torch.randn(2, 3, 256, 320, generator=g) * 50).to(DEV)
    sizes = [(256, 300), (240, 320)]
    tg = [{"boxes": torch.tensor([[10., 10, 120, 150], [90, 60, 280, 220]], device=DEV), "labels": torch.tensor([3, 7], device=DEV)},
          {"boxes": torch.tensor([[30., 40, 200, 200]], device=DEV), "labels": torch.tensor([5], device=DEV)}]
    losses = model(imgs, sizes, tg)
    assert set(losses) == {"loss_objectness", "loss_rpn_box_reg", "loss_classifier", "loss_box_reg", "loss_mask"}
    total = sum(losses.values())
    assert torch.isfinite(total)
    total.backward()
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
        else:
            assert p.grad is None, n
    model.eval()
    with torch.no_grad():
        dets = model(imgs, sizes)
    assert len(dets) == 2 and all(d["boxes"].shape[1] == 4 for d in dets)


def test_conv_autograd_matches_torch(built_lib):
    """_ConvFn backward (ReLU mask -> tcgen05 dgrad with folded BN scale, wgrad, bias, residual) vs autograd
    of the same expression in fp32."""
    from mrb_b200.model.backend import B200Backend
    be = B200Backend()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 24, 40, generator=g)
    w = torch.randn(128, 64, 3, 3, generator=g) / 24
    scale, shift = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    res = torch.randn(2, 128, 24, 40, generator=g)
    go = torch.randn(2, 128, 24, 40, generator=g)
    x16, w16f, r16 = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), res.to(torch.bfloat16).float()
    xr, wr, rr = x16.clone().requires_grad_(True), w16f.clone().requires_grad_(True), r16.clone().requires_grad_(True)
    y = torch.relu(torch.nn.functional.conv2d(xr, wr, padding=1) * scale[None, :, None, None] + shift[None, :, None, None] + rr)
    y.backward(go.to(torch.bfloat16).float())
    xd = x.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wd = w16f.to(DEV).requires_grad_(True)
    rd = res.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    yd = be.conv(xd, wd, scale.to(DEV), shift.to(DEV), None, rd, 1, 1, True)
    assert _rel(yd.detach(), y.detach()) < 1e-2
    yd.backward(go.to(DEV).to(torch.bfloat16))
    assert _rel(xd.grad, xr.grad) < 2e-2
    assert _rel(wd.grad, wr.grad) < 2e-2
    assert _rel(rd.grad, rr.grad) < 1e-2


def test_sgd_momentum_step_matches_torch_sgd(built_lib):
    """mrb_sgd_momentum_step == torch.optim.SGD(momentum, weight_decay) over several steps (reference
    solver/build.py:7-20), with the bf16 copy refreshed and the gradient buffer zeroed in the same pass."""
    from mrb_b200 import ops
    g = torch.Generator().manual_seed(11)
    n = 4 * 5000 + 3                                      # ragged tail on purpose
    p0 = torch.randn(n, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], lr=0.05, momentum=0.9, weight_decay=1e-2)
    p, m, p16 = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    for it in range(4):
        gr = torch.randn(n, generator=g)
        pr.grad = gr.clone()
        opt.step()
        gd = (gr * 2.0).to(DEV)                           # grad_scale 0.5 undoes the doubling (the 1/world fold)
        ops.sgd_momentum_step(p, gd, m, p16, 0.05, 0.9, 1e-2, 0.5, True)
        assert torch.count_nonzero(gd) == 0
        assert torch.allclose(p.cpu(), pr.detach(), rtol=1e-5, atol=1e-6), it
        assert torch.equal(p16, p.to(torch.bfloat16))


def test_wgrad_and_bias_grad_accumulate(built_lib):
    from mrb_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 64, 20, 28, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    go = torch.randn(2, 128, 20, 28, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    fresh = ops.conv2d_wgrad(x, go, (128, 64, 3, 3), 1, 1)
    acc = torch.full((128, 64, 3, 3), 1.5, device=DEV).contiguous(memory_format=torch.channels_last)
    ops.conv2d_wgrad(x, go, (128, 64, 3, 3), 1, 1, accumulate_into=acc)
    assert _rel(acc - 1.5, fresh) < 1e-5
    bf = ops.bias_grad(go)
    ba = torch.full((128,), -2.0, device=DEV)
    ops.bias_grad(go, accumulate_into=ba)
    assert _rel(ba + 2.0, bf) < 1e-5


def test_param_arena_step_equals_flat_sgd(built_lib):
    """One full train step with gradients red.add-ed into the ParamArena and the fused update kernel must leave
    the same parameters as autograd gradients + FlatSGD (same seed => same sampling), and must leave the bf16
    operand copies and the (zeroed) accumulators consistent."""
    from mrb_b200.model import GeneralizedRCNN
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.optim import FlatSGD, ParamArena
    cfg = _tiny_cfg()
    g = torch.Generator().manual_seed(2)
    imgs = (torch.randn(2, 3, 256, 320, generator=g) * 50).to(DEV)
    sizes = [(256, 300), (240, 320)]
    tg = [{"boxes": torch.tensor([[10., 10, 120, 150], [90, 60, 280, 220]], device=DEV), "labels": torch.tensor([3, 7], device=DEV)},
          {"boxes": torch.tensor([[30., 40, 200, 200]], device=DEV), "labels": torch.tensor([5], device=DEV)}]
    results = []
    for kind in ("flat", "arena"):
        torch.manual_seed(0)
        model = GeneralizedRCNN(cfg, B200Backend()).to(DEV).train()
        for p in model.parameters():
            if p.dim() == 4:
                p.data = p.data.contiguous(memory_format=torch.channels_last)
        if kind == "flat":
            opt = FlatSGD(model.named_parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
        else:
            opt = ParamArena(model.named_parameters(), model.be, lr=1e-3, momentum=0.9, weight_decay=1e-4)
            model.be.enable_overlap(True)          # gradient-sink kernels on the second stream
            model.cfg.parallel_heads = True        # mask branch on its own stream
        before = {n: p.detach().float().clone() for n, p in model.named_parameters() if p.requires_grad}
        for it in range(2):
            torch.manual_seed(100 + it)
            loss = sum(model(imgs, sizes, tg).values())
            assert torch.isfinite(loss), (kind, it)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if kind == "flat":
                model.be.refresh_weights([p for p in model.parameters() if p.requires_grad])
            if it == 0:   # the first update is a deterministic function of the seed (same sampling on both paths)
                results.append({n: p.detach().float() - before[n] for n, p in model.named_parameters() if p.requires_grad})
        if kind == "arena":
            assert torch.count_nonzero(opt.grad) == 0
            assert torch.equal(opt.param16, opt.param.to(torch.bfloat16))
            for n, p in model.named_parameters():
                if p.requires_grad and p.dim() in (2, 4):
                    assert model.be._weight16(p).data_ptr() == opt.views16[id(p)].data_ptr(), n
    flat, arena = results
    for n in flat:
        # identical math up to the fp32 summation order of the split-K red.adds / atomics
        d = (flat[n] - arena[n]).abs().max().item()
        ref = flat[n].abs().max().item()
        assert d <= 5e-2 * ref + 1e-8, (n, d, ref)


def test_deconv2x2_matches_conv_transpose(built_lib):
    """Mask-head ConvTranspose2d(2, 2) + ReLU on the conv engine (strided in-place sub-pixel planes) vs autograd."""
    import torch.nn.functional as F
    from mrb_b200.model.backend import B200Backend
    be = B200Backend()
    g = torch.Generator().manual_seed(31)
    x = torch.randn(5, 64, 14, 14, generator=g).relu()
    wt = torch.randn(64, 32, 2, 2, generator=g) / 8
    b = torch.randn(32, generator=g) * 0.1
    go = torch.randn(5, 32, 28, 28, generator=g)
    x16, w16 = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    xr, wr, br = x16.clone().requires_grad_(True), w16.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.relu(F.conv_transpose2d(xr, wr, br, stride=2))
    y.backward(go.to(torch.bfloat16).float())
    xd = x.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wd, bd = w16.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    yd = be.deconv2x2(xd, wd, bd, relu=True, premask_x=True)
    assert yd.shape == y.shape and yd.is_contiguous(memory_format=torch.channels_last)
    assert _rel(yd.detach(), y.detach()) < 1e-2
    yd.backward(go.to(DEV).to(torch.bfloat16))
    assert _rel(xd.grad, xr.grad * (x16 > 0)) < 2e-2          # premask_x: the producer's ReLU mask rides along
    assert _rel(wd.grad, wr.grad) < 2e-2
    assert _rel(bd.grad, br.grad) < 2e-2


def test_conv_select_matches_index_expression(built_lib):
    """be.conv_select == conv(x)[arange, labels] (mask logits + mask_head/loss.py:120-126), values and gradients."""
    import torch.nn.functional as F
    from mrb_b200.model.backend import B200Backend
    be = B200Backend()
    g = torch.Generator().manual_seed(33)
    x = torch.randn(6, 64, 12, 12, generator=g).relu()
    wt = torch.randn(81, 64, 1, 1, generator=g) / 8
    b = torch.randn(81, generator=g) * 0.1
    lab = torch.tensor([3, 80, 0, 17, 17, 42])
    go = torch.randn(6, 12, 12, generator=g)
    x16, w16 = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    xr, wr, br = x16.clone().requires_grad_(True), w16.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.conv2d(xr, wr, br)[torch.arange(6), lab]
    y.backward(go)
    xd = x.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wd, bd = w16.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    yd = be.conv_select(xd, wd, bd, lab.to(DEV))
    assert yd.shape == y.shape and _rel(yd.detach(), y.detach()) < 1e-2
    yd.backward(go.to(DEV))
    assert _rel(xd.grad, xr.grad) < 2e-2
    assert _rel(wd.grad, wr.grad) < 2e-2
    assert _rel(bd.grad, br.grad) < 2e-2
