"""Grouped 3x3 conv composed from per-super-group calls of the dense engine (mrb_b200/grouped.py).
Passed on the driver's B200 at the end of round 1 (XPASS): a normal parity test since round 2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
DEV = "cuda:0"


@pytest.mark.parametrize("mode", ["native", "composed"])
@pytest.mark.parametrize("c,groups,h,w,stride", [(256, 32, 20, 28, 1), (128, 4, 13, 21, 1), (512, 32, 25, 42, 1), (256, 32, 20, 28, 2),
                                                 (1024, 32, 13, 21, 1)])
def test_grouped_conv_matches_torch(built_lib, monkeypatch, mode, c, groups, h, w, stride):
    """X-101-32x8d layer geometries (reference resnet.py:302-311: groups=32, width 8 * 2^stage => 8/16/32/64 channels per
    group; stride 2 in the 3x3 of a stage's first block): forward (+BN+ReLU), data and weight gradient vs fp32 F.conv2d."""
    from mrb_b200.grouped import conv2d_grouped
    if mode == "composed" and (c > 256 or stride == 2):
        pytest.skip("composed form: covered on the small cases")
    monkeypatch.setenv("MRB_GROUPED", mode)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, c, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(c, c // groups, 3, 3, generator=g) / (9 * c // groups) ** 0.5).to(torch.bfloat16).float()
    scale, shift = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    go = torch.randn(2, c, ho, wo, generator=g).to(torch.bfloat16)
    xr, wr = x.float().requires_grad_(True), wt.clone().requires_grad_(True)
    y = torch.relu(F.conv2d(xr, wr, stride=stride, padding=1, groups=groups) * scale[None, :, None, None] + shift[None, :, None, None])
    y.backward(go.float())
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wd = wt.to(DEV).requires_grad_(True)
    yd = conv2d_grouped(xd, wd, groups, scale.to(DEV), shift.to(DEV), pad=1, relu=True, stride=stride)
    assert yd.shape == y.shape

    def rel(a, b):
        a, b = a.float().cpu(), b.float().cpu()
        return float((a - b).abs().max() / (b.abs().max() + 1e-12))
    assert rel(yd.detach(), y.detach()) < 1e-2
    yd.backward(go.to(DEV))
    assert rel(xd.grad, xr.grad) < 2e-2
    assert rel(wd.grad, wr.grad) < 2e-2


@pytest.mark.parametrize("c,groups", [(256, 32), (512, 32), (128, 4), (2048, 32)])
def test_grouped_prep_kernels_match_the_python_expansion(built_lib, c, groups):
    """csrc/grouped_prep.cu (one launch each) == the torch composition pinned on CPU by tests/test_grouped_cpu.py."""
    from mrb_b200 import ops
    from mrb_b200.grouped import collapse_group_grads, expand_group_weights
    g = torch.Generator().manual_seed(c + groups)
    cg = c // groups
    w16 = torch.randn(c, cg, 3, 3, generator=g).bfloat16().to(DEV).contiguous(memory_format=torch.channels_last)
    scale = (torch.rand(c, generator=g) + 0.5).to(DEV)
    w_exp, wd = ops.grouped_expand_weights(w16, groups, scale, True, True)
    want_exp = expand_group_weights(w16, groups)
    assert torch.equal(w_exp, want_exp)
    want_wd = ops.grouped_dgrad_weights(want_exp, scale)
    assert float((wd.float() - want_wd.float()).abs().max()) <= 1e-2 * float(want_wd.float().abs().max())     # bf16(w * scale) either way
    if c % 128 == 0:
        gw128 = torch.randn(c, 9, 128, generator=g).to(DEV)
        got = ops.grouped_collapse_wgrad(gw128, (c, cg, 3, 3), groups)
        v = gw128.view(c // 128, 2, 64, 9, 2, 64)
        diag = torch.stack([v[:, 0, :, :, 0], v[:, 1, :, :, 1]], 1).reshape(c, 3, 3, 64).permute(0, 3, 1, 2)
        assert torch.equal(got, collapse_group_grads(diag, groups))
        acc = torch.ones(c, cg, 3, 3, device=DEV)                       # NCHW-contiguous destination, accumulate
        ops.grouped_collapse_wgrad(gw128, (c, cg, 3, 3), groups, accumulate_into=acc)
        assert torch.equal(acc, got + 1)
