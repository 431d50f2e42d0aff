"""Grouped-conv groundwork (X-101-32x8d config): the block-diagonal weight expansion used to run a grouped 3x3 layer
as 64 -> 64 super-group convolutions on the dense engine, pinned on CPU against F.conv2d(groups=...)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maskrcnn-benchmark_b200"))


@pytest.mark.parametrize("c,groups", [(256, 32), (128, 8), (64, 1), (128, 2), (512, 32)])
def test_expanded_weights_reproduce_grouped_conv(c, groups):
    from mrb_b200.grouped import SG, collapse_group_grads, expand_group_weights
    g = torch.Generator().manual_seed(c + groups)
    x = torch.randn(2, c, 9, 11, generator=g, dtype=torch.float64)
    w = torch.randn(c, c // groups, 3, 3, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, padding=1, groups=groups)
    w_exp = expand_group_weights(w, groups)
    assert w_exp.shape == (c, SG, 3, 3)
    got = torch.cat([F.conv2d(x[:, s * SG:(s + 1) * SG], w_exp[s * SG:(s + 1) * SG], padding=1) for s in range(c // SG)], 1)
    torch.testing.assert_close(got, want, rtol=1e-12, atol=1e-12)
    # gradient direction: dense super-group weight gradients collapse to the grouped gradient
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    go = torch.randn(want.shape, generator=g, dtype=torch.float64)
    F.conv2d(xr, wr, padding=1, groups=groups).backward(go)
    we = w_exp.clone().requires_grad_(True)
    torch.cat([F.conv2d(x[:, s * SG:(s + 1) * SG], we[s * SG:(s + 1) * SG], padding=1) for s in range(c // SG)], 1).backward(go)
    torch.testing.assert_close(collapse_group_grads(we.grad, groups), wr.grad, rtol=1e-12, atol=1e-12)
    assert torch.equal(collapse_group_grads(w_exp, groups), w)


def test_geometry_is_checked():
    from mrb_b200.grouped import check_geometry
    with pytest.raises(RuntimeError):
        check_geometry(96, 96, 3)        # 96 % 64 != 0
    with pytest.raises(RuntimeError):
        check_geometry(256, 256, 2)      # 128 channels per group > 64
