"""Second half of __graft_entry__.smoke(): the tcgen05 conv engine (forward with fused BN/residual/ReLU, data
gradient, weight gradient) against fp32 F.conv2d on the same bf16-rounded operands, through layers.Conv2d as well,
and one tiny Mask R-CNN train step of the harness (forward + backward + fused SGD)."""
import torch
import torch.nn.functional as F


def run(dev):
    from maskrcnn_benchmark import layers
    from mrb_b200 import ops
    g = torch.Generator().manual_seed(0)
    n, ci, co, h, w = 2, 64, 128, 24, 40
    x = torch.randn(n, ci, h, w, generator=g).bfloat16()
    wt = (torch.randn(co, ci, 3, 3, generator=g) / 24).bfloat16()
    sc, sh = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
    res = torch.randn(n, co, h, w, generator=g).bfloat16()
    go = torch.randn(n, co, h, w, generator=g).bfloat16()
    cl = dict(memory_format=torch.channels_last)
    xd, wd, rd, gd = (t.to(dev).contiguous(**cl) for t in (x, wt, res, go))
    y = ops.conv2d_fwd(xd, wd, sc.to(dev), sh.to(dev), rd, 1, 1, True)
    ref = F.relu(F.conv2d(x.float(), wt.float(), None, 1, 1) * sc[None, :, None, None] + sh[None, :, None, None] + res.float())
    torch.testing.assert_close(y.float().cpu(), ref, rtol=2e-2, atol=2e-2)
    gx = ops.conv2d_dgrad(gd, wd, (n, ci, h, w), None, None, None, 1, 1)
    gw = ops.conv2d_wgrad(xd, gd, wt.shape, 1, 1)
    xr, wr = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    F.conv2d(xr, wr, None, 1, 1).backward(go.float())
    torch.testing.assert_close(gx.float().cpu(), xr.grad, rtol=2e-2, atol=5e-2)
    torch.testing.assert_close(gw.cpu(), wr.grad, rtol=1e-3, atol=5e-2)
    # the drop-in boundary: layers.Conv2d on a CUDA tensor is served by the engine
    conv = layers.Conv2d(ci, co, 3, padding=1).to(dev)
    with torch.no_grad():
        conv.weight.copy_(wt.float())
        conv.bias.zero_()
    from mrb_b200 import engine
    before = engine.STATS["engine"]
    yl = conv(x.float().to(dev))
    assert engine.STATS["engine"] == before + 1 and yl.dtype == torch.float32
    torch.testing.assert_close(yl.cpu(), F.conv2d(x.float(), wt.float(), None, 1, 1), rtol=1e-3, atol=1e-3)
    # one tiny train step of the harness model
    from mrb_b200.model import RCNNConfig, build_model
    from mrb_b200.optim import ParamArena
    torch.manual_seed(0)
    cfg = RCNNConfig(stem_out=8, width_per_group=8, res2_out=32, fpn_out=32, mlp_head_dim=64, mask_conv_layers=(16, 16, 16, 16),
                     pre_nms_top_n_train=200, post_nms_top_n_train=200, fpn_post_nms_top_n_train=200, roi_batch_size=64)
    model = build_model(cfg, device=dev).train()
    opt = ParamArena(model.named_parameters(), model.be, lr=1e-3)
    images = torch.randn(2, 3, 128, 160, generator=g).to(dev) * 40
    bx = torch.tensor([[10., 12., 70., 90.], [60., 30., 140., 100.], [5., 60., 50., 110.]], device=dev)
    targets = [{"boxes": bx, "labels": torch.tensor([3, 17, 60], device=dev)} for _ in range(2)]
    losses = model(images, [(128, 150), (120, 160)], targets)
    loss = sum(losses.values())
    loss.backward()
    opt.sync()
    opt.step()
    assert torch.isfinite(loss), losses
