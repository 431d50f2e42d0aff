"""GPU parity of the tcgen05 conv engine.  Oracle for this floating-point kernel = plain PyTorch fp32
convolution (the reference pins no conv arithmetic: layers.Conv2d is ATen) evaluated on
the SAME bf16-rounded operands; the kernel accumulates in fp32, so the only difference is summation
order: tolerance 1e-4 of the output scale (north_star: "within 1e-4 rel for ... convs")."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops(built_lib):
    assert torch.cuda.is_available()
    from mrb_b200 import ops
    return ops


def _mk(n, c, h, w, co, k, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(co, c, k, k, generator=g) / (c * k * k) ** 0.5).to(torch.bfloat16)
    return x, wt


def _assert_close(got, want, tol=1e-4):
    got, want = got.float().cpu(), want.float().cpu()
    scale = float(want.abs().max()) + 1e-6
    err = float((got - want).abs().max())
    assert err <= tol * scale, "max err %g vs scale %g" % (err, scale)


CASES = [
    # n, cin, h, w, cout, k, stride, pad
    (2, 64, 40, 56, 256, 1, 1, 0),     # 1x1 expand (pure GEMM path)
    (2, 256, 40, 56, 64, 1, 1, 0),     # 1x1 reduce, K = 4 blocks
    (1, 64, 40, 56, 64, 3, 1, 1),      # 3x3, exact 8x16 tiles
    (2, 128, 25, 42, 128, 3, 1, 1),    # 3x3, partial tiles in both dims
    (1, 256, 13, 21, 256, 3, 1, 1),    # P6-like tiny map
    (2, 256, 50, 84, 512, 1, 2, 0),    # 1x1 stride 2 (downsample / STRIDE_IN_1X1), 2 Cout tiles
    (1, 128, 25, 41, 256, 1, 2, 0),    # stride 2 with odd width
    (1, 24, 20, 30, 48, 3, 1, 1),      # Cin not a multiple of 64 (TMA zero-fill along channels)
    (1, 256, 30, 40, 81, 1, 1, 0),     # mask predictor: odd Cout
    (1, 256, 30, 40, 12, 1, 1, 0),     # RPN bbox head: tiny Cout
    (1, 1024, 14, 14, 2048, 1, 1, 0),  # deep K, 8 Cout tiles
    (300, 512, 1, 1, 1024, 1, 1, 0),   # fully connected (ROI box head style)
]


@pytest.mark.parametrize("n,c,h,w,co,k,stride,pad", CASES)
def test_conv_fwd_plain(ops, n, c, h, w, co, k, stride, pad):
    x, wt = _mk(n, c, h, w, co, k, 1)
    want = F.conv2d(x.float(), wt.float(), stride=stride, padding=pad)
    got = ops.conv2d_fwd(x.to(DEV), wt.to(DEV), stride=stride, pad=pad, out_dtype=torch.float32)
    assert got.shape == want.shape
    _assert_close(got, want)


@pytest.mark.parametrize("n,c,h,w,co,k,stride,pad", CASES[:6])
def test_conv_fwd_fused_epilogue(ops, n, c, h, w, co, k, stride, pad):
    x, wt = _mk(n, c, h, w, co, k, 2)
    g = torch.Generator().manual_seed(3)
    scale = torch.rand(co, generator=g) + 0.5
    bias = torch.randn(co, generator=g)
    conv = F.conv2d(x.float(), wt.float(), stride=stride, padding=pad)
    res = torch.randn(conv.shape, generator=g).to(torch.bfloat16)
    want = torch.relu(conv * scale[None, :, None, None] + bias[None, :, None, None] + res.float())
    got = ops.conv2d_fwd(x.to(DEV), wt.to(DEV), scale.to(DEV), bias.to(DEV), res.to(DEV), stride, pad, relu=True,
                         out_dtype=torch.float32)
    _assert_close(got, want)
    got16 = ops.conv2d_fwd(x.to(DEV), wt.to(DEV), scale.to(DEV), bias.to(DEV), res.to(DEV), stride, pad, relu=True)
    assert got16.dtype == torch.bfloat16
    _assert_close(got16, want, tol=1e-2)  # one bf16 rounding of the result
    assert got16.is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("n,c,h,w,co,k,stride,pad", CASES[:8])
def test_conv_dgrad(ops, n, c, h, w, co, k, stride, pad):
    x, wt = _mk(n, c, h, w, co, k, 4)
    g = torch.Generator().manual_seed(5)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    go = torch.randn(n, co, ho, wo, generator=g).to(torch.bfloat16)
    xf = x.float().requires_grad_(True)
    F.conv2d(xf, wt.float(), stride=stride, padding=pad).backward(go.float())
    got = ops.conv2d_dgrad(go.to(DEV), wt.to(DEV), x.shape, stride=stride, pad=pad, out_dtype=torch.float32)
    _assert_close(got, xf.grad)


def test_conv_dgrad_fused(ops):
    n, c, h, w, co, k = 2, 64, 24, 40, 128, 3
    x, wt = _mk(n, c, h, w, co, k, 6)
    g = torch.Generator().manual_seed(7)
    go = torch.randn(n, co, h, w, generator=g).to(torch.bfloat16)
    scale = torch.rand(co, generator=g) + 0.5
    add = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
    mask = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
    xf = x.float().requires_grad_(True)
    # weights are scaled then rounded to bf16 inside the kernel's weight prep: mirror that
    wd = (wt.float() * scale[:, None, None, None]).to(torch.bfloat16).float()
    F.conv2d(xf, wd, padding=1).backward(go.float())
    want = (xf.grad + add.float()) * (mask.float() > 0)
    got = ops.conv2d_dgrad(go.to(DEV), wt.to(DEV), x.shape, scale.to(DEV), add.to(DEV), mask.to(DEV), 1, 1,
                           out_dtype=torch.float32)
    _assert_close(got, want)


WG_CASES = [
    (2, 64, 40, 56, 256, 1, 1, 0), (2, 256, 40, 56, 128, 1, 1, 0), (1, 128, 24, 40, 128, 3, 1, 1),
    (2, 128, 25, 42, 256, 3, 1, 1), (2, 256, 50, 84, 512, 1, 2, 0), (1, 128, 25, 41, 256, 1, 2, 0),
    (1, 24, 20, 30, 48, 3, 1, 1), (1, 256, 30, 40, 81 + 7, 1, 1, 0), (1, 512, 14, 14, 2048, 1, 1, 0),
    (300, 512, 1, 1, 1024, 1, 1, 0), (2, 256, 13, 21, 16, 1, 1, 0),
]


@pytest.mark.parametrize("n,c,h,w,co,k,stride,pad", WG_CASES)
def test_conv_wgrad(ops, n, c, h, w, co, k, stride, pad):
    x, wt = _mk(n, c, h, w, co, k, 9)
    g = torch.Generator().manual_seed(10)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    go = torch.randn(n, co, ho, wo, generator=g).to(torch.bfloat16)
    wf = wt.float().requires_grad_(True)
    F.conv2d(x.float(), wf, stride=stride, padding=pad).backward(go.float())
    got = ops.conv2d_wgrad(x.to(DEV), go.to(DEV), wt.shape, stride, pad)
    assert got.shape == wf.grad.shape
    _assert_close(got, wf.grad, tol=2e-4)   # split-K fp32 atomics: order differs


def test_conv_dgrad_stride2_accumulate_and_mask(ops):
    """First bottleneck of a stage: grad_x = dgrad(conv1, s=2) + dgrad(downsample, s=2), masked by x > 0."""
    n, c, h, w, co1, co2 = 2, 256, 26, 42, 128, 512
    x, w1 = _mk(n, c, h, w, co1, 1, 11)
    _, w2 = _mk(n, c, h, w, co2, 1, 12)
    g = torch.Generator().manual_seed(13)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    g1 = torch.randn(n, co1, ho, wo, generator=g).to(torch.bfloat16)
    g2 = torch.randn(n, co2, ho, wo, generator=g).to(torch.bfloat16)
    xf = x.float().requires_grad_(True)
    (F.conv2d(xf, w1.float(), stride=2) * 1.0).backward(g1.float())
    a = xf.grad.clone().to(torch.bfloat16).float()   # first branch is rounded to bf16 before the second accumulates
    xf.grad = None
    F.conv2d(xf, w2.float(), stride=2).backward(g2.float())
    want = (a + xf.grad) * (x.float() > 0)
    gx = ops.conv2d_dgrad(g1.to(DEV), w1.to(DEV), x.shape, stride=2)
    gx = ops.conv2d_dgrad(g2.to(DEV), w2.to(DEV), x.shape, relu_mask=x.to(DEV), stride=2, accumulate_into=gx)
    _assert_close(gx, want, tol=1e-2)


def test_conv_fwd_residual_upsampled(ops):
    """FPN inner block: conv1x1(C_i) + bias + nearest_upsample_2x(P_{i+1}) with the upsample folded into the epilogue."""
    n, c, h, w, co = 2, 512, 50, 84, 256
    x, wt = _mk(n, c, h, w, co, 1, 21)
    g = torch.Generator().manual_seed(22)
    top = torch.randn(n, co, h // 2, w // 2, generator=g).to(torch.bfloat16)
    bias = torch.randn(co, generator=g)
    want = F.conv2d(x.float(), wt.float()) + bias[None, :, None, None] + F.interpolate(top.float(), scale_factor=2, mode="nearest")
    got = ops.conv2d_fwd(x.to(DEV), wt.to(DEV), None, bias.to(DEV), top.to(DEV), 1, 0, False, torch.float32, residual_up2=True)
    _assert_close(got, want)


def test_wgrad_scale_and_bias_grad(ops):
    n, c, h, w, co, k = 2, 128, 25, 42, 256, 3
    x, wt = _mk(n, c, h, w, co, k, 23)
    g = torch.Generator().manual_seed(24)
    go = torch.randn(n, co, h, w, generator=g).to(torch.bfloat16)
    scale = torch.rand(co, generator=g) + 0.5
    wf = wt.float().requires_grad_(True)
    F.conv2d(x.float(), wf, padding=1).backward(go.float())
    got = ops.conv2d_wgrad(x.to(DEV), go.to(DEV), wt.shape, 1, 1, scale.to(DEV))
    _assert_close(got, wf.grad * scale[:, None, None, None], tol=2e-4)
    _assert_close(ops.bias_grad(go.to(DEV)), go.float().sum((0, 2, 3)), tol=1e-4)
    odd = torch.randn(3, 16, 7, 5, generator=g).to(torch.bfloat16)
    _assert_close(ops.bias_grad(odd.to(DEV)), odd.float().sum((0, 2, 3)), tol=1e-4)


def test_conv_rejects_unsupported(ops):
    x, wt = _mk(1, 64, 16, 16, 64, 3, 8)
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.conv2d_fwd(x.to(DEV), wt.to(DEV), stride=2, pad=1)  # 3x3 stride 2: not in the R-50 hot path
    with pytest.raises(RuntimeError):
        ops.conv2d_fwd(x.float().to(DEV), wt.to(DEV))


def test_conv_rectangular_kernel_and_padding(ops):
    """kh != kw with separate paddings (forward, data gradient, weight gradient)."""
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 64, 19, 27, generator=g).to(torch.bfloat16)
    wt = (torch.randn(128, 64, 3, 1, generator=g) / 14).to(torch.bfloat16)
    go = torch.randn(2, 128, 19, 27, generator=g).to(torch.bfloat16)
    xf, wf = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    y = F.conv2d(xf, wf, padding=(1, 0))
    y.backward(go.float())
    xd, wd, gd = x.to(DEV), wt.to(DEV), go.to(DEV)
    _assert_close(ops.conv2d_fwd(xd, wd, pad=(1, 0), out_dtype=torch.float32), y.detach())
    _assert_close(ops.conv2d_dgrad(gd, wd, x.shape, pad=(1, 0), out_dtype=torch.float32), xf.grad, tol=2e-4)
    _assert_close(ops.conv2d_wgrad(xd, gd, wt.shape, 1, (1, 0)), wf.grad, tol=2e-4)


def test_conv_strided_windows(ops):
    """Operands / results that are strided NHWC windows of larger tensors are used in place (TMA strides for reads,
    epilogue pitches for writes): output rows interleaved into a taller tensor; overlapping 4-pixel input windows."""
    g = torch.Generator().manual_seed(22)
    # (a) write every other row of a [1, 2H, W, 2C]-like buffer, read it back as grad_out for dgrad / wgrad
    n, c, h, w, co = 1, 64, 24, 14, 128
    x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(co, c, 1, 1, generator=g) / 8).to(torch.bfloat16)
    want = F.conv2d(x.float(), wt.float())
    big = torch.full((n, co, 2 * h, w), 7.0, dtype=torch.bfloat16, device=DEV).contiguous(memory_format=torch.channels_last)
    win = torch.as_strided(big, (n, co, h, w), (2 * h * w * co, 1, 2 * w * co, co), w * co)     # the odd rows
    ops.conv2d_fwd(x.to(DEV), wt.to(DEV), out=win)
    _assert_close(big[:, :, 1::2].float(), want.to(torch.bfloat16).float(), tol=1e-2)
    assert bool((big[:, :, 0::2] == 7.0).all())                                                  # untouched rows
    go = torch.randn(n, co, 2 * h, w, generator=g).to(torch.bfloat16)
    god = go.to(DEV).contiguous(memory_format=torch.channels_last)
    gwin = torch.as_strided(god, (n, co, h, w), (2 * h * w * co, 1, 2 * w * co, co), w * co)
    xf, wf = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    F.conv2d(xf, wf).backward(go[:, :, 1::2].float())
    _assert_close(ops.conv2d_dgrad(gwin, wt.to(DEV), x.shape, out_dtype=torch.float32), xf.grad, tol=2e-4)
    _assert_close(ops.conv2d_wgrad(x.to(DEV), gwin, wt.shape), wf.grad, tol=2e-4)
    # (b) overlapping windows: a 1x4 conv over 16 channels == a 1x1 conv over the 64-element window view
    x16 = torch.randn(2, 16, 10, 35, generator=g).to(torch.bfloat16)
    w16 = (torch.randn(32, 16, 1, 4, generator=g) / 8).to(torch.bfloat16)
    want = F.conv2d(x16.float(), w16.float())                                                    # [2, 32, 10, 32]
    xd = x16.to(DEV).contiguous(memory_format=torch.channels_last)
    xv = torch.as_strided(xd, (2, 64, 10, 32), (10 * 35 * 16, 1, 35 * 16, 16))
    wv = w16.to(DEV).permute(0, 2, 3, 1).reshape(32, 64, 1, 1).contiguous(memory_format=torch.channels_last)
    _assert_close(ops.conv2d_fwd(xv, wv, out_dtype=torch.float32), want)


def test_prepare_dgrad_weights_batched(ops):
    """[Cout][taps][Cin] -> [Cin][flipped taps][Cout] (* scale[Cout]) for many layers in one launch: tiled path
    (channels % 8 == 0) and the generic fallback (odd channel counts), > 40 layers to cross a batch boundary."""
    g = torch.Generator().manual_seed(41)
    shapes = [(64, 64, 1), (256, 64, 3), (72, 200, 3), (128, 24, 1), (2048, 512, 1)] + [(64, 32, 3)] * 40 + [(20, 12, 3)]
    ws, scs = [], []
    for i, (co, ci, k) in enumerate(shapes):
        ws.append(torch.randn(co, ci, k, k, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last))
        scs.append((torch.rand(co, generator=g) + 0.5).to(DEV) if i % 2 == 0 else None)
    outs = ops.prepare_dgrad_weights(ws, scs)
    for w, sc, o in zip(ws, scs, outs):
        co, ci, k, _ = w.shape
        wf = w.float() * (sc[:, None, None, None] if sc is not None else 1.0)
        want = wf.flip(2, 3).permute(1, 2, 3, 0).reshape(-1).to(torch.bfloat16)    # [ci][kh'][kw'][co]
        assert torch.equal(o, want), (co, ci, k)


def test_pool_kernels(ops):
    g = torch.Generator().manual_seed(42)
    for (n, c, h, w, k, s, p) in [(2, 64, 37, 50, 3, 2, 1), (1, 256, 25, 42, 1, 2, 0), (1, 8, 9, 9, 2, 2, 0)]:
        x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
        got = ops.max_pool_nhwc(x.to(DEV), k, s, p)
        want = F.max_pool2d(x.float(), k, s, p).to(torch.bfloat16)
        assert got.shape == want.shape and torch.equal(got.cpu(), want)
    for (n, c, h, w) in [(2, 64, 20, 28), (1, 16, 25, 13)]:
        x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
        got = ops.sum_pool2x2_nhwc(x.to(DEV))
        want = F.avg_pool2d(x.float(), 2, ceil_mode=True, count_include_pad=True, divisor_override=1)
        assert got.shape == want.shape
        _assert_close(got, want, tol=1e-2)


def test_bias_grad_shapes(ops):
    g = torch.Generator().manual_seed(43)
    for (n, c, h, w) in [(2, 256, 50, 84), (1024, 1024, 1, 1), (3, 408, 7, 5), (2, 8, 33, 17), (1, 6, 9, 4)]:
        go = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
        _assert_close(ops.bias_grad(go.to(DEV)), go.float().sum((0, 2, 3)), tol=2e-4)


def test_conv_full_size_properties(ops):
    """BASELINE-size layers (2 x 256 x 200 x 336, the P2 plane of an 800x1344 batch) through size-independent
    properties that are EXACT for bf16 operands with fp32 accumulation: identity / delta kernels reproduce or shift the
    input (forward and data gradient), and the weight gradient satisfies a checksum-of-checksums identity."""
    g = torch.Generator().manual_seed(51)
    n, c, h, w = 2, 256, 200, 336
    x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    eye = torch.eye(c)
    w1 = eye.view(c, c, 1, 1).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.conv2d_fwd(x, w1), x)                                   # 1x1 identity (flattened GEMM path)
    d3 = torch.zeros(c, c, 3, 3)
    d3[:, :, 1, 1] = eye
    d3 = d3.to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.conv2d_fwd(x, d3, pad=1), x)                            # centre tap: identity
    assert torch.equal(ops.conv2d_dgrad(x, d3, x.shape, pad=1), x)                 # and its data gradient
    s3 = torch.zeros(c, c, 3, 3)
    s3[:, :, 0, 0] = eye                                                           # y[h, w] = x[h-1, w-1], zero outside
    s3 = s3.to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    want = F.pad(x, (1, 0, 1, 0))[:, :, :h, :w]
    assert torch.equal(ops.conv2d_fwd(x, s3, pad=1), want)                         # im2col-in-TMA coordinates + zero fill
    # weight gradient: sum over (co, ci) of dW[:, :, r, q] == sum over pixels of (sum_co g)(sum_ci x shifted by the tap)
    go = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    dw = ops.conv2d_wgrad(x, go, (c, c, 3, 3), 1, 1)
    gs, xs = go.float().sum(1).double(), x.float().sum(1).double()
    xs = F.pad(xs, (1, 1, 1, 1))
    for r in range(3):
        for q in range(3):
            want_s = float((gs * xs[:, r:r + h, q:q + w]).sum())
            got_s = float(dw[:, :, r, q].double().sum())
            scale = float((gs.abs() * xs[:, r:r + h, q:q + w].abs()).sum())
            assert abs(got_s - want_s) <= 1e-5 * scale, (r, q, got_s, want_s)


# ---------------------------------------------------------------- tile plans (MRB_CONV_TILE override: th,tw,bn,epi,sets)
TILE_PLANS = ["10,12,128,1,2", "10,12,128,1,3", "7,18,64,1,2", "5,25,256,0", "3,42,128,1,3", "9,14,128,1,2", "1,128,128,1,3", "4,32,256,0"]


@pytest.mark.parametrize("plan", TILE_PLANS)
def test_conv_tile_plans_fwd_and_dgrad(ops, monkeypatch, plan):
    """Every tile rectangle (th x tw <= 128 rows of the M = 128 tile), N tile, epilogue mode and buffer-ring depth the
    planner may choose gives the same result: forward with BN + residual + ReLU, data gradient with add + ReLU mask."""
    monkeypatch.setenv("MRB_CONV_TILE", plan)
    n, c, h, w, co = 2, 128, 25, 42, 256
    x, wt = _mk(n, c, h, w, co, 3, 11)
    g = torch.Generator().manual_seed(12)
    scale, bias = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
    conv = F.conv2d(x.float(), wt.float(), padding=1)
    res = torch.randn(conv.shape, generator=g).to(torch.bfloat16)
    want = torch.relu(conv * scale[None, :, None, None] + bias[None, :, None, None] + res.float())
    got = ops.conv2d_fwd(x.to(DEV), wt.to(DEV), scale.to(DEV), bias.to(DEV), res.to(DEV), 1, 1, relu=True)
    _assert_close(got, want, tol=1e-2)
    go = torch.randn(n, co, h, w, generator=g).to(torch.bfloat16)
    add = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
    mask = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
    xr = x.float().requires_grad_(True)
    F.conv2d(xr, wt.float(), padding=1).backward(go.float())
    want_gx = (xr.grad + add.float()) * (mask.float() > 0)
    gx = ops.conv2d_dgrad(go.to(DEV), wt.to(DEV), (n, c, h, w), None, add.to(DEV), mask.to(DEV), 1, 1)
    _assert_close(gx, want_gx, tol=1e-2)
    if plan.startswith("1,128"):
        # the flattened 1x1 path with the same plan
        x1, w1 = _mk(n, c, h, w, co, 1, 13)
        want1 = torch.relu(F.conv2d(x1.float(), w1.float()) * scale[None, :, None, None] + bias[None, :, None, None] + res.float())
        got1 = ops.conv2d_fwd(x1.to(DEV), w1.to(DEV), scale.to(DEV), bias.to(DEV), res.to(DEV), 1, 0, relu=True)
        _assert_close(got1, want1, tol=1e-2)


def test_conv_random_weights_full_size_p2(ops):
    """Random-weight BASELINE-size cases against fp32 F.conv2d computed on the host: 3x3 256->256 and 1x1 64->256 on the
    2 x 200 x 336 P2 plane (fused BN + ReLU), sampled rows (the full fp32 reference of the 3x3 takes ~20 s of CPU)."""
    g = torch.Generator().manual_seed(21)
    for (c, co, k) in ((256, 256, 3), (64, 256, 1)):
        x = torch.randn(2, c, 200, 336, generator=g).to(torch.bfloat16)
        wt = (torch.randn(co, c, k, k, generator=g) / (c * k * k) ** 0.5).to(torch.bfloat16)
        scale, bias = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
        got = ops.conv2d_fwd(x.to(DEV), wt.to(DEV), scale.to(DEV), bias.to(DEV), None, 1, k // 2, relu=True,
                             out_dtype=torch.float32).cpu()
        for (r0, r1) in ((0, 9), (95, 106), (191, 200)):          # top edge, interior, bottom edge (with halo)
            lo, hi = max(r0 - k // 2, 0), min(r1 + k // 2, 200)
            ref = F.conv2d(x[:, :, lo:hi].float(), wt.float(), padding=(0, k // 2))
            if k == 3:
                # rows of `ref` are valid convolution rows lo+1 .. hi-2; pad the image borders explicitly
                xp = F.pad(x[:, :, lo:hi].float(), (0, 0, 1 if lo == 0 else 0, 1 if hi == 200 else 0))
                ref = F.conv2d(xp, wt.float(), padding=(0, 1))
                first = lo + 1 - (1 if lo == 0 else 0)
            else:
                first = lo
            ref = torch.relu(ref * scale[None, :, None, None] + bias[None, :, None, None])
            a, b = max(r0, first), min(r1, first + ref.shape[2])
            _assert_close(got[:, :, a:b], ref[:, :, a - first:b - first])
