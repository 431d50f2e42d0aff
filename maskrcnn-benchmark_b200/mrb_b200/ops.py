"""Host-side Python API over the C ABI for the pieces that are not part of the reference's `_C`
surface: the tcgen05 conv engine (forward / data-gradient) and the fused multi-level ROIAlign.

Tensors are logical NCHW in torch.channels_last memory format (== NHWC in memory), bf16 for conv
operands.  Everything here raises if libmrb_b200.so or a CUDA device is missing -- no fallback."""
import ctypes

import torch

from maskrcnn_benchmark import _C as _c

lib = _c.lib
lib.mrb_conv2d_dgrad_workspace_bytes.restype = ctypes.c_size_t
lib.mrb_conv2d_dgrad_workspace_bytes.argtypes = [ctypes.POINTER(_c.ConvParams)]

_DT = {torch.float32: 0, torch.bfloat16: 1}

# launch accounting for bench.py: number of kernels of libmrb_b200.so launched through this module, and
# (when a list is installed) the geometry of every tcgen05 conv launch
STATS = {"launches": 0, "conv_calls": None}


def _count(n, conv_key=None):
    STATS["launches"] += n
    if conv_key is not None and STATS["conv_calls"] is not None:
        STATS["conv_calls"].append(conv_key)


def _kk(p):
    return p.kh if p.kh == p.kw else (p.kh, p.kw)


def _pp(p):
    return p.pad if not p.flags else (p.pad, p.pad_w)


def _nhwc(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s: expected a CUDA tensor (no CPU path)" % name)
    if t.dim() != 4:
        raise RuntimeError("%s: expected a 4-D tensor" % name)
    if not t.is_contiguous(memory_format=torch.channels_last):
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def _view(t, name):
    """NHWC operand: returns (tensor, pitch).  Dense channels_last tensors have pitch None; a strided window of a
    larger NHWC tensor (channel stride 1, other strides multiples of 8 elements) is passed in place with its
    {image, row, pixel} element strides; anything else is copied to dense channels_last."""
    if not t.is_cuda:
        raise RuntimeError("%s: expected a CUDA tensor (no CPU path)" % name)
    if t.dim() != 4:
        raise RuntimeError("%s: expected a 4-D tensor" % name)
    if t.is_contiguous(memory_format=torch.channels_last):
        return t, None
    n, _, h, _ = t.shape
    sn, sc, sh, sw = t.stride()
    if sc == 1 and sh % 8 == 0 and sw % 8 == 0 and sh > 0 and sw > 0 and (n == 1 or (sn % 8 == 0 and sn > 0)) \
            and t.data_ptr() % 16 == 0:
        return t, (sn if n > 1 else sh * h, sh, sw)
    return t.contiguous(memory_format=torch.channels_last), None


FLAG_PAD_W, FLAG_GROUPED64 = 1, 2          # mrb_conv_params.flags (include/mrb_b200.h)


_PARAMS = {}


def _params(x_shape, w_shape, stride, pad, relu, out_dtype, out_hw=None, x_pitch=None, y_pitch=None, grouped=False):
    """mrb_conv_params of a launch (+ output size); memoised per geometry -- the structs are never mutated after creation."""
    key = (tuple(x_shape), tuple(w_shape), stride, pad, bool(relu), out_dtype, None if out_hw is None else tuple(out_hw),
           x_pitch, y_pitch, grouped)
    hit = _PARAMS.get(key)
    if hit is None:
        hit = _PARAMS[key] = _make_params(x_shape, w_shape, stride, pad, relu, out_dtype, out_hw, x_pitch, y_pitch, grouped)
    return hit


def _make_params(x_shape, w_shape, stride, pad, relu, out_dtype, out_hw=None, x_pitch=None, y_pitch=None, grouped=False):
    n, c, h, w = x_shape
    co, ci, kh, kw = w_shape
    if grouped:
        if ci != 64 or co != c or c % 64:
            raise RuntimeError("grouped conv2d: expanded weight [C, 64, kh, kw] with C == Cin %% 64 == 0 expected")
    elif ci != c:
        raise RuntimeError("conv2d: weight expects %d input channels, got %d" % (ci, c))
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    p = _c.ConvParams()
    p.batch, p.height, p.width, p.cin = n, h, w, c
    p.cout, p.kh, p.kw = co, kh, kw
    p.stride, p.pad, p.relu, p.out_dtype = stride, ph, int(bool(relu)), _DT[out_dtype]
    if pw != ph:
        p.pad_w, p.flags = pw, FLAG_PAD_W
    if grouped:
        p.flags |= FLAG_GROUPED64
    ho = (h + 2 * ph - kh) // stride + 1
    wo = (w + 2 * pw - kw) // stride + 1
    if out_hw is not None:
        p.out_h, p.out_w = ho, wo = int(out_hw[0]), int(out_hw[1])
    if x_pitch is not None:
        p.x_pitch[0], p.x_pitch[1], p.x_pitch[2] = x_pitch
    if y_pitch is not None:
        p.y_pitch[0], p.y_pitch[1], p.y_pitch[2] = y_pitch
    return p, ho, wo


def conv2d_fwd(x, weight, scale=None, bias=None, residual=None, stride=1, pad=0, relu=False,
               out_dtype=torch.bfloat16, out_hw=None, residual_up2=False, out=None, grouped=False):
    """y = act(conv(x, weight) * scale[c] + bias[c] + residual).  x, weight bf16 channels_last.
    residual_up2: `residual` has half the output resolution and is read through a nearest 2x upsample.
    pad: int or (pad_h, pad_w).  x may be a strided NHWC window (see _view); `out` (optional) is a preallocated
    result, possibly a strided window of a larger tensor, written in place."""
    x, x_pitch = _view(x, "conv2d_fwd(x)")
    weight = _nhwc(weight, "conv2d_fwd(weight)")
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise RuntimeError("conv2d_fwd: bf16 operands required")
    y_pitch = None
    if out is not None:
        out_v, y_pitch = _view(out, "conv2d_fwd(out)")
        if out_v is not out:
            raise RuntimeError("conv2d_fwd: `out` must be NHWC (channel stride 1, other strides multiples of 8)")
        out_dtype = out.dtype
        if residual is not None and y_pitch is not None:
            raise RuntimeError("conv2d_fwd: residual with a strided `out` is not supported")
    p, ho, wo = _params(x.shape, weight.shape, stride, pad, relu, out_dtype, out_hw, x_pitch, y_pitch, grouped)
    if out is None:
        out = torch.empty((p.batch, p.cout, ho, wo), dtype=out_dtype, device=x.device, memory_format=torch.channels_last)
    elif tuple(out.shape) != (p.batch, p.cout, ho, wo):
        raise RuntimeError("conv2d_fwd: out shaped %s, expected %s" % (tuple(out.shape), (p.batch, p.cout, ho, wo)))
    if residual is not None:
        residual = _nhwc(residual, "conv2d_fwd(residual)")
        want = (p.batch, p.cout, (ho + 1) // 2, (wo + 1) // 2) if residual_up2 else tuple(out.shape)
        if residual.dtype != torch.bfloat16 or tuple(residual.shape) != want:
            raise RuntimeError("conv2d_fwd: residual must be bf16 and shaped %s" % (want,))
    for v in (scale, bias):
        if v is not None and (v.dtype != torch.float32 or v.numel() != p.cout or not v.is_contiguous()):
            raise RuntimeError("conv2d_fwd: scale/bias must be contiguous fp32 [Cout]")
    with _c.on_device(x.device):
        fn = lib.mrb_conv2d_fwd_up2 if residual_up2 else lib.mrb_conv2d_fwd
        _c.check(fn(ctypes.byref(p), _c._ptr(x), _c._ptr(weight), _c._ptr(scale), _c._ptr(bias),
                    _c._ptr(residual), _c._ptr(out), _c._stream()), "mrb_conv2d_fwd")
    _count(1, ("fwd", p.batch, p.cin, p.height, p.width, p.cout, _kk(p), p.stride, _pp(p)) + ((p.cin // 64,) if grouped else ()))
    return out


# ------------------------------------------------------------------------- grouped conv (block-diagonal super-groups)
def grouped_dgrad_weights(w_exp, scale=None):
    """Expanded grouped filter [C, 64, kh, kw] (row co: its filter against the 64 input channels of co's super-group)
    -> flat bf16 [Cin][kh][kw][64] for the data gradient: entry (ci, tap, k) = w_exp[64*(ci//64) + k, ci % 64, flipped tap]
    (* scale[co], the frozen-BN scale of the forward epilogue)."""
    c, sg, kh, kw = w_exp.shape
    w = w_exp.float().view(c // 64, 64, sg, kh, kw)                   # [B, co_l, ci_l, kh, kw]
    if scale is not None:
        w = w * scale.float().view(c // 64, 64, 1, 1, 1)
    w = w.flip(3, 4).permute(0, 2, 3, 4, 1)                           # [B, ci_l, kh', kw', co_l]
    return w.reshape(-1).to(torch.bfloat16).contiguous()


def grouped_expand_weights(w16, groups, scale=None, fwd=True, dgrad=False):
    """One launch: the block-diagonal expansions of a grouped filter.  w16: bf16 [C, C/groups, kh, kw] channels_last (KRSC).
    -> (w_exp [C, 64, kh, kw] channels_last bf16 or None, wd_flat bf16 [C*kh*kw*64] or None); see csrc/grouped_prep.cu."""
    w16 = _nhwc(w16, "grouped_expand_weights(w16)")
    if w16.dtype != torch.bfloat16:
        raise RuntimeError("grouped_expand_weights: bf16 weight required")
    c, cg, kh, kw = w16.shape
    w_exp = torch.empty((c, 64, kh, kw), dtype=torch.bfloat16, device=w16.device).contiguous(memory_format=torch.channels_last) if fwd else None
    wd = torch.empty(c * kh * kw * 64, dtype=torch.bfloat16, device=w16.device) if dgrad else None
    with _c.on_device(w16.device):
        _c.check(lib.mrb_grouped_expand_weights(_c._ptr(w16), _c._ptr(scale), _c._ptr(w_exp), _c._ptr(wd), c, kh * kw, groups,
                                                _c._stream()), "mrb_grouped_expand_weights")
    _count(1)
    return w_exp, wd


def grouped_collapse_wgrad(gw128, w_shape, groups, accumulate_into=None):
    """Expanded weight gradient [C, taps, 128] fp32 (conv2d_wgrad_grouped_raw) -> the grouped layout [C, C/groups, kh, kw]
    (channels_last memory), or ADDED into `accumulate_into` (any dense layout of that shape)."""
    c, cg, kh, kw = w_shape
    if accumulate_into is not None:
        out, acc = accumulate_into, 1
        if out.dtype != torch.float32 or tuple(out.shape) != tuple(w_shape):
            raise RuntimeError("grouped_collapse_wgrad: accumulate_into must be fp32 shaped like the weight")
    else:
        out, acc = torch.empty(tuple(w_shape), dtype=torch.float32, device=gw128.device).contiguous(memory_format=torch.channels_last), 0
    s = out.stride()
    with _c.on_device(gw128.device):
        # tap stride: taps are enumerated (r, q) row-major; a dense [.., kh, kw] block has stride(q) = s[3], stride(r) = s[2] = kw * s[3]
        if s[2] != kw * s[3]:
            raise RuntimeError("grouped_collapse_wgrad: unsupported weight layout")
        _c.check(lib.mrb_grouped_collapse_wgrad(_c._ptr(gw128), _c._ptr(out), c, kh * kw, groups, ctypes.c_longlong(s[0]),
                                                ctypes.c_longlong(s[1]), ctypes.c_longlong(s[3]), acc, _c._stream()),
                 "mrb_grouped_collapse_wgrad")
    _count(1)
    return out


def conv2d_dgrad_grouped(grad_out, wd_flat, x_shape, stride=1, pad=1, add=None, relu_mask=None):
    """Data gradient of a grouped (MRB_CONV_GROUPED64) convolution from the prepared weights of grouped_dgrad_weights."""
    if stride != 1:
        raise RuntimeError("conv2d_dgrad_grouped: stride 1 only")
    grad_out = _nhwc(grad_out, "conv2d_dgrad_grouped(grad_out)")
    n, c, h, w = x_shape
    k2 = wd_flat.numel() // (c * 64)
    k = int(round(k2 ** 0.5))
    p, ho, wo = _params(tuple(x_shape), (c, 64, k, k), 1, pad, False, torch.bfloat16, None, None, None, True)
    if tuple(grad_out.shape) != (n, c, ho, wo) or grad_out.dtype != torch.bfloat16 or wd_flat.dtype != torch.bfloat16:
        raise RuntimeError("conv2d_dgrad_grouped: bad grad_out / weights")
    gx = torch.empty(tuple(x_shape), dtype=torch.bfloat16, device=grad_out.device, memory_format=torch.channels_last)
    for t in (add, relu_mask):
        if t is not None and (t.dtype != torch.bfloat16 or tuple(t.shape) != tuple(x_shape)):
            raise RuntimeError("conv2d_dgrad_grouped: add/relu_mask must be bf16 and shaped like x")
    add = _nhwc(add, "add") if add is not None else None
    relu_mask = _nhwc(relu_mask, "relu_mask") if relu_mask is not None else None
    with _c.on_device(grad_out.device):
        _c.check(lib.mrb_conv2d_dgrad_prepared(ctypes.byref(p), _c._ptr(grad_out), _c._ptr(wd_flat), _c._ptr(add),
                                               _c._ptr(relu_mask), _c._ptr(gx), _c._stream()), "mrb_conv2d_dgrad_prepared(grouped)")
    _count(1, ("dgrad", n, c, h, w, c, k, 1, pad, c // 64))
    return gx


def conv2d_wgrad_grouped(x, grad_out, k, stride=1, pad=1, scale=None, raw=False):
    """Weight gradient of a grouped convolution as the EXPANDED fp32 [C, 64, k, k] (row co against the 64 input channels
    of its super-group): the kernel produces each 128-channel Cout tile against its own 128 input channels, the two
    diagonal 64 x 64 blocks are kept.  raw=True returns the kernel's [C, k*k, 128] buffer (for grouped_collapse_wgrad)."""
    if stride != 1:
        raise RuntimeError("conv2d_wgrad_grouped: stride 1 only")
    x = _nhwc(x, "conv2d_wgrad_grouped(x)")
    grad_out = _nhwc(grad_out, "conv2d_wgrad_grouped(grad_out)")
    n, c, h, w = x.shape
    p, ho, wo = _params(x.shape, (c, 64, k, k), 1, pad, False, torch.float32, None, None, None, True)
    if tuple(grad_out.shape) != (n, c, ho, wo) or c % 128:
        raise RuntimeError("conv2d_wgrad_grouped: grad_out shape mismatch or C %% 128 != 0")
    gw = torch.empty((c, k * k, 128), dtype=torch.float32, device=x.device)
    with _c.on_device(x.device):
        _c.check(lib.mrb_conv2d_wgrad(ctypes.byref(p), _c._ptr(x), _c._ptr(grad_out), _c._ptr(scale), _c._ptr(gw), _c._stream()),
                 "mrb_conv2d_wgrad(grouped)")
    _count(1, ("wgrad", n, c, h, w, c, k, 1, pad, c // 64))
    if raw:
        return gw
    gw = gw.view(c // 128, 2, 64, k * k, 2, 64)                       # [tile, half, co_l, tap, half', ci_l]
    diag = torch.stack([gw[:, 0, :, :, 0], gw[:, 1, :, :, 1]], 1)     # [tile, half, co_l, tap, ci_l]
    return diag.reshape(c, k, k, 64).permute(0, 3, 1, 2)              # logical [C, 64, k, k]


def prepare_dgrad_weights(weights, scales, out=None):
    """Flipped/transposed (+ per-Cout scaled) bf16 copies of many conv weights in one launch.
    weights: list of bf16 channels_last [Cout,Cin,kh,kw]; scales: list of fp32 [Cout] or None.  Returns the list of
    prepared tensors (flat bf16, Cout*kh*kw*Cin elements), reusing `out` when given."""
    n = len(weights)
    if out is None:
        out = [torch.empty(w.numel(), dtype=torch.bfloat16, device=w.device) for w in weights]
    if n == 0:
        return out
    ws = [_nhwc(w, "prepare_dgrad_weights") for w in weights]
    wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
    sp = (ctypes.c_void_p * n)(*[(s.data_ptr() if s is not None else None) for s in scales])
    op = (ctypes.c_void_p * n)(*[o.data_ptr() for o in out])
    co = (ctypes.c_int * n)(*[w.shape[0] for w in ws])
    tp = (ctypes.c_int * n)(*[w.shape[2] * w.shape[3] for w in ws])
    ci = (ctypes.c_int * n)(*[w.shape[1] for w in ws])
    with _c.on_device(ws[0].device):
        _c.check(lib.mrb_conv2d_prepare_dgrad_weights(n, wp, sp, op, co, tp, ci, _c._stream()), "mrb_conv2d_prepare_dgrad_weights")
    _count((n + 39) // 40)
    return out


def conv2d_dgrad(grad_out, weight, x_shape, scale=None, add=None, relu_mask=None, stride=1, pad=0,
                 out_dtype=torch.bfloat16, accumulate_into=None, prepared=None):
    """grad_x = conv_transpose(grad_out, weight * scale[cout]) (+ add) masked by (relu_mask > 0).
    `accumulate_into` (bf16, shaped like x) makes the call in-place: result = accumulate_into + dgrad,
    written back into it (the only form of `add` the stride-2 path supports)."""
    grad_out, y_pitch = _view(grad_out, "conv2d_dgrad(grad_out)")
    weight = _nhwc(weight, "conv2d_dgrad(weight)")
    if grad_out.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise RuntimeError("conv2d_dgrad: bf16 operands required")
    p, ho, wo = _params(tuple(x_shape), weight.shape, stride, pad, False, out_dtype, None, None, y_pitch)
    if tuple(grad_out.shape) != (p.batch, p.cout, ho, wo):
        raise RuntimeError("conv2d_dgrad: grad_out shape %s != %s" % (tuple(grad_out.shape), (p.batch, p.cout, ho, wo)))
    if accumulate_into is not None:
        if add is not None:
            raise RuntimeError("conv2d_dgrad: pass either add or accumulate_into")
        gx = accumulate_into
        if gx.dtype != torch.bfloat16 or tuple(gx.shape) != tuple(x_shape) or \
                not gx.is_contiguous(memory_format=torch.channels_last):
            raise RuntimeError("conv2d_dgrad: accumulate_into must be bf16 channels_last shaped like x")
        add = gx
    else:
        gx = torch.empty(tuple(x_shape), dtype=out_dtype, device=grad_out.device, memory_format=torch.channels_last)
    for t in (add, relu_mask):
        if t is not None and (t.dtype != torch.bfloat16 or tuple(t.shape) != tuple(x_shape)):
            raise RuntimeError("conv2d_dgrad: add/relu_mask must be bf16 and shaped like x")
    if add is not None and add is not gx:
        add = _nhwc(add, "add")
    relu_mask = _nhwc(relu_mask, "relu_mask") if relu_mask is not None else None
    with _c.on_device(grad_out.device):
        if prepared is not None:
            # `prepared` already holds the flipped/transposed/scaled weights (prepare_dgrad_weights)
            _c.check(lib.mrb_conv2d_dgrad_prepared(ctypes.byref(p), _c._ptr(grad_out), _c._ptr(prepared), _c._ptr(add),
                                                   _c._ptr(relu_mask), _c._ptr(gx), _c._stream()), "mrb_conv2d_dgrad_prepared")
            _count(1, ("dgrad", p.batch, p.cin, p.height, p.width, p.cout, _kk(p), p.stride, _pp(p)))
            return gx
        nbytes = lib.mrb_conv2d_dgrad_workspace_bytes(ctypes.byref(p))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=grad_out.device)
        _c.check(lib.mrb_conv2d_dgrad(ctypes.byref(p), _c._ptr(grad_out), _c._ptr(weight), _c._ptr(scale), _c._ptr(add),
                                      _c._ptr(relu_mask), _c._ptr(gx), _c._ptr(ws), ctypes.c_size_t(nbytes),
                                      _c._stream()), "mrb_conv2d_dgrad")
    _count(2, ("dgrad", p.batch, p.cin, p.height, p.width, p.cout, _kk(p), p.stride, _pp(p)))
    return gx


def conv2d_wgrad(x, grad_out, w_shape, stride=1, pad=0, scale=None, accumulate_into=None):
    """grad_weight (fp32, logical [Cout, Cin, kh, kw], channels_last memory == KRSC) on the tcgen05 engine,
    optionally multiplied by a per-Cout `scale` (the frozen-BN scale of the forward epilogue).
    `accumulate_into` (fp32, that shape and memory order): the result is ADDED into it (red.add) instead of being
    returned in a fresh zero-filled tensor."""
    x, x_pitch = _view(x, "conv2d_wgrad(x)")
    grad_out, y_pitch = _view(grad_out, "conv2d_wgrad(grad_out)")
    if x.dtype != torch.bfloat16 or grad_out.dtype != torch.bfloat16:
        raise RuntimeError("conv2d_wgrad: bf16 operands required")
    p, ho, wo = _params(x.shape, tuple(w_shape), stride, pad, False, torch.float32, None, x_pitch, y_pitch)
    if tuple(grad_out.shape) != (p.batch, p.cout, ho, wo):
        raise RuntimeError("conv2d_wgrad: grad_out shape mismatch")
    if accumulate_into is not None:
        gw = accumulate_into
        if gw.dtype != torch.float32 or tuple(gw.shape) != tuple(w_shape) or \
                not gw.is_contiguous(memory_format=torch.channels_last):
            raise RuntimeError("conv2d_wgrad: accumulate_into must be fp32 channels_last shaped like the weight")
        fn = lib.mrb_conv2d_wgrad_accumulate
    else:
        gw = torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device).contiguous(memory_format=torch.channels_last)
        fn = lib.mrb_conv2d_wgrad
    with _c.on_device(x.device):
        _c.check(fn(ctypes.byref(p), _c._ptr(x), _c._ptr(grad_out), _c._ptr(scale), _c._ptr(gw), _c._stream()),
                 "mrb_conv2d_wgrad")
    _count(1, ("wgrad", p.batch, p.cin, p.height, p.width, p.cout, _kk(p), p.stride, _pp(p)))
    return gw


def bias_grad(grad_out, accumulate_into=None):
    """sum over N, H, W of an NHWC bf16 gradient -> fp32 [C] (added into `accumulate_into` when given)."""
    grad_out = _nhwc(grad_out, "bias_grad(grad_out)")
    if grad_out.dtype != torch.bfloat16:
        raise RuntimeError("bias_grad: bf16 gradient required")
    n, c, h, w = grad_out.shape
    if accumulate_into is not None:
        out, fn = accumulate_into, lib.mrb_bias_grad_accumulate
        if out.dtype != torch.float32 or out.numel() != c or not out.is_contiguous():
            raise RuntimeError("bias_grad: accumulate_into must be contiguous fp32 [C]")
    else:
        out, fn = torch.empty(c, dtype=torch.float32, device=grad_out.device), lib.mrb_bias_grad
    with _c.on_device(grad_out.device):
        _c.check(fn(_c._ptr(grad_out), _c._ptr(out), ctypes.c_longlong(n * h * w), c, _c._stream()), "mrb_bias_grad")
    _count(1)
    return out


def max_pool_nhwc(x, kernel, stride, pad):
    """F.max_pool2d forward on an NHWC bf16 tensor (no indices: for inputs that are not differentiated)."""
    x = _nhwc(x, "max_pool_nhwc")
    if x.dtype != torch.bfloat16:
        raise RuntimeError("max_pool_nhwc: bf16 input required")
    n, c, h, w = x.shape
    ho, wo = (h + 2 * pad - kernel) // stride + 1, (w + 2 * pad - kernel) // stride + 1
    out = torch.empty((n, c, ho, wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with _c.on_device(x.device):
        _c.check(lib.mrb_max_pool_nhwc(_c._ptr(x), _c._ptr(out), n, h, w, c, kernel, stride, pad, _c._stream()), "mrb_max_pool_nhwc")
    _count(1)
    return out


def sum_pool2x2_nhwc(g):
    """Sum over 2x2 blocks (ceil-sized output) of an NHWC bf16 tensor: backward of the nearest-2x upsample."""
    g = _nhwc(g, "sum_pool2x2_nhwc")
    if g.dtype != torch.bfloat16:
        raise RuntimeError("sum_pool2x2_nhwc: bf16 input required")
    n, c, h, w = g.shape
    out = torch.empty((n, c, (h + 1) // 2, (w + 1) // 2), dtype=g.dtype, device=g.device, memory_format=torch.channels_last)
    with _c.on_device(g.device):
        _c.check(lib.mrb_sum_pool2x2_nhwc(_c._ptr(g), _c._ptr(out), n, h, w, c, _c._stream()), "mrb_sum_pool2x2_nhwc")
    _count(1)
    return out


def sgd_momentum_step(param, grad, momentum_buf, param_bf16, lr, momentum, weight_decay, grad_scale=1.0, zero_grad=True):
    """In-place fused SGD update of flat fp32 tensors (see mrb_sgd_momentum_step): also refreshes the bf16 operand
    copy `param_bf16` (may be None) and zeroes `grad`."""
    n = param.numel()
    for t in (param, grad, momentum_buf):
        if t.dtype != torch.float32 or t.numel() != n or not t.is_contiguous() or not t.is_cuda:
            raise RuntimeError("sgd_momentum_step: param/grad/momentum_buf must be contiguous fp32 CUDA tensors of one size")
    if param_bf16 is not None and (param_bf16.dtype != torch.bfloat16 or param_bf16.numel() != n or not param_bf16.is_contiguous()):
        raise RuntimeError("sgd_momentum_step: param_bf16 must be a contiguous bf16 tensor of the same size")
    with _c.on_device(param.device):
        _c.check(lib.mrb_sgd_momentum_step(_c._ptr(param), _c._ptr(grad), _c._ptr(momentum_buf), _c._ptr(param_bf16),
                                           ctypes.c_longlong(n), ctypes.c_float(lr), ctypes.c_float(momentum),
                                           ctypes.c_float(weight_decay), ctypes.c_float(grad_scale), int(bool(zero_grad)),
                                           _c._stream()), "mrb_sgd_momentum_step")
    _count(1)


# ------------------------------------------------------------------------- deformable conv (NHWC bf16, tensor-core path)
def _dcn_geom(x, om, k, stride, pad, dil, modulated):
    n, c, h, w = x.shape
    ho = (h + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    wo = (w + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    if om.dtype != torch.float32 or om.dim() != 4 or tuple(om.shape[0:1] + om.shape[2:]) != (n, ho, wo) or \
            not om.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("dcn: offset/mask tensor must be fp32 channels_last [N, >= %d, %d, %d]" % (k * k * (3 if modulated else 2), ho, wo))
    if om.shape[1] < k * k * (3 if modulated else 2):
        raise RuntimeError("dcn: offset/mask tensor has too few channels")
    return n, c, h, w, ho, wo


def dcn_sample_nhwc(x, om, k=3, stride=1, pad=1, dil=1, modulated=False):
    """Deformable im2col on NHWC bf16: -> cols, logical [N, k*k*C, Ho, Wo] channels_last (memory [pix][tap*C + c]).
    om: fp32 channels_last [N, OC, Ho, Wo]: channels (2t, 2t+1) = offsets of tap t, (2*k*k + t) = mask logit (v2)."""
    x = _nhwc(x, "dcn_sample_nhwc(x)")
    if x.dtype != torch.bfloat16:
        raise RuntimeError("dcn_sample_nhwc: bf16 input required")
    n, c, h, w, ho, wo = _dcn_geom(x, om, k, stride, pad, dil, modulated)
    cols = torch.empty((n, k * k * c, ho, wo), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    with _c.on_device(x.device):
        _c.check(lib.mrb_dcn_sample_nhwc(_c._ptr(x), _c._ptr(om), _c._ptr(cols), n, h, w, c, ho, wo, k, k, stride, pad, dil,
                                         om.shape[1], int(bool(modulated)), _c._stream()), "mrb_dcn_sample_nhwc")
    _count(1)
    return cols


def dcn_backward_nhwc(x, om, gcols, k=3, stride=1, pad=1, dil=1, modulated=False, need_grad_x=True):
    """-> (grad_x fp32 NHWC [N, C, H, W] logical or None, grad_om fp32 channels_last shaped like om; unused channels 0)."""
    x = _nhwc(x, "dcn_backward_nhwc(x)")
    gcols = _nhwc(gcols, "dcn_backward_nhwc(gcols)")
    n, c, h, w, ho, wo = _dcn_geom(x, om, k, stride, pad, dil, modulated)
    if gcols.dtype != torch.bfloat16 or tuple(gcols.shape) != (n, k * k * c, ho, wo):
        raise RuntimeError("dcn_backward_nhwc: gcols must be bf16 [N, k*k*C, Ho, Wo]")
    gx = torch.zeros((n, c, h, w), dtype=torch.float32, device=x.device).contiguous(memory_format=torch.channels_last) if need_grad_x else None
    gom = torch.zeros_like(om)
    with _c.on_device(x.device):
        _c.check(lib.mrb_dcn_backward_nhwc(_c._ptr(x), _c._ptr(om), _c._ptr(gcols), _c._ptr(gx), _c._ptr(gom), n, h, w, c, ho, wo,
                                           k, k, stride, pad, dil, om.shape[1], int(bool(modulated)), _c._stream()),
                 "mrb_dcn_backward_nhwc")
    _count(1)
    return gx, gom


# ------------------------------------------------------------------------- mask targets from polygons
class PolygonSet:
    """Polygon segmentations of the instances of a batch, packed for mrb_mask_targets_polygons.
    instances: list (one entry per instance) of lists of flat [x0, y0, x1, y1, ...] coordinate sequences."""

    def __init__(self, instances, device):
        xy, pstart, istart = [], [0], [0]
        for polys in instances:
            for p in polys:
                p = [float(v) for v in p]
                if len(p) % 2:
                    raise RuntimeError("PolygonSet: odd number of coordinates")
                xy.extend(p)
                pstart.append(len(xy) // 2)
            istart.append(len(pstart) - 1)
        if not xy:
            xy = [0.0, 0.0]
        # ONE pinned, asynchronous host -> device copy (a pageable torch.tensor(..., device=cuda) blocks the host until the
        # stream has drained): vertices, then the two CSR arrays bit-cast into the same fp32 buffer
        host = torch.empty(len(xy) + len(pstart) + len(istart), dtype=torch.float32).pin_memory() if torch.device(device).type == "cuda" \
            else torch.empty(len(xy) + len(pstart) + len(istart), dtype=torch.float32)
        host[:len(xy)] = torch.tensor(xy, dtype=torch.float32)
        host[len(xy):].view(torch.int32).copy_(torch.tensor(pstart + istart, dtype=torch.int32))
        dev = host.to(device, non_blocking=True)
        self.xy = dev[:len(xy)]
        self.poly_start = dev[len(xy):len(xy) + len(pstart)].view(torch.int32)
        self.inst_start = dev[len(xy) + len(pstart):].view(torch.int32)
        self.num_instances = len(instances)


def mask_targets_polygons(polyset, rois, inst_of_roi, m):
    """[R, m, m] fp32 {0,1}: instance inst_of_roi[r] rasterised on the m x m grid of box rois[r] (xyxy)."""
    rois = rois.float().contiguous()
    idx = inst_of_roi.to(torch.int32).contiguous()
    if not rois.is_cuda:
        raise RuntimeError("mask_targets_polygons: expected CUDA tensors (no CPU path)")
    r = rois.shape[0]
    out = torch.empty((r, m, m), dtype=torch.float32, device=rois.device)
    if r:
        with _c.on_device(rois.device):
            _c.check(lib.mrb_mask_targets_polygons(_c._ptr(polyset.xy), _c._ptr(polyset.poly_start), _c._ptr(polyset.inst_start),
                                                   _c._ptr(rois), _c._ptr(idx), _c._ptr(out), r, m, _c._stream()),
                     "mrb_mask_targets_polygons")
        _count(1)
    return out


def mask_targets_rect(gt_boxes, rois, m):
    """[R, m, m] fp32 {0,1}: the matched ground-truth RECTANGLE gt_boxes[r] cropped to box rois[r] and sampled at the m x m cell
    centres (project_masks_on_boxes for rectangle instances), one launch.  rois: [R, 4] or [R, 5] (image index first)."""
    if not rois.is_cuda:
        raise RuntimeError("mask_targets_rect: expected CUDA tensors (no CPU path)")
    gt_boxes = gt_boxes.float().contiguous()
    rois = rois.float().contiguous()
    r, w = rois.shape
    out = torch.empty((r, m, m), dtype=torch.float32, device=rois.device)
    if r:
        off = 1 if w == 5 else 0
        with _c.on_device(rois.device):
            _c.check(lib.mrb_mask_targets_rect(_c._ptr(gt_boxes), _c_void_p_offset(rois, off), w, _c._ptr(out), r, m, _c._stream()),
                     "mrb_mask_targets_rect")
        _count(1)
    return out


def _c_void_p_offset(t, elems):
    return ctypes.c_void_p(t.data_ptr() + elems * t.element_size())


# ------------------------------------------------------------------------- fused FPN ROIAlign
class _RoiAlignFpn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rois, scales, pooled, sampling_ratio, out_nhwc, k_min, k_max, s0, lvl0, *feats):
        feats = [_nhwc(f, "roi_align_fpn(feat)") for f in feats]
        dt = feats[0].dtype
        if dt not in _DT or any(f.dtype != dt for f in feats):
            raise RuntimeError("roi_align_fpn: feature maps must all be bf16 or all fp32")
        n, c = feats[0].shape[:2]
        r = rois.shape[0]
        rois = rois.float().contiguous()
        L = len(feats)
        out = torch.empty((r, pooled, pooled, c) if out_nhwc else (r, c, pooled, pooled), dtype=dt, device=rois.device)
        hs = (ctypes.c_int * L)(*[f.shape[2] for f in feats])
        ws = (ctypes.c_int * L)(*[f.shape[3] for f in feats])
        sc = (ctypes.c_float * L)(*[float(s) for s in scales])
        ptrs = (ctypes.c_void_p * L)(*[f.data_ptr() for f in feats])
        geom = (L, n, c, pooled, sampling_ratio, k_min, k_max, float(s0), lvl0, _DT[dt], int(bool(out_nhwc)))
        if r > 0:
            with _c.on_device(rois.device):
                _c.check(lib.mrb_roi_align_fpn_fwd(ptrs, hs, ws, sc, L, _c._ptr(rois), _c._ptr(out), r, n, c, pooled,
                                                   sampling_ratio, k_min, k_max, ctypes.c_float(s0), lvl0, _DT[dt],
                                                   int(bool(out_nhwc)), _c._stream()), "mrb_roi_align_fpn_fwd")
            _count(1)
        ctx.save_for_backward(rois)
        ctx.geom = geom
        ctx.shapes = [tuple(f.shape) for f in feats]
        ctx.scales = [float(s) for s in scales]
        ctx.dt = dt
        if out_nhwc:
            out = out.permute(0, 3, 1, 2)  # logical [R, C, P, P], channels_last memory
        return out

    @staticmethod
    def backward(ctx, gout):
        (rois,) = ctx.saved_tensors
        L, n, c, pooled, sr, k_min, k_max, s0, lvl0, dtc, out_nhwc = ctx.geom
        dt = ctx.dt
        if out_nhwc:
            gout = gout.permute(0, 2, 3, 1)
        gout = gout.to(dt).contiguous()
        grads = [torch.zeros((s[0], s[2], s[3], s[1]), dtype=torch.float32, device=rois.device) for s in ctx.shapes]
        r = rois.shape[0]
        if r > 0:
            hs = (ctypes.c_int * L)(*[s[2] for s in ctx.shapes])
            ws = (ctypes.c_int * L)(*[s[3] for s in ctx.shapes])
            sc = (ctypes.c_float * L)(*ctx.scales)
            ptrs = (ctypes.c_void_p * L)(*[g.data_ptr() for g in grads])
            with _c.on_device(rois.device):
                _c.check(lib.mrb_roi_align_fpn_bwd(_c._ptr(gout), ptrs, hs, ws, sc, L, _c._ptr(rois), r, n, c, pooled, sr,
                                                   k_min, k_max, ctypes.c_float(s0), lvl0, dtc, out_nhwc, _c._stream()),
                         "mrb_roi_align_fpn_bwd")
            _count(1)
        outs = [g.permute(0, 3, 1, 2).to(dt) for g in grads]  # logical NCHW, channels_last memory
        return (None,) * 9 + tuple(outs)


def roi_align_fpn(feats, rois, scales, pooled, sampling_ratio, out_nhwc=False, k_min=2, k_max=5, canonical_scale=224.0,
                  canonical_level=4):
    """Multi-level ROIAlign == the reference `Pooler` (modeling/poolers.py:45-121) in one launch."""
    if k_max - k_min + 1 != len(feats):
        raise RuntimeError("roi_align_fpn: need one feature map per level k_min..k_max")
    return _RoiAlignFpn.apply(rois, tuple(scales), pooled, sampling_ratio, out_nhwc, k_min, k_max, canonical_scale,
                              canonical_level, *feats)


# --------------------------------------------------------------------------------- batched NMS
def nms_batched(boxes, scores, sizes, threshold, presorted=False):
    """Independent NMS problems stored back to back (sizes[p] rows each).
    -> (keep int64 [sum sizes]: per problem, kept indices relative to the problem, ascending, first
    counts[p] entries valid; counts int32 [len(sizes)] on device).  No host synchronisation."""
    if not boxes.is_cuda:
        raise RuntimeError("nms_batched: expected CUDA tensors (no CPU path)")
    boxes = boxes.float().contiguous()
    scores = scores.float().contiguous()
    p = len(sizes)
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + int(s))
    if offs[-1] != boxes.shape[0] or scores.numel() != boxes.shape[0]:
        raise RuntimeError("nms_batched: sizes do not add up to the number of boxes")
    offs_c = (ctypes.c_int * (p + 1))(*offs)
    keep = torch.empty(max(offs[-1], 1), dtype=torch.int64, device=boxes.device)
    counts = torch.zeros(max(p, 1), dtype=torch.int32, device=boxes.device)
    with _c.on_device(boxes.device):
        nbytes = lib.mrb_nms_batched_workspace_bytes(offs_c, p)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=boxes.device)
        if presorted:    # rows of every problem already in descending-score order: no rank sort
            _c.check(lib.mrb_nms_batched_presorted(_c._ptr(boxes), offs_c, p, ctypes.c_float(threshold), _c._ptr(keep),
                                                   _c._ptr(counts), _c._ptr(ws), ctypes.c_size_t(nbytes), _c._stream()),
                     "mrb_nms_batched_presorted")
        else:
            _c.check(lib.mrb_nms_batched(_c._ptr(boxes), _c._ptr(scores), offs_c, p, ctypes.c_float(threshold), _c._ptr(keep),
                                         _c._ptr(counts), _c._ptr(ws), ctypes.c_size_t(nbytes), _c._stream()),
                     "mrb_nms_batched")
    _count(3)
    return keep[:offs[-1]], counts[:p]


# --------------------------------------------------------------------------------- detection glue (csrc/detect_glue.cu)
lib.mrb_rpn_anchor_match_workspace_bytes.restype = ctypes.c_size_t


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s: expected CUDA tensors (no CPU path)" % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


def rpn_decode(logits, deltas, anchors, topk_idx, image_w, image_h, boxes_out, scores_out, weights=(1.0, 1.0, 1.0, 1.0),
               xform_clip=None):
    """One level of RPNPostProcessor.forward_for_single_feature_map after its top-k (rpn/inference.py:91-111): decode
    (box_coder.py:52-95) + clip_to_image + sigmoid of the k selected anchors of every image, one launch.
    logits [N, A], deltas [N, A, 4], anchors [A, 4], topk_idx [N, k] int64 -> boxes_out [N, k, 4], scores_out [N, k]
    (caller-allocated, contiguous fp32: slices of the buffers the batched NMS reads)."""
    import math
    n, a = logits.shape
    k = topk_idx.shape[1]
    logits, deltas, anchors = _f32c(logits, "rpn_decode"), _f32c(deltas, "rpn_decode"), _f32c(anchors, "rpn_decode")
    if topk_idx.dtype != torch.int64 or not topk_idx.is_contiguous():
        topk_idx = topk_idx.long().contiguous()
    if not (boxes_out.is_contiguous() and scores_out.is_contiguous() and boxes_out.dtype == torch.float32
            and scores_out.dtype == torch.float32 and boxes_out.numel() == n * k * 4 and scores_out.numel() == n * k):
        raise RuntimeError("rpn_decode: outputs must be contiguous fp32 [N, k, 4] / [N, k]")
    w = (ctypes.c_float * 4)(*[float(x) for x in weights])
    clip = math.log(1000.0 / 16) if xform_clip is None else float(xform_clip)
    with _c.on_device(logits.device):
        _c.check(lib.mrb_rpn_decode(_c._ptr(logits), _c._ptr(deltas), _c._ptr(anchors), _c._ptr(topk_idx), _c._ptr(image_w),
                                    _c._ptr(image_h), _c._ptr(boxes_out), _c._ptr(scores_out), n, a, k, w,
                                    ctypes.c_float(clip), _c._stream()), "mrb_rpn_decode")
    _count(1)


def rpn_collect(boxes, scores, keep, counts, ks, num_images, post_nms_top_n, fpn_post_nms_top_n, per_batch, gt_boxes=None,
                gt_count=None):
    """Everything RPNPostProcessor does after the NMS (inference.py:116-123, :154-181, :53-74) in one launch.
    boxes / scores / keep / counts: the (level-major, image-minor) problems of nms_batched, level l = num_images x ks[l] rows.
    -> (boxes [N, W + gmax, 4], scores [N, W + gmax], valid [N, W + gmax] bool)."""
    boxes, scores = _f32c(boxes, "rpn_collect"), _f32c(scores, "rpn_collect")
    L = len(ks)
    kc = (ctypes.c_int * L)(*[int(k) for k in ks])
    w = lib.mrb_rpn_collect_width(kc, L, num_images, post_nms_top_n, fpn_post_nms_top_n, int(bool(per_batch)))
    gmax = 0 if gt_boxes is None else gt_boxes.shape[1]
    dev = boxes.device
    ob = torch.empty((num_images, w + gmax, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((num_images, w + gmax), dtype=torch.float32, device=dev)
    ov = torch.empty((num_images, w + gmax), dtype=torch.bool, device=dev)
    with _c.on_device(dev):
        _c.check(lib.mrb_rpn_collect(_c._ptr(boxes), _c._ptr(scores), _c._ptr(keep), _c._ptr(counts), kc, L, num_images,
                                     post_nms_top_n, fpn_post_nms_top_n, int(bool(per_batch)), int(not per_batch),
                                     _c._ptr(gt_boxes), _c._ptr(gt_count), gmax, _c._ptr(ob), _c._ptr(os_), _c._ptr(ov),
                                     _c._stream()), "mrb_rpn_collect")
    _count(1)
    return ob, os_, ov


def pad_targets(targets, device):
    """Ground truth of a batch as fixed-shape tensors: boxes [N, Gmax, 4] fp32, labels [N, Gmax] int64, count [N] int32."""
    n = len(targets)
    gmax = max(1, max(int(t["boxes"].shape[0]) for t in targets))
    gb = torch.zeros((n, gmax, 4), dtype=torch.float32, device=device)
    gl = torch.zeros((n, gmax), dtype=torch.int64, device=device)
    for i, t in enumerate(targets):
        g = int(t["boxes"].shape[0])
        if g:
            gb[i, :g] = t["boxes"]
            gl[i, :g] = t["labels"]
    cnt = torch.zeros((n,), dtype=torch.int32, device=device)
    for i, t in enumerate(targets):          # scalar fills, not a host -> device copy: the step may be under graph capture
        cnt[i:i + 1].fill_(int(t["boxes"].shape[0]))
    return gb, gl, cnt


def roi_assign_sample(boxes, valid, rand_keys, gt_boxes, gt_labels, gt_count, batch_size_per_image, positive_fraction, fg_iou,
                      bg_iou, weights, mask_rois_per_image=0, with_index=False):
    """FastRCNNLossComputation.subsample (box_head/loss.py:41-118) + the mask branch's positives-first list, one launch.
    -> dict(rois [N*S, 5], labels [N, S], reg_targets [N, S, 4], gt_index [N, S]) (+ mask_rois [N*M, 5], mask_labels [N*M],
    mask_weight [N*M], mask_gt_index [N, M] when mask_rois_per_image = M > 0)."""
    boxes, rand_keys, gt_boxes = _f32c(boxes, "roi_assign_sample"), _f32c(rand_keys, "roi_assign_sample"), _f32c(gt_boxes, "roi_assign_sample")
    n, p, _ = boxes.shape
    if valid.dtype != torch.bool or not valid.is_contiguous():
        valid = valid.bool().contiguous()
    s, m = int(batch_size_per_image), int(mask_rois_per_image)
    dev = boxes.device
    out = {"rois": torch.empty((n * s, 5), dtype=torch.float32, device=dev),
           "labels": torch.empty((n, s), dtype=torch.int64, device=dev),
           "reg_targets": torch.empty((n, s, 4), dtype=torch.float32, device=dev),
           "gt_index": torch.empty((n, s), dtype=torch.int64, device=dev)}
    if m > 0:
        out.update({"mask_rois": torch.empty((n * m, 5), dtype=torch.float32, device=dev),
                    "mask_labels": torch.empty((n * m,), dtype=torch.int64, device=dev),
                    "mask_weight": torch.empty((n * m,), dtype=torch.float32, device=dev),
                    "mask_gt_index": torch.empty((n, m), dtype=torch.int64, device=dev)})
    if with_index:
        out["index"] = torch.empty((n, s), dtype=torch.int64, device=dev)      # proposal index of every output row
    w = (ctypes.c_float * 4)(*[float(x) for x in weights])
    with _c.on_device(dev):
        _c.check(lib.mrb_roi_assign_sample(
            _c._ptr(boxes), _c._ptr(valid), _c._ptr(rand_keys), _c._ptr(gt_boxes), _c._ptr(gt_labels), _c._ptr(gt_count), n, p,
            gt_boxes.shape[1], s, ctypes.c_float(positive_fraction), ctypes.c_float(fg_iou), ctypes.c_float(bg_iou), w, m,
            _c._ptr(out["rois"]), _c._ptr(out["labels"]), _c._ptr(out["reg_targets"]), _c._ptr(out["gt_index"]),
            _c._ptr(out.get("mask_rois")), _c._ptr(out.get("mask_labels")), _c._ptr(out.get("mask_weight")),
            _c._ptr(out.get("mask_gt_index")), _c._ptr(out.get("index")), _c._stream()), "mrb_roi_assign_sample")
    _count(1)
    return out


def rpn_anchor_match(anchors, gt_boxes, gt_count, image_w, image_h, fg_iou, bg_iou, straddle_thresh):
    """RPN anchor labelling (rpn/loss.py:40-90, matcher.py:42-112 with low-quality matches) for the batch: two launches.
    -> labels [N, A] fp32 (1 / 0 / -1), matched_gt [N, A] int32."""
    anchors, gt_boxes = _f32c(anchors, "rpn_anchor_match"), _f32c(gt_boxes, "rpn_anchor_match")
    n, gmax, _ = gt_boxes.shape
    a = anchors.shape[0]
    dev = anchors.device
    labels = torch.empty((n, a), dtype=torch.float32, device=dev)
    matched = torch.empty((n, a), dtype=torch.int32, device=dev)
    with _c.on_device(dev):
        nbytes = lib.mrb_rpn_anchor_match_workspace_bytes(n, a, gmax)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _c.check(lib.mrb_rpn_anchor_match(_c._ptr(anchors), _c._ptr(gt_boxes), _c._ptr(gt_count), _c._ptr(image_w), _c._ptr(image_h),
                                          n, a, gmax, ctypes.c_float(fg_iou), ctypes.c_float(bg_iou),
                                          ctypes.c_float(straddle_thresh), _c._ptr(labels), _c._ptr(matched), _c._ptr(ws),
                                          ctypes.c_size_t(nbytes), _c._stream()), "mrb_rpn_anchor_match")
    _count(3)
    return labels, matched


def rpn_decode_packed(head_out, apl, anchors, topk_idx, image_w, image_h, boxes_out, scores_out, weights=(1.0, 1.0, 1.0, 1.0),
                      xform_clip=None):
    """rpn_decode reading logits and deltas in place from the RPN head's NHWC output `head_out` [N, H, W, ld] fp32 (apl logits,
    4 apl deltas, padding per location)."""
    import math
    if not (head_out.is_cuda and head_out.dtype == torch.float32 and head_out.is_contiguous() and head_out.dim() == 4):
        raise RuntimeError("rpn_decode_packed: head output must be a contiguous fp32 CUDA tensor [N, H, W, ld]")
    n, h, w_, ld = head_out.shape
    a = h * w_ * apl
    k = topk_idx.shape[1]
    anchors = _f32c(anchors, "rpn_decode_packed")
    if topk_idx.dtype != torch.int64 or not topk_idx.is_contiguous():
        topk_idx = topk_idx.long().contiguous()
    w = (ctypes.c_float * 4)(*[float(x) for x in weights])
    clip = math.log(1000.0 / 16) if xform_clip is None else float(xform_clip)
    with _c.on_device(head_out.device):
        _c.check(lib.mrb_rpn_decode_packed(_c._ptr(head_out), apl, ld, _c._ptr(anchors), _c._ptr(topk_idx), _c._ptr(image_w),
                                           _c._ptr(image_h), _c._ptr(boxes_out), _c._ptr(scores_out), n, a, k, w,
                                           ctypes.c_float(clip), _c._stream()), "mrb_rpn_decode_packed")
    _count(1)


def rpn_sample(labels, matched, rand_keys, anchors, gt_boxes, batch_size_per_image, positive_fraction, weights=(1.0, 1.0, 1.0, 1.0)):
    """BalancedPositiveNegativeSampler over the labelled anchors of the batch + encode of the sampled positives, one launch.
    labels [N, A] fp32 (1 / 0 / -1), matched [N, A] int32, rand_keys [N, A] iid uniform -> the six inputs of rpn_loss:
    (pos_idx [N, P], pos_ok [N, P] bool, reg_targets [N, P, 4], sel_idx [N, P + B], sel_label, sel_weight)."""
    labels, rand_keys, anchors, gt_boxes = (_f32c(t, "rpn_sample") for t in (labels, rand_keys, anchors, gt_boxes))
    n, a = labels.shape
    b = int(batch_size_per_image)
    p = int(b * positive_fraction)
    dev = labels.device
    pos_idx = torch.empty((n, p), dtype=torch.int64, device=dev)
    pos_ok = torch.empty((n, p), dtype=torch.bool, device=dev)
    reg_t = torch.empty((n, p, 4), dtype=torch.float32, device=dev)
    sel_idx = torch.empty((n, p + b), dtype=torch.int64, device=dev)
    sel_lab = torch.empty((n, p + b), dtype=torch.float32, device=dev)
    sel_w = torch.empty((n, p + b), dtype=torch.float32, device=dev)
    w = (ctypes.c_float * 4)(*[float(x) for x in weights])
    with _c.on_device(dev):
        _c.check(lib.mrb_rpn_sample(_c._ptr(labels), _c._ptr(matched), _c._ptr(rand_keys), _c._ptr(anchors), _c._ptr(gt_boxes), n, a,
                                    gt_boxes.shape[1], b, ctypes.c_float(positive_fraction), w, _c._ptr(pos_idx), _c._ptr(pos_ok),
                                    _c._ptr(reg_t), _c._ptr(sel_idx), _c._ptr(sel_lab), _c._ptr(sel_w), _c._stream()),
                 "mrb_rpn_sample")
    _count(1)
    return pos_idx, pos_ok, reg_t, sel_idx, sel_lab, sel_w


def rpn_topk_decode(head_out, apl, anchors, k, image_w, image_h, boxes_out, scores_out, weights=(1.0, 1.0, 1.0, 1.0), xform_clip=None):
    """objectness.topk(k, sorted=True) + rpn_decode_packed in ONE launch (a thread-block cluster per image): the k best anchors of
    the level by logit, descending (ascending anchor index among equal logits), decoded and clipped, with their sigmoids."""
    import math
    if not (head_out.is_cuda and head_out.dtype == torch.float32 and head_out.is_contiguous() and head_out.dim() == 4):
        raise RuntimeError("rpn_topk_decode: head output must be a contiguous fp32 CUDA tensor [N, H, W, ld]")
    n, h, w_, ld = head_out.shape
    a = h * w_ * apl
    anchors = _f32c(anchors, "rpn_topk_decode")
    w = (ctypes.c_float * 4)(*[float(x) for x in weights])
    clip = math.log(1000.0 / 16) if xform_clip is None else float(xform_clip)
    with _c.on_device(head_out.device):
        _c.check(lib.mrb_rpn_topk_decode(_c._ptr(head_out), apl, ld, _c._ptr(anchors), _c._ptr(image_w), _c._ptr(image_h),
                                         _c._ptr(boxes_out), _c._ptr(scores_out), n, a, int(k), w, ctypes.c_float(clip), _c._stream()),
                 "mrb_rpn_topk_decode")
    _count(1)


def box_postprocess(outputs, num_classes, proposals, valid, image_w, image_h, score_thresh, weights, nms_thresh, detections_per_img,
                    xform_clip=None):
    """PostProcessor.forward (box_head/inference.py:45-149) without host synchronisation: softmax + per-class decode + clip +
    score threshold (one launch), one batched NMS over the N x (C-1) (image, class) problems, top detections_per_img per image
    (one cluster launch).  outputs [N*P, ld >= 5C] fp32 (C logits, 4C regression outputs), proposals [N, P, 4], valid [N, P].
    -> (boxes [N, D, 4], scores [N, D], labels [N, D] int64, count [N] int32); rows beyond count are zero."""
    import math
    if not (outputs.is_cuda and outputs.dtype == torch.float32 and outputs.is_contiguous()):
        raise RuntimeError("box_postprocess: outputs must be a contiguous fp32 CUDA tensor (no CPU path)")
    n, p, _ = proposals.shape
    proposals = _f32c(proposals, "box_postprocess")
    if valid.dtype != torch.bool or not valid.is_contiguous():
        valid = valid.bool().contiguous()
    ld = outputs.shape[1]
    dev = outputs.device
    nprob = n * (num_classes - 1)
    boxes = torch.empty((nprob * p, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((nprob * p,), dtype=torch.float32, device=dev)
    w = (ctypes.c_float * 4)(*[float(x) for x in weights])
    clip = math.log(1000.0 / 16) if xform_clip is None else float(xform_clip)
    with _c.on_device(dev):
        _c.check(lib.mrb_box_post_decode(_c._ptr(outputs), ld, num_classes, _c._ptr(proposals), _c._ptr(valid), _c._ptr(image_w),
                                         _c._ptr(image_h), n, p, ctypes.c_float(score_thresh), w, ctypes.c_float(clip),
                                         _c._ptr(boxes), _c._ptr(scores), _c._stream()), "mrb_box_post_decode")
    _count(1)
    keep, counts = nms_batched(boxes, scores, [p] * nprob, nms_thresh)
    d = int(detections_per_img)
    ob = torch.empty((n, d, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((n, d), dtype=torch.float32, device=dev)
    ol = torch.empty((n, d), dtype=torch.int64, device=dev)
    on = torch.empty((n,), dtype=torch.int32, device=dev)
    with _c.on_device(dev):
        _c.check(lib.mrb_box_post_select(_c._ptr(boxes), _c._ptr(scores), _c._ptr(keep), _c._ptr(counts), n, p, num_classes, d,
                                         _c._ptr(ob), _c._ptr(os_), _c._ptr(ol), _c._ptr(on), _c._stream()), "mrb_box_post_select")
    _count(1)
    return ob, os_, ol, on


# --------------------------------------------------------------------------------- fused losses (csrc/loss_glue.cu)
def _gscalar(g, like):
    if g is None:
        return torch.zeros((), dtype=torch.float32, device=like.device)
    return g.detach().to(torch.float32).contiguous()


class _RpnLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sel_idx, sel_label, sel_weight, pos_idx, pos_ok, reg_t, apl, beta, *outs):
        outs = [o if o.is_contiguous() else o.contiguous() for o in outs]
        n, ld = outs[0].shape[0], outs[0].shape[3]
        L = len(outs)
        hw = (ctypes.c_int * L)(*[o.shape[1] * o.shape[2] for o in outs])
        ptrs = (ctypes.c_void_p * L)(*[o.data_ptr() for o in outs])
        result = torch.empty(3, dtype=torch.float32, device=outs[0].device)
        with _c.on_device(outs[0].device):
            _c.check(lib.mrb_rpn_loss_fwd(ptrs, hw, L, n, apl, ld, _c._ptr(sel_idx), _c._ptr(sel_label), _c._ptr(sel_weight),
                                          sel_idx.shape[1], _c._ptr(pos_idx), _c._ptr(pos_ok), _c._ptr(reg_t), pos_idx.shape[1],
                                          ctypes.c_float(beta), _c._ptr(result), _c._stream()), "mrb_rpn_loss_fwd")
        _count(1)
        ctx.save_for_backward(sel_idx, sel_label, sel_weight, pos_idx, pos_ok, reg_t, result, *outs)
        ctx.geom = (apl, beta)
        return result[0], result[1]

    @staticmethod
    def backward(ctx, g_obj, g_box):
        sel_idx, sel_label, sel_weight, pos_idx, pos_ok, reg_t, result = ctx.saved_tensors[:7]
        outs = ctx.saved_tensors[7:]
        apl, beta = ctx.geom
        n, ld = outs[0].shape[0], outs[0].shape[3]
        L = len(outs)
        flat = torch.zeros(sum(o.numel() for o in outs), dtype=torch.float32, device=outs[0].device)
        grads, off = [], 0
        for o in outs:
            grads.append(flat[off:off + o.numel()].view(o.shape))
            off += o.numel()
        hw = (ctypes.c_int * L)(*[o.shape[1] * o.shape[2] for o in outs])
        ptrs = (ctypes.c_void_p * L)(*[o.data_ptr() for o in outs])
        gptrs = (ctypes.c_void_p * L)(*[g.data_ptr() for g in grads])
        go, gb = _gscalar(g_obj, flat), _gscalar(g_box, flat)
        with _c.on_device(flat.device):
            _c.check(lib.mrb_rpn_loss_bwd(ptrs, gptrs, hw, L, n, apl, ld, _c._ptr(sel_idx), _c._ptr(sel_label), _c._ptr(sel_weight),
                                          sel_idx.shape[1], _c._ptr(pos_idx), _c._ptr(pos_ok), _c._ptr(reg_t), pos_idx.shape[1],
                                          ctypes.c_float(beta), _c._ptr(result), _c._ptr(go), _c._ptr(gb), _c._stream()),
                     "mrb_rpn_loss_bwd")
        _count(1)
        return (None,) * 8 + tuple(grads)


def rpn_loss(outs, apl, sel_idx, sel_label, sel_weight, pos_idx, pos_ok, reg_targets, beta=1.0 / 9):
    """RPNLossComputation.__call__ (rpn/loss.py:92-131) on the head's NHWC outputs `outs` (one [N, H, W, ld] fp32 tensor per
    level: apl logits, 4 apl deltas, padding): -> (loss_objectness, loss_rpn_box_reg).  sel_* [N, S], pos_* [N, P]."""
    sel_idx = sel_idx.long().contiguous()
    pos_idx = pos_idx.long().contiguous()
    pos_ok = pos_ok.bool().contiguous()
    return _RpnLossFn.apply(sel_idx, sel_label.float().contiguous(), sel_weight.float().contiguous(), pos_idx, pos_ok,
                            reg_targets.float().contiguous(), int(apl), float(beta), *outs)


class _BoxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, labels, reg_t, num_classes):
        if not o.is_contiguous():
            o = o.contiguous()
        r, ld = o.shape
        result = torch.empty(3, dtype=torch.float32, device=o.device)
        rows = torch.empty(3 * r, dtype=torch.float32, device=o.device)
        with _c.on_device(o.device):
            _c.check(lib.mrb_box_loss_fwd(_c._ptr(o), ld, num_classes, _c._ptr(labels), _c._ptr(reg_t), r, _c._ptr(rows),
                                          _c._ptr(result), _c._stream()), "mrb_box_loss_fwd")
        _count(2)
        ctx.save_for_backward(o, labels, reg_t, result)
        ctx.nc = num_classes
        return result[0], result[1]

    @staticmethod
    def backward(ctx, g_cls, g_box):
        o, labels, reg_t, result = ctx.saved_tensors
        r, ld = o.shape
        d_o = torch.empty_like(o)
        gc, gb = _gscalar(g_cls, o), _gscalar(g_box, o)
        with _c.on_device(o.device):
            _c.check(lib.mrb_box_loss_bwd(_c._ptr(o), ld, ctx.nc, _c._ptr(labels), _c._ptr(reg_t), r, _c._ptr(result), _c._ptr(gc),
                                          _c._ptr(gb), _c._ptr(d_o), _c._stream()), "mrb_box_loss_bwd")
        _count(1)
        return d_o, None, None, None


def box_head_loss(outputs, labels, reg_targets, num_classes):
    """FastRCNNLossComputation.__call__ (box_head/loss.py:120-167) on the predictor output [R, ld >= 5 C] fp32 (C class logits,
    then 4 C regression outputs): -> (loss_classifier, loss_box_reg).  labels [R] int64 (-1 = not sampled)."""
    if outputs.dtype != torch.float32 or not outputs.is_cuda:
        raise RuntimeError("box_head_loss: expected a fp32 CUDA tensor (no CPU path)")
    return _BoxLossFn.apply(outputs, labels.reshape(-1).long().contiguous(), reg_targets.reshape(-1, 4).float().contiguous(),
                            int(num_classes))


class _MaskLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, labels, targets, weights):
        r, c, h, w = y.shape
        if y.dtype != torch.bfloat16 or not y.is_contiguous(memory_format=torch.channels_last):
            y = y.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        rows = torch.empty(r, dtype=torch.float32, device=y.device)
        result = torch.empty(2, dtype=torch.float32, device=y.device)
        with _c.on_device(y.device):
            _c.check(lib.mrb_mask_loss_fwd(_c._ptr(y), c, h * w, _c._ptr(labels), _c._ptr(targets), _c._ptr(weights), r,
                                           _c._ptr(rows), _c._ptr(result), _c._stream()), "mrb_mask_loss_fwd")
        _count(2)
        ctx.save_for_backward(y, labels, targets, weights, result)
        return result[0]

    @staticmethod
    def backward(ctx, g):
        y, labels, targets, weights, result = ctx.saved_tensors
        r, c, h, w = y.shape
        gy = torch.empty_like(y, memory_format=torch.channels_last)
        gg = _gscalar(g, y)
        with _c.on_device(y.device):
            _c.check(lib.mrb_mask_loss_bwd(_c._ptr(y), c, h * w, _c._ptr(labels), _c._ptr(targets), _c._ptr(weights), r,
                                           _c._ptr(result), _c._ptr(gg), _c._ptr(gy), _c._stream()), "mrb_mask_loss_bwd")
        _count(1)
        return gy, None, None, None


def mask_head_loss(logits_nhwc, labels, targets, weights):
    """MaskRCNNLossComputation.__call__ (mask_head/loss.py:100-133), fixed-shape form: logits [R, C, M, M] bf16 channels_last
    (C % 8 == 0), labels [R] (class plane read per ROI), targets [R, M, M] fp32, weights [R] -> weighted mean of the per-ROI
    mean BCE-with-logits."""
    if logits_nhwc.shape[1] % 8:
        raise RuntimeError("mask_head_loss: channel count must be a multiple of 8")
    return _MaskLossFn.apply(logits_nhwc, labels.long().contiguous(), targets.float().contiguous(), weights.float().contiguous())
