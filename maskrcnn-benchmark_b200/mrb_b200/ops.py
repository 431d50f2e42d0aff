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


def _nhwc(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s: expected a CUDA tensor (no CPU path)" % name)
    if t.dim() != 4:
        raise RuntimeError("%s: expected a 4-D tensor" % name)
    if not t.is_contiguous(memory_format=torch.channels_last):
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def _params(x_shape, w_shape, stride, pad, relu, out_dtype):
    n, c, h, w = x_shape
    co, ci, kh, kw = w_shape
    if ci != c:
        raise RuntimeError("conv2d: weight expects %d input channels, got %d" % (ci, c))
    p = _c.ConvParams()
    p.batch, p.height, p.width, p.cin = n, h, w, c
    p.cout, p.kh, p.kw = co, kh, kw
    p.stride, p.pad, p.relu, p.out_dtype = stride, pad, int(bool(relu)), _DT[out_dtype]
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    return p, ho, wo


def conv2d_fwd(x, weight, scale=None, bias=None, residual=None, stride=1, pad=0, relu=False,
               out_dtype=torch.bfloat16):
    """y = act(conv(x, weight) * scale[c] + bias[c] + residual).  x, weight bf16 channels_last."""
    x = _nhwc(x, "conv2d_fwd(x)")
    weight = _nhwc(weight, "conv2d_fwd(weight)")
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise RuntimeError("conv2d_fwd: bf16 operands required")
    p, ho, wo = _params(x.shape, weight.shape, stride, pad, relu, out_dtype)
    out = torch.empty((p.batch, p.cout, ho, wo), dtype=out_dtype, device=x.device,
                      memory_format=torch.channels_last)
    if residual is not None:
        residual = _nhwc(residual, "conv2d_fwd(residual)")
        if residual.dtype != torch.bfloat16 or residual.shape != out.shape:
            raise RuntimeError("conv2d_fwd: residual must be bf16 and shaped like the output")
    for v in (scale, bias):
        if v is not None and (v.dtype != torch.float32 or v.numel() != p.cout or not v.is_contiguous()):
            raise RuntimeError("conv2d_fwd: scale/bias must be contiguous fp32 [Cout]")
    with torch.cuda.device(x.device):
        _c.check(lib.mrb_conv2d_fwd(ctypes.byref(p), _c._ptr(x), _c._ptr(weight), _c._ptr(scale), _c._ptr(bias),
                                    _c._ptr(residual), _c._ptr(out), _c._stream()), "mrb_conv2d_fwd")
    return out


def conv2d_dgrad(grad_out, weight, x_shape, scale=None, add=None, relu_mask=None, stride=1, pad=0,
                 out_dtype=torch.bfloat16):
    """grad_x = conv_transpose(grad_out, weight * scale[cout]) (+ add) masked by (relu_mask > 0)."""
    grad_out = _nhwc(grad_out, "conv2d_dgrad(grad_out)")
    weight = _nhwc(weight, "conv2d_dgrad(weight)")
    if grad_out.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise RuntimeError("conv2d_dgrad: bf16 operands required")
    p, ho, wo = _params(tuple(x_shape), weight.shape, stride, pad, False, out_dtype)
    if tuple(grad_out.shape) != (p.batch, p.cout, ho, wo):
        raise RuntimeError("conv2d_dgrad: grad_out shape %s != %s" % (tuple(grad_out.shape), (p.batch, p.cout, ho, wo)))
    gx = torch.empty(tuple(x_shape), dtype=out_dtype, device=grad_out.device, memory_format=torch.channels_last)
    for t in (add, relu_mask):
        if t is not None and (t.dtype != torch.bfloat16 or tuple(t.shape) != tuple(x_shape)):
            raise RuntimeError("conv2d_dgrad: add/relu_mask must be bf16 and shaped like x")
    add = _nhwc(add, "add") if add is not None else None
    relu_mask = _nhwc(relu_mask, "relu_mask") if relu_mask is not None else None
    with torch.cuda.device(grad_out.device):
        nbytes = lib.mrb_conv2d_dgrad_workspace_bytes(ctypes.byref(p))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=grad_out.device)
        _c.check(lib.mrb_conv2d_dgrad(ctypes.byref(p), _c._ptr(grad_out), _c._ptr(weight), _c._ptr(scale), _c._ptr(add),
                                      _c._ptr(relu_mask), _c._ptr(gx), _c._ptr(ws), ctypes.c_size_t(nbytes),
                                      _c._stream()), "mrb_conv2d_dgrad")
    return gx
