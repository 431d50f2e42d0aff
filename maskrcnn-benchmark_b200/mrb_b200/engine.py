"""Dispatch of `maskrcnn_benchmark.layers.{Conv2d, ConvTranspose2d}` (reference layers/misc.py:30-60) onto the
tcgen05 conv engine of libmrb_b200.so.

This is the drop-in boundary for the dense convolutions: the reference's UNMODIFIED module graph
(modeling/backbone/resnet.py, fpn.py, make_layers.py, roi_heads/*) instantiates `layers.Conv2d`, and every
forward of a CUDA tensor lands here.  Semantics are torch.nn.Conv2d's: same logical NCHW shape and the input's
dtype come back; the result is in torch.channels_last memory (NHWC, what the engine reads and writes) so that
a chain of engine convs never transposes.  Operands are rounded to bf16, accumulation is fp32 (the
reference's mixed-precision mode, `DTYPE: float16` + apex O1, does the same with fp16).

Unsupported geometries (groups > 1 before the grouped path is enabled, dilation, Cin % 8 != 0 other than the
7x7 stem, fp16/fp64 inputs) return None: the caller falls back to ATen and the event is logged once per
signature on the `mrb_b200` logger.  MRB_CONV_ENGINE=0 disables the dispatch altogether (A/B switch).
The fused path (conv + FrozenBN + residual + ReLU in one epilogue) is mrb_b200.fuse.fuse_model()."""
import logging
import os

import torch

log = logging.getLogger("mrb_b200")
_BACKEND = None
_LOGGED = set()
STATS = {"engine": 0, "aten": 0}


def default_backend():
    """Process-wide B200Backend (bf16 operand caches keyed per nn.Parameter, flipped dgrad weights)."""
    global _BACKEND
    if _BACKEND is None:
        from mrb_b200.model.backend import B200Backend
        _BACKEND = B200Backend()
    return _BACKEND


def set_default_backend(be):
    global _BACKEND
    _BACKEND = be


def enabled():
    return os.environ.get("MRB_CONV_ENGINE", "1") != "0"


def _fallback(why, mod, x):
    key = (why, type(mod).__name__, tuple(mod.weight.shape), tuple(getattr(mod, "stride", ())), str(x.dtype))
    if key not in _LOGGED:
        _LOGGED.add(key)
        log.warning("mrb_b200: ATen fallback for %s weight=%s stride=%s dtype=%s: %s", type(mod).__name__,
                    tuple(mod.weight.shape), tuple(getattr(mod, "stride", ())), x.dtype, why)
    STATS["aten"] += 1
    return None


def _pair_eq(v, a):
    return tuple(v) == (a, a)


def conv2d_module(mod, x):
    """Forward of a torch.nn.Conv2d-shaped module `mod` on CUDA tensor x through the engine, or None."""
    if not enabled() or not x.is_cuda:
        return None
    if x.dtype not in (torch.float32, torch.bfloat16):
        return _fallback("input dtype", mod, x)
    if mod.padding_mode != "zeros" or isinstance(mod.padding, str):
        return _fallback("padding mode", mod, x)
    if tuple(mod.dilation) != (1, 1):
        return _fallback("dilation", mod, x)
    co, ci, kh, kw = mod.weight.shape
    sh, sw = mod.stride
    ph, pw = mod.padding
    be = default_backend()
    out_fp32 = x.dtype == torch.float32
    if mod.groups != 1:
        if hasattr(be, "grouped_conv") and sh == sw:
            y = be.grouped_conv(x, mod.weight, mod.bias, mod.groups, sh, (ph, pw), out_fp32=out_fp32)
            if y is not None:
                STATS["engine"] += 1
                return y
        return _fallback("groups", mod, x)
    if sh != sw or sh not in (1, 2):
        return _fallback("stride", mod, x)
    if ci == 3 and (kh, kw, sh, ph, pw) == (7, 7, 2, 3, 3) and co % 8 == 0:
        # the 7x7/2 stem (resnet.py:353): 4x4 conv on the 2x2 space-to-depth image, see B200Backend.stem
        if torch.is_grad_enabled() and (mod.weight.requires_grad or x.requires_grad):
            return _fallback("trainable stem (the engine's stem path has no backward; FREEZE_CONV_BODY_AT >= 1 freezes it)", mod, x)
        y = be.stem(x, mod.weight, None, None, relu=False, bias=mod.bias, out_fp32=out_fp32)
        STATS["engine"] += 1
        return y
    if ci % 8:
        return _fallback("Cin % 8", mod, x)
    if sh == 2 and not ((kh, kw, ph, pw) == (1, 1, 0, 0) or getattr(be, "stride2_3x3", False)):
        return _fallback("stride-2 kxk", mod, x)
    x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = be.conv(x16, mod.weight, bias=mod.bias, stride=sh, pad=(ph, pw) if ph != pw else ph, out_fp32=out_fp32)
    STATS["engine"] += 1
    return y


def conv_transpose2d_module(mod, x):
    """ConvTranspose2d(k=2, s=2, p=0) (the mask head's conv5_mask, roi_mask_predictors.py:17-19) or None."""
    if not enabled() or not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
        return None
    ci, co, kh, kw = mod.weight.shape
    if (kh, kw) != (2, 2) or not _pair_eq(mod.stride, 2) or not _pair_eq(mod.padding, 0) or \
            not _pair_eq(mod.output_padding, 0) or mod.groups != 1 or not _pair_eq(mod.dilation, 1) or ci % 8 or co % 8:
        return _fallback("deconv geometry", mod, x)
    be = default_backend()
    y = be.deconv2x2(x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last), mod.weight, mod.bias)
    STATS["engine"] += 1
    return y.to(x.dtype)
