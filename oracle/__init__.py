"""Oracle loader -- TEST INFRASTRUCTURE ONLY.

`oracle` is the CPU restatement (mrb_oracle*.c, built by oracle/Makefile) of the
reference's csrc hot path plus, when available, `oracle/_ref` (the reference's
own csrc/cpu sources AND its csrc/cuda kernels compiled in place by
oracle/build_ref.py; the CUDA ones are the GPU-side checker).  Only tests/,
bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() import
this package, and only as the checker.  Nothing under maskrcnn-benchmark_b200/
may import it.
"""
import ctypes
import importlib.util
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build():
    """Compile libmrb_oracle.so (gcc, seconds) and, if the reference tree is here, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libmrb_oracle.so"])
    if os.path.isdir(os.environ.get("MRB_REFERENCE", "/root/reference")):
        from . import build_ref
        build_ref.build()
        build_ref.build_cuda()


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmrb_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_nms.restype = ctypes.c_int64
    return _LIB


def ref():
    """The reference's own CPU `_C` subset (nms, roi_align_forward); None if never built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "mrb_ref_C.so")
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location("mrb_ref_C", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _REF = mod
    return _REF


_REF_CUDA = None


def ref_cuda():
    """The reference's own CUDA kernels (csrc/cuda/*.cu compiled for sm_100a by build_ref.build_cuda): module with the 14
    `_C` names in their *_cuda forms, or None if never built / not loadable.  GPU-side checker only."""
    global _REF_CUDA
    if _REF_CUDA is None:
        path = os.path.join(_HERE, "_ref", "cuda", "mrb_ref_cuda.so")
        if not os.path.exists(path):
            return None
        try:
            spec = importlib.util.spec_from_file_location("mrb_ref_cuda", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        except Exception:
            return None
        _REF_CUDA = mod
    return _REF_CUDA


def _f32(t):
    t = torch.as_tensor(t)
    assert t.dtype == torch.float32 and t.device.type == "cpu"
    return t.contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


_F = ctypes.c_float
_I = ctypes.c_int
_L = ctypes.c_int64


def nms(dets, scores, threshold, order=None):
    dets, scores = _f32(dets), _f32(scores)
    n = dets.shape[0]
    keep = torch.empty(n, dtype=torch.int64)
    if order is not None:
        order = order.to(torch.int64).contiguous()
    k = lib().orc_nms(_p(dets), _p(scores), _L(n), _F(threshold),
                      _p(order) if order is not None else None, _p(keep))
    return keep[:k].clone()


def roi_align_forward(inp, rois, spatial_scale, ph, pw, sampling_ratio):
    inp, rois = _f32(inp), _f32(rois)
    n, c, h, w = inp.shape
    r = rois.shape[0]
    out = torch.empty(r, c, ph, pw, dtype=torch.float32)
    lib().orc_roi_align_fwd(_p(inp), _p(rois), _I(r), _I(c), _I(h), _I(w), _I(ph), _I(pw),
                            _F(spatial_scale), _I(sampling_ratio), _p(out))
    return out


def roi_align_backward(grad, rois, spatial_scale, ph, pw, bs, ch, h, w, sampling_ratio):
    grad, rois = _f32(grad), _f32(rois)
    r = rois.shape[0]
    gin = torch.empty(bs, ch, h, w, dtype=torch.float32)
    lib().orc_roi_align_bwd(_p(grad), _p(rois), _I(r), _I(bs), _I(ch), _I(h), _I(w), _I(ph), _I(pw),
                            _F(spatial_scale), _I(sampling_ratio), _p(gin))
    return gin


def roi_pool_forward(inp, rois, spatial_scale, ph, pw):
    inp, rois = _f32(inp), _f32(rois)
    n, c, h, w = inp.shape
    r = rois.shape[0]
    out = torch.empty(r, c, ph, pw, dtype=torch.float32)
    argmax = torch.empty(r, c, ph, pw, dtype=torch.int32)
    lib().orc_roi_pool_fwd(_p(inp), _p(rois), _I(r), _I(c), _I(h), _I(w), _I(ph), _I(pw),
                           _F(spatial_scale), _p(out), _p(argmax))
    return out, argmax


def roi_pool_backward(grad, rois, argmax, bs, ch, h, w):
    grad, rois = _f32(grad), _f32(rois)
    argmax = argmax.to(torch.int32).contiguous()
    r, _, ph, pw = grad.shape
    gin = torch.empty(bs, ch, h, w, dtype=torch.float32)
    lib().orc_roi_pool_bwd(_p(grad), _p(rois), _p(argmax), _I(r), _I(bs), _I(ch), _I(h), _I(w),
                           _I(ph), _I(pw), _p(gin))
    return gin


def sigmoid_focalloss_forward(logits, targets, num_classes, gamma, alpha):
    logits = _f32(logits)
    targets = targets.to(torch.int32).contiguous()
    out = torch.empty_like(logits)
    lib().orc_sigmoid_focal_fwd(_p(logits), _p(targets), _L(logits.shape[0]), _I(num_classes),
                                _F(gamma), _F(alpha), _p(out))
    return out


def sigmoid_focalloss_backward(logits, targets, d_losses, num_classes, gamma, alpha):
    logits, d_losses = _f32(logits), _f32(d_losses)
    targets = targets.to(torch.int32).contiguous()
    out = torch.empty_like(logits)
    lib().orc_sigmoid_focal_bwd(_p(logits), _p(targets), _p(d_losses), _L(logits.shape[0]),
                                _I(num_classes), _F(gamma), _F(alpha), _p(out))
    return out


# ------------------------------------------------------------------ deformable conv / psroi
class _Dcn(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "batch", "cin", "H", "W", "cout", "kh", "kw", "sh", "sw", "ph", "pw", "dh", "dw", "groups", "dg")]


class _Ps(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "channels", "height", "width", "pooled", "output_dim", "group_size", "part_size", "sample_per_part",
        "num_classes", "channels_each_class", "no_trans")] + [("spatial_scale", ctypes.c_float),
                                                              ("trans_std", ctypes.c_float)]


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _dcn(inp, weight, stride, padding, dilation, groups, dg):
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    n, c, h, w = inp.shape
    p = _Dcn(n, c, h, w, weight.shape[0], weight.shape[2], weight.shape[3], sh, sw, ph, pw, dh, dw, groups, dg)
    ho = (h + 2 * ph - (dh * (p.kh - 1) + 1)) // sh + 1
    wo = (w + 2 * pw - (dw * (p.kw - 1) + 1)) // sw + 1
    return p, ho, wo


def deform_conv_forward(inp, offset, mask, weight, bias, stride=1, padding=0, dilation=1, groups=1, dg=1):
    """mask=None -> DCNv1 (deform_conv_forward), else DCNv2 (modulated_deform_conv_forward)."""
    inp, offset, weight = _f32(inp), _f32(offset), _f32(weight)
    mask = _f32(mask) if mask is not None else None
    bias = _f32(bias) if bias is not None else None
    p, ho, wo = _dcn(inp, weight, stride, padding, dilation, groups, dg)
    out = torch.empty(p.batch, p.cout, ho, wo, dtype=torch.float32)
    lib().orc_deform_conv_fwd(ctypes.byref(p), _p(inp), _p(offset), _p(mask) if mask is not None else None,
                              _p(weight), _p(bias) if bias is not None else None, _p(out))
    return out


def deform_conv_backward(inp, offset, mask, weight, grad_out, stride=1, padding=0, dilation=1, groups=1, dg=1,
                         with_bias=False, scale=1.0):
    """-> dict(grad_input, grad_offset, grad_mask, grad_weight, grad_bias)"""
    inp, offset, weight, grad_out = _f32(inp), _f32(offset), _f32(weight), _f32(grad_out)
    mask = _f32(mask) if mask is not None else None
    p, ho, wo = _dcn(inp, weight, stride, padding, dilation, groups, dg)
    gi, go, gw = torch.zeros_like(inp), torch.zeros_like(offset), torch.zeros_like(weight)
    gm = torch.zeros_like(mask) if mask is not None else None
    gb = torch.zeros(p.cout) if with_bias else None
    lib().orc_deform_conv_bwd(ctypes.byref(p), _p(inp), _p(offset), _p(mask) if mask is not None else None,
                              _p(weight), _p(grad_out), _p(gi), _p(go), _p(gm) if gm is not None else None,
                              _p(gw), _p(gb) if gb is not None else None, _F(scale))
    return dict(grad_input=gi, grad_offset=go, grad_mask=gm, grad_weight=gw, grad_bias=gb)


def _ps(data, trans, no_trans, spatial_scale, output_dim, group_size, pooled, part_size, spp, trans_std):
    n, c, h, w = data.shape
    ncls = 1 if no_trans else trans.shape[1] // 2
    cec = output_dim if no_trans else output_dim // ncls
    return _Ps(c, h, w, pooled, output_dim, group_size, part_size, spp, ncls, cec, int(bool(no_trans)),
               spatial_scale, trans_std)


def deform_psroi_forward(data, rois, trans, no_trans, spatial_scale, output_dim, group_size, pooled, part_size,
                         spp, trans_std):
    data, rois = _f32(data), _f32(rois)
    trans = _f32(trans) if not no_trans else None
    a = _ps(data, trans, no_trans, spatial_scale, output_dim, group_size, pooled, part_size, spp, trans_std)
    r = rois.shape[0]
    out = torch.empty(r, output_dim, pooled, pooled)
    cnt = torch.empty(r, output_dim, pooled, pooled)
    lib().orc_deform_psroi_fwd(ctypes.byref(a), _I(r), _p(data), _p(rois), _p(trans) if trans is not None else None,
                               _p(out), _p(cnt))
    return out, cnt


def deform_psroi_backward(out_grad, data, rois, trans, top_count, no_trans, spatial_scale, output_dim, group_size,
                          pooled, part_size, spp, trans_std):
    out_grad, data, rois, top_count = _f32(out_grad), _f32(data), _f32(rois), _f32(top_count)
    trans = _f32(trans) if not no_trans else None
    a = _ps(data, trans, no_trans, spatial_scale, output_dim, group_size, pooled, part_size, spp, trans_std)
    gi = torch.zeros_like(data)
    gt = torch.zeros_like(trans) if trans is not None else None
    lib().orc_deform_psroi_bwd(ctypes.byref(a), _I(rois.shape[0]), _p(out_grad), _p(data), _p(rois),
                               _p(trans) if trans is not None else None, _p(top_count), _p(gi),
                               _p(gt) if gt is not None else None)
    return gi, gt
