"""maskrcnn_benchmark._C -- the reference's native extension surface, served by libmrb_b200.so.

The reference builds `_C` from csrc/vision.cpp:9-25 (14 pybind functions).  This module exposes
the same 14 names with the same positional signatures and return conventions, as a thin ctypes
marshalling layer over the C ABI declared in include/mrb_b200.h: tensors -> device pointers,
sizes, the current CUDA stream.  All arithmetic happens in hand-written sm_100a kernels.

There is NO CPU implementation and no fallback: a CPU tensor, a missing libmrb_b200.so or a
kernel error raises RuntimeError (the reference raised "Not implemented on the CPU" for 12 of
the 14 functions, csrc/ROIAlign.h:44 etc.; here that holds for all 14).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmrb_b200.so")

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_int64 = ctypes.c_int64
_c_size_t = ctypes.c_size_t


class DcnParams(ctypes.Structure):
    """struct mrb_dcn_params (include/mrb_b200.h)"""
    _fields_ = [(n, _c_int) for n in (
        "batch", "cin", "height", "width", "cout", "kh", "kw", "stride_h", "stride_w", "pad_h", "pad_w",
        "dil_h", "dil_w", "groups", "deformable_groups")]


class ConvParams(ctypes.Structure):
    """struct mrb_conv_params (include/mrb_b200.h)"""
    _fields_ = [(n, _c_int) for n in (
        "batch", "height", "width", "cin", "cout", "kh", "kw", "stride", "pad", "relu", "out_dtype", "out_h", "out_w",
        "pad_w", "flags")] + [("x_pitch", ctypes.c_longlong * 3), ("y_pitch", ctypes.c_longlong * 3)]


def _load():
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            "maskrcnn_benchmark._C: %s is missing. Build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    lib.mrb_error_string.restype = ctypes.c_char_p
    lib.mrb_nms_workspace_bytes.restype = _c_size_t
    lib.mrb_nms_batched_workspace_bytes.restype = _c_size_t
    lib.mrb_deform_conv_workspace_bytes.restype = _c_size_t
    lib.mrb_deform_conv_workspace_bytes.argtypes = [ctypes.POINTER(DcnParams)]
    if lib.mrb_version() != 100:
        raise RuntimeError("maskrcnn_benchmark._C: libmrb_b200.so version mismatch; rebuild")
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s (code %d)" % (what, lib.mrb_error_string(int(rc)).decode(), rc))


def _ptr(t):
    return _c_void_p(t.data_ptr()) if t is not None else _c_void_p(0)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw-handle query is ~10x cheaper than building a
    torch.cuda.Stream object: the eager reference graph makes ~270 library calls per step)."""
    if _raw_stream is not None:
        return _c_void_p(_raw_stream(torch.cuda.current_device()))
    return _c_void_p(torch.cuda.current_stream().cuda_stream)


class on_device:
    """`with torch.cuda.device(t.device)` that costs nothing when t already lives on the current device."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.ctx = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            return self.ctx.__exit__(*a)
        return False


def _require_cuda(name, *tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "maskrcnn_benchmark._C.%s: expected CUDA tensors; this build has no CPU path "
                "(Blackwell sm_100a kernels only)" % name)


def _require_f32(name, *tensors):
    for t in tensors:
        if t.dtype != torch.float32:
            raise RuntimeError("maskrcnn_benchmark._C.%s: expected float32 tensors, got %s" % (name, t.dtype))


def _is_channels_last(t):
    return t.dim() == 4 and t.shape[1] > 1 and (t.shape[2] > 1 or t.shape[3] > 1) and \
        t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def _layout_of(t):
    """(tensor, MRB_LAYOUT_*): channels_last tensors are consumed in place as NHWC."""
    if _is_channels_last(t):
        return t, 1
    return t.contiguous(), 0


# ----------------------------------------------------------------------------------------- nms
def nms(dets, scores, threshold):
    """csrc/nms.h:10-28.  dets [N,4] xyxy fp32, scores [N] -> int64 kept indices, ascending.

    Suppression rule and arithmetic follow the reference CPU kernel (IoU >= threshold,
    "+1" areas; csrc/cpu/nms_cpu.cpp:22,49-60)."""
    _require_cuda("nms", dets, scores)
    if dets.numel() == 0:
        # csrc/nms.h:17-18: an empty CUDA input returns an empty int64 CPU tensor
        return torch.empty((0,), dtype=torch.int64, device="cpu")
    _require_f32("nms", dets, scores)
    dets = dets.contiguous()
    scores = scores.contiguous()
    n = dets.shape[0]
    if dets.dim() != 2 or dets.shape[1] != 4 or scores.numel() != n:
        raise RuntimeError("nms: dets must be [N,4] and scores [N]")
    with torch.cuda.device(dets.device):
        ws_bytes = lib.mrb_nms_workspace_bytes(_c_int(n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dets.device)
        keep = torch.empty(n, dtype=torch.int64, device=dets.device)
        cnt = torch.empty(1, dtype=torch.int32, device=dets.device)
        check(lib.mrb_nms(_ptr(dets), _ptr(scores), _c_int(n), _c_float(threshold), _ptr(keep), _ptr(cnt),
                          _ptr(ws), _c_size_t(ws_bytes), _stream()), "mrb_nms")
        k = int(cnt.item())  # 4-byte D2H, only to size the returned tensor
    return keep[:k]


# ------------------------------------------------------------------------------------ ROIAlign
def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    """csrc/ROIAlign.h:11-25 -> [R, C, pooled_height, pooled_width]"""
    _require_cuda("roi_align_forward", input, rois)
    _require_f32("roi_align_forward", input, rois)
    n, c, h, w = input.shape
    r = rois.shape[0]
    out = torch.empty((r, c, pooled_height, pooled_width), dtype=input.dtype, device=input.device)
    if out.numel() == 0:
        return out
    x, layout = _layout_of(input)
    rois = rois.contiguous()
    with torch.cuda.device(input.device):
        check(lib.mrb_roi_align_fwd(_ptr(x), _ptr(rois), _ptr(out), _c_int(r), _c_int(n), _c_int(c), _c_int(h),
                                    _c_int(w), _c_int(pooled_height), _c_int(pooled_width),
                                    _c_float(spatial_scale), _c_int(sampling_ratio), _c_int(layout), _stream()),
              "mrb_roi_align_fwd")
    return out


def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height,
                       width, sampling_ratio):
    """csrc/ROIAlign.h:27-45 -> grad_input [batch_size, channels, height, width]"""
    _require_cuda("roi_align_backward", grad, rois)
    _require_f32("roi_align_backward", grad, rois)
    gin = torch.empty((batch_size, channels, height, width), dtype=grad.dtype, device=grad.device)
    if gin.numel() == 0:
        return gin
    grad = grad.contiguous()
    rois = rois.contiguous()
    with torch.cuda.device(grad.device):
        check(lib.mrb_roi_align_bwd(_ptr(grad), _ptr(rois), _ptr(gin), _c_int(rois.shape[0]), _c_int(batch_size),
                                    _c_int(channels), _c_int(height), _c_int(width), _c_int(pooled_height),
                                    _c_int(pooled_width), _c_float(spatial_scale), _c_int(sampling_ratio),
                                    _c_int(0), _stream()), "mrb_roi_align_bwd")
    return gin


# ------------------------------------------------------------------------------------- ROIPool
def roi_pool_forward(input, rois, spatial_scale, pooled_height, pooled_width):
    """csrc/ROIPool.h:11-24 -> (output, argmax int32)"""
    _require_cuda("roi_pool_forward", input, rois)
    _require_f32("roi_pool_forward", input, rois)
    n, c, h, w = input.shape
    r = rois.shape[0]
    out = torch.empty((r, c, pooled_height, pooled_width), dtype=input.dtype, device=input.device)
    argmax = torch.zeros((r, c, pooled_height, pooled_width), dtype=torch.int32, device=input.device)
    if out.numel() == 0:
        return out, argmax
    x = input.contiguous()
    rois = rois.contiguous()
    with torch.cuda.device(input.device):
        check(lib.mrb_roi_pool_fwd(_ptr(x), _ptr(rois), _ptr(out), _ptr(argmax), _c_int(r), _c_int(n), _c_int(c),
                                   _c_int(h), _c_int(w), _c_int(pooled_height), _c_int(pooled_width),
                                   _c_float(spatial_scale), _stream()), "mrb_roi_pool_fwd")
    return out, argmax


def roi_pool_backward(grad, input, rois, argmax, spatial_scale, pooled_height, pooled_width, batch_size,
                      channels, height, width):
    """csrc/ROIPool.h:26-45 -> grad_input"""
    _require_cuda("roi_pool_backward", grad, rois, argmax)
    _require_f32("roi_pool_backward", grad, rois)
    if argmax.dtype != torch.int32:
        raise RuntimeError("roi_pool_backward: argmax must be int32")
    gin = torch.empty((batch_size, channels, height, width), dtype=grad.dtype, device=grad.device)
    if gin.numel() == 0:
        return gin
    grad = grad.contiguous()
    rois = rois.contiguous()
    argmax = argmax.contiguous()
    with torch.cuda.device(grad.device):
        check(lib.mrb_roi_pool_bwd(_ptr(grad), _ptr(rois), _ptr(argmax), _ptr(gin), _c_int(rois.shape[0]),
                                   _c_int(batch_size), _c_int(channels), _c_int(height), _c_int(width),
                                   _c_int(pooled_height), _c_int(pooled_width), _stream()), "mrb_roi_pool_bwd")
    return gin


# ---------------------------------------------------------------------------- SigmoidFocalLoss
def _focal_args(name, logits, targets, num_classes):
    _require_cuda(name, logits, targets)
    _require_f32(name, logits)
    if logits.dim() != 2:
        raise RuntimeError("logits should be NxClass")
    if logits.shape[1] != num_classes:
        raise RuntimeError("logits.size(1) should be num_classes")
    if targets.dtype != torch.int32:
        raise RuntimeError("%s: targets must be int32 (reference reads them as int, "
                           "SigmoidFocalLoss_cuda.cu:133)" % name)
    return logits.contiguous(), targets.contiguous()


def sigmoid_focalloss_forward(logits, targets, num_classes, gamma, alpha):
    """csrc/SigmoidFocalLoss.h:10-24 -> losses [A, num_classes]"""
    logits, targets = _focal_args("sigmoid_focalloss_forward", logits, targets, num_classes)
    out = torch.empty_like(logits)
    if out.numel() == 0:
        return out
    with torch.cuda.device(logits.device):
        check(lib.mrb_sigmoid_focal_fwd(_ptr(logits), _ptr(targets), _ptr(out), _c_int64(logits.shape[0]),
                                        _c_int(num_classes), _c_float(gamma), _c_float(alpha), _stream()),
              "mrb_sigmoid_focal_fwd")
    return out


def sigmoid_focalloss_backward(logits, targets, d_losses, num_classes, gamma, alpha):
    """csrc/SigmoidFocalLoss.h:26-41 -> d_logits [A, num_classes]"""
    logits, targets = _focal_args("sigmoid_focalloss_backward", logits, targets, num_classes)
    _require_cuda("sigmoid_focalloss_backward", d_losses)
    _require_f32("sigmoid_focalloss_backward", d_losses)
    d_losses = d_losses.contiguous()
    out = torch.empty_like(logits)
    if out.numel() == 0:
        return out
    with torch.cuda.device(logits.device):
        check(lib.mrb_sigmoid_focal_bwd(_ptr(logits), _ptr(targets), _ptr(d_losses), _ptr(out),
                                        _c_int64(logits.shape[0]), _c_int(num_classes), _c_float(gamma),
                                        _c_float(alpha), _stream()), "mrb_sigmoid_focal_bwd")
    return out


# ------------------------------------------------------------------------ deformable conv (v1)
def _dcn_params(input, weight, kH, kW, dH, dW, padH, padW, dilH, dilW, group, deformable_group):
    p = DcnParams()
    p.batch, p.cin, p.height, p.width = (int(s) for s in input.shape)
    p.cout = int(weight.shape[0])
    p.kh, p.kw = int(kH), int(kW)
    p.stride_h, p.stride_w = int(dH), int(dW)
    p.pad_h, p.pad_w = int(padH), int(padW)
    p.dil_h, p.dil_w = int(dilH), int(dilW)
    p.groups, p.deformable_groups = int(group), int(deformable_group)
    return p


def _dcn_out_hw(p):
    ho = (p.height + 2 * p.pad_h - (p.dil_h * (p.kh - 1) + 1)) // p.stride_h + 1
    wo = (p.width + 2 * p.pad_w - (p.dil_w * (p.kw - 1) + 1)) // p.stride_w + 1
    return ho, wo


def _dcn_ws(p, device):
    nbytes = lib.mrb_deform_conv_workspace_bytes(ctypes.byref(p))
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device), nbytes


def _dcn_check_shapes(name, p, input, offset, weight, mask=None):
    ho, wo = _dcn_out_hw(p)
    if weight.shape[1] * p.groups != p.cin or tuple(weight.shape[2:]) != (p.kh, p.kw):
        raise RuntimeError("%s: weight shape %s does not match input channels %d / kernel (%d,%d)" %
                           (name, tuple(weight.shape), p.cin, p.kh, p.kw))
    if tuple(offset.shape) != (p.batch, p.deformable_groups * 2 * p.kh * p.kw, ho, wo):
        raise RuntimeError("%s: invalid offset shape %s, expected %s" %
                           (name, tuple(offset.shape), (p.batch, p.deformable_groups * 2 * p.kh * p.kw, ho, wo)))
    if mask is not None and tuple(mask.shape) != (p.batch, p.deformable_groups * p.kh * p.kw, ho, wo):
        raise RuntimeError("%s: invalid mask shape %s" % (name, tuple(mask.shape)))
    return ho, wo


def deform_conv_forward(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW,
                        dilationH, group, deformable_group, im2col_step):
    """csrc/deform_conv.h:11-43.  Note W-before-H argument order.  Writes `output` in place;
    `columns` / `ones` are ignored (no column matrix is materialised)."""
    _require_cuda("deform_conv_forward", input, weight, offset, output)
    _require_f32("deform_conv_forward", input, weight, offset, output)
    squeeze = input.dim() == 3
    if squeeze:
        input = input.unsqueeze(0)
        offset = offset.unsqueeze(0)
    input, weight, offset = input.contiguous(), weight.contiguous(), offset.contiguous()
    p = _dcn_params(input, weight, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group)
    ho, wo = _dcn_check_shapes("deform_conv_forward", p, input, offset, weight)
    out = output
    want = (p.batch, p.cout, ho, wo)
    if tuple(out.shape) != want or not out.is_contiguous():
        out.resize_(want)
    with torch.cuda.device(input.device):
        ws, nbytes = _dcn_ws(p, input.device)
        check(lib.mrb_deform_conv_fwd(ctypes.byref(p), _ptr(input), _ptr(offset), _c_void_p(0), _ptr(weight),
                                      _c_void_p(0), _ptr(out), _ptr(ws), _c_size_t(nbytes), _stream()),
              "mrb_deform_conv_fwd")
    if squeeze:
        output.resize_((p.cout, ho, wo))
    return 1


def deform_conv_backward_input(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW,
                               dH, padW, padH, dilationW, dilationH, group, deformable_group, im2col_step):
    """csrc/deform_conv.h:45-78.  Overwrites gradInput and gradOffset."""
    _require_cuda("deform_conv_backward_input", input, offset, gradOutput, gradInput, gradOffset, weight)
    _require_f32("deform_conv_backward_input", input, offset, gradOutput, gradInput, gradOffset, weight)
    input, weight, offset = input.contiguous(), weight.contiguous(), offset.contiguous()
    gradOutput = gradOutput.contiguous()
    p = _dcn_params(input, weight, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group)
    _dcn_check_shapes("deform_conv_backward_input", p, input, offset, weight)
    if not gradInput.is_contiguous() or gradInput.shape != input.shape:
        gradInput.resize_(input.shape)
    if not gradOffset.is_contiguous() or gradOffset.shape != offset.shape:
        gradOffset.resize_(offset.shape)
    with torch.cuda.device(input.device):
        ws, nbytes = _dcn_ws(p, input.device)
        check(lib.mrb_deform_conv_bwd(ctypes.byref(p), _ptr(input), _ptr(offset), _c_void_p(0), _ptr(weight),
                                      _ptr(gradOutput), _ptr(gradInput), _ptr(gradOffset), _c_void_p(0),
                                      _c_void_p(0), _c_void_p(0), _c_float(1.0), _ptr(ws), _c_size_t(nbytes),
                                      _stream()), "mrb_deform_conv_bwd(input)")
    return 1


def deform_conv_backward_parameters(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW,
                                    padH, dilationW, dilationH, group, deformable_group, scale, im2col_step):
    """csrc/deform_conv.h:80-113.  Accumulates scale * dW into gradWeight."""
    _require_cuda("deform_conv_backward_parameters", input, offset, gradOutput, gradWeight)
    _require_f32("deform_conv_backward_parameters", input, offset, gradOutput, gradWeight)
    input, offset, gradOutput = input.contiguous(), offset.contiguous(), gradOutput.contiguous()
    if not gradWeight.is_contiguous():
        raise RuntimeError("deform_conv_backward_parameters: gradWeight must be contiguous")
    p = _dcn_params(input, gradWeight, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group)
    _dcn_check_shapes("deform_conv_backward_parameters", p, input, offset, gradWeight)
    with torch.cuda.device(input.device):
        ws, nbytes = _dcn_ws(p, input.device)
        check(lib.mrb_deform_conv_bwd(ctypes.byref(p), _ptr(input), _ptr(offset), _c_void_p(0), _c_void_p(0),
                                      _ptr(gradOutput), _c_void_p(0), _c_void_p(0), _c_void_p(0),
                                      _ptr(gradWeight), _c_void_p(0), _c_float(scale), _ptr(ws),
                                      _c_size_t(nbytes), _stream()), "mrb_deform_conv_bwd(parameters)")
    return 1


# ------------------------------------------------------------------------ deformable conv (v2)
def modulated_deform_conv_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w,
                                  stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                                  deformable_group, with_bias):
    """csrc/deform_conv.h:115-150.  H-before-W argument order.  Writes `output` in place."""
    _require_cuda("modulated_deform_conv_forward", input, weight, offset, mask, output)
    _require_f32("modulated_deform_conv_forward", input, weight, offset, mask, output)
    if not input.is_contiguous() or not weight.is_contiguous():
        # deform_conv_cuda.cu:504-505
        raise RuntimeError("input tensor has to be contiguous" if not input.is_contiguous()
                           else "weight tensor has to be contiguous")
    offset, mask = offset.contiguous(), mask.contiguous()
    p = _dcn_params(input, weight, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                    group, deformable_group)
    ho, wo = _dcn_check_shapes("modulated_deform_conv_forward", p, input, offset, weight, mask)
    want = (p.batch, p.cout, ho, wo)
    if tuple(output.shape) != want or not output.is_contiguous():
        output.resize_(want)
    b = bias.contiguous() if with_bias else None
    with torch.cuda.device(input.device):
        ws, nbytes = _dcn_ws(p, input.device)
        check(lib.mrb_deform_conv_fwd(ctypes.byref(p), _ptr(input), _ptr(offset), _ptr(mask), _ptr(weight),
                                      _ptr(b), _ptr(output), _ptr(ws), _c_size_t(nbytes), _stream()),
              "mrb_deform_conv_fwd(modulated)")


def modulated_deform_conv_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight,
                                   grad_bias, grad_offset, grad_mask, grad_output, kernel_h, kernel_w, stride_h,
                                   stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group,
                                   with_bias):
    """csrc/deform_conv.h:152-191.  grad_input/grad_offset/grad_mask overwritten; grad_weight/grad_bias
    accumulated (callers pass zeros: layers/dcn/deform_conv_func.py:218-224)."""
    ts = (input, weight, offset, mask, grad_input, grad_weight, grad_offset, grad_mask, grad_output)
    _require_cuda("modulated_deform_conv_backward", *ts)
    _require_f32("modulated_deform_conv_backward", *ts)
    if not input.is_contiguous() or not weight.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous" if not input.is_contiguous()
                           else "weight tensor has to be contiguous")
    offset, mask, grad_output = offset.contiguous(), mask.contiguous(), grad_output.contiguous()
    p = _dcn_params(input, weight, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                    group, deformable_group)
    _dcn_check_shapes("modulated_deform_conv_backward", p, input, offset, weight, mask)
    for g, ref in ((grad_input, input), (grad_offset, offset), (grad_mask, mask), (grad_weight, weight)):
        if g.shape != ref.shape or not g.is_contiguous():
            raise RuntimeError("modulated_deform_conv_backward: gradient buffers must be contiguous and "
                               "shaped like their primal")
    gb = grad_bias if with_bias else None
    with torch.cuda.device(input.device):
        ws, nbytes = _dcn_ws(p, input.device)
        check(lib.mrb_deform_conv_bwd(ctypes.byref(p), _ptr(input), _ptr(offset), _ptr(mask), _ptr(weight),
                                      _ptr(grad_output), _ptr(grad_input), _ptr(grad_offset), _ptr(grad_mask),
                                      _ptr(grad_weight), _ptr(gb), _c_float(1.0), _ptr(ws), _c_size_t(nbytes),
                                      _stream()), "mrb_deform_conv_bwd(modulated)")


# ------------------------------------------------------------------ deformable PS-ROI pooling
def deform_psroi_pooling_forward(input, bbox, trans, out, top_count, no_trans, spatial_scale, output_dim,
                                 group_size, pooled_size, part_size, sample_per_part, trans_std):
    """csrc/deform_pool.h:11-39.  Writes `out` and `top_count` in place."""
    _require_cuda("deform_psroi_pooling_forward", input, bbox, out, top_count)
    _require_f32("deform_psroi_pooling_forward", input, bbox, out, top_count)
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")  # deform_pool_cuda.cu:45
    n, c, h, w = input.shape
    r = bbox.shape[0]
    if c != output_dim * group_size * group_size:
        raise RuntimeError("input channels must equal output_dim*group_size^2")
    bbox = bbox.contiguous()
    tr = trans.contiguous() if (not no_trans and trans.numel() > 0) else None
    for t in (out, top_count):
        if tuple(t.shape) != (r, output_dim, pooled_size, pooled_size) or not t.is_contiguous():
            raise RuntimeError("deform_psroi_pooling_forward: out/top_count must be contiguous "
                               "[R, output_dim, pooled, pooled]")
    if r == 0:
        return
    with torch.cuda.device(input.device):
        check(lib.mrb_deform_psroi_fwd(_ptr(input), _ptr(bbox), _ptr(tr), _ptr(out), _ptr(top_count), _c_int(n),
                                       _c_int(c), _c_int(h), _c_int(w), _c_int(r), _c_int(int(no_trans)),
                                       _c_int(2 if tr is None else int(tr.shape[1])),
                                       _c_float(spatial_scale), _c_int(output_dim), _c_int(group_size),
                                       _c_int(pooled_size), _c_int(part_size), _c_int(sample_per_part),
                                       _c_float(trans_std), _stream()), "mrb_deform_psroi_fwd")


def deform_psroi_pooling_backward(out_grad, input, bbox, trans, top_count, input_grad, trans_grad, no_trans,
                                  spatial_scale, output_dim, group_size, pooled_size, part_size, sample_per_part,
                                  trans_std):
    """csrc/deform_pool.h:41-70.  ACCUMULATES into input_grad / trans_grad (callers pass zeros:
    layers/dcn/deform_pool_func.py:69-70; deform_pool_kernel_cuda.cu:241-260 uses atomicAdd)."""
    _require_cuda("deform_psroi_pooling_backward", out_grad, input, bbox, top_count, input_grad)
    _require_f32("deform_psroi_pooling_backward", out_grad, input, bbox, top_count, input_grad)
    if not out_grad.is_contiguous() or not input.is_contiguous():
        raise RuntimeError("out_grad/input tensor has to be contiguous")  # deform_pool_cuda.cu:62-63
    n, c, h, w = input.shape
    r = bbox.shape[0]
    bbox = bbox.contiguous()
    use_trans = (not no_trans) and trans.numel() > 0
    tr = trans.contiguous() if use_trans else None
    tg = trans_grad if use_trans else None
    if not input_grad.is_contiguous() or (tg is not None and not tg.is_contiguous()):
        raise RuntimeError("gradient buffers must be contiguous")
    if r == 0:
        return
    with torch.cuda.device(input.device):
        check(lib.mrb_deform_psroi_bwd(_ptr(out_grad), _ptr(input), _ptr(bbox), _ptr(tr), _ptr(top_count),
                                       _ptr(input_grad), _ptr(tg), _c_int(n), _c_int(c), _c_int(h), _c_int(w),
                                       _c_int(r), _c_int(int(no_trans)),
                                       _c_int(2 if tr is None else int(tr.shape[1])), _c_float(spatial_scale),
                                       _c_int(output_dim), _c_int(group_size), _c_int(pooled_size),
                                       _c_int(part_size), _c_int(sample_per_part), _c_float(trans_std),
                                       _stream()), "mrb_deform_psroi_bwd")


__all__ = [
    "nms", "roi_align_forward", "roi_align_backward", "roi_pool_forward", "roi_pool_backward",
    "sigmoid_focalloss_forward", "sigmoid_focalloss_backward", "deform_conv_forward",
    "deform_conv_backward_input", "deform_conv_backward_parameters", "modulated_deform_conv_forward",
    "modulated_deform_conv_backward", "deform_psroi_pooling_forward", "deform_psroi_pooling_backward",
]
