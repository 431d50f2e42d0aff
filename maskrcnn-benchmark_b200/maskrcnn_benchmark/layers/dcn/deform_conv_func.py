"""Autograd functions for DCNv1 / DCNv2 (reference layers/dcn/deform_conv_func.py:9-258).

Both call the `_C` entry points with the reference's argument orders (v1: W before H; v2: H before W)
so that `_C` stays interchangeable with the reference extension."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from maskrcnn_benchmark import _C


def _conv_output_size(input, weight, padding, dilation, stride):
    size = [input.size(0), weight.size(0)]
    for d in range(input.dim() - 2):
        kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
        size.append((input.size(d + 2) + 2 * padding[d] - kernel) // stride[d] + 1)
    if not all(s > 0 for s in size):
        raise ValueError("convolution input is too small (output would be {})".format("x".join(map(str, size))))
    return tuple(size)


class DeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        if not input.is_cuda:
            raise NotImplementedError
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        ctx.groups, ctx.deformable_groups, ctx.im2col_step = groups, deformable_groups, im2col_step
        ctx.save_for_backward(input, offset, weight)
        output = input.new_empty(_conv_output_size(input, weight, ctx.padding, ctx.dilation, ctx.stride))
        ctx.bufs_ = [input.new_empty(0), input.new_empty(0)]  # columns, ones: unused by this backend
        step = min(ctx.im2col_step, input.shape[0])
        assert input.shape[0] % step == 0, "im2col step must divide batchsize"
        _C.deform_conv_forward(
            input, weight, offset, output, ctx.bufs_[0], ctx.bufs_[1], weight.size(3), weight.size(2),
            ctx.stride[1], ctx.stride[0], ctx.padding[1], ctx.padding[0], ctx.dilation[1], ctx.dilation[0],
            ctx.groups, ctx.deformable_groups, step)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        if not grad_output.is_cuda:
            raise NotImplementedError
        grad_input = grad_offset = grad_weight = None
        step = min(ctx.im2col_step, input.shape[0])
        assert input.shape[0] % step == 0, "im2col step must divide batchsize"
        geom = (weight.size(3), weight.size(2), ctx.stride[1], ctx.stride[0], ctx.padding[1], ctx.padding[0],
                ctx.dilation[1], ctx.dilation[0], ctx.groups, ctx.deformable_groups)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_input = torch.zeros_like(input, memory_format=torch.contiguous_format)
            grad_offset = torch.zeros_like(offset, memory_format=torch.contiguous_format)
            _C.deform_conv_backward_input(input, offset, grad_output, grad_input, grad_offset, weight,
                                          ctx.bufs_[0], *geom, step)
        if ctx.needs_input_grad[2]:
            grad_weight = torch.zeros_like(weight, memory_format=torch.contiguous_format)
            _C.deform_conv_backward_parameters(input, offset, grad_output, grad_weight, ctx.bufs_[0],
                                               ctx.bufs_[1], *geom, 1, step)
        return grad_input, grad_offset, grad_weight, None, None, None, None, None

    _output_size = staticmethod(_conv_output_size)


class ModulatedDeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        if not input.is_cuda:
            raise NotImplementedError
        ctx.stride, ctx.padding, ctx.dilation = stride, padding, dilation
        ctx.groups, ctx.deformable_groups = groups, deformable_groups
        ctx.with_bias = bias is not None
        # `_C.modulated_deform_conv_*` insist on NCHW-contiguous input / weight (deform_conv_cuda.cu:504-505).  Every tensor
        # of the reference is; here the producer may be an engine conv (channels_last), so normalise at the call site
        input, weight = input.contiguous(), weight.contiguous()
        if not ctx.with_bias:
            bias = input.new_empty(1)  # placeholder, never read
        if weight.requires_grad or mask.requires_grad or offset.requires_grad or input.requires_grad:
            ctx.save_for_backward(input, offset, mask, weight, bias)
        output = input.new_empty(ModulatedDeformConvFunction._infer_shape(ctx, input, weight))
        ctx._bufs = [input.new_empty(0), input.new_empty(0)]
        _C.modulated_deform_conv_forward(
            input, weight, bias, ctx._bufs[0], offset, mask, output, ctx._bufs[1], weight.shape[2],
            weight.shape[3], ctx.stride, ctx.stride, ctx.padding, ctx.padding, ctx.dilation, ctx.dilation,
            ctx.groups, ctx.deformable_groups, ctx.with_bias)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight, bias = ctx.saved_tensors
        cf = torch.contiguous_format
        grad_input = torch.zeros_like(input, memory_format=cf)
        grad_offset = torch.zeros_like(offset, memory_format=cf)
        grad_mask = torch.zeros_like(mask, memory_format=cf)
        grad_weight = torch.zeros_like(weight, memory_format=cf)
        grad_bias = torch.zeros_like(bias)
        _C.modulated_deform_conv_backward(
            input, weight, bias, ctx._bufs[0], offset, mask, ctx._bufs[1], grad_input, grad_weight, grad_bias,
            grad_offset, grad_mask, grad_output, weight.shape[2], weight.shape[3], ctx.stride, ctx.stride,
            ctx.padding, ctx.padding, ctx.dilation, ctx.dilation, ctx.groups, ctx.deformable_groups,
            ctx.with_bias)
        if not ctx.with_bias:
            grad_bias = None
        return grad_input, grad_offset, grad_mask, grad_weight, grad_bias, None, None, None, None, None

    @staticmethod
    def _infer_shape(ctx, input, weight):
        kh, kw = weight.shape[2:4]
        ho = (input.shape[2] + 2 * ctx.padding - (ctx.dilation * (kh - 1) + 1)) // ctx.stride + 1
        wo = (input.shape[3] + 2 * ctx.padding - (ctx.dilation * (kw - 1) + 1)) // ctx.stride + 1
        return input.size(0), weight.size(0), ho, wo


deform_conv = DeformConvFunction.apply
modulated_deform_conv = ModulatedDeformConvFunction.apply
