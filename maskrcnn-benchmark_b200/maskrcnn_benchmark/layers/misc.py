"""Empty-batch-safe nn wrappers and DFConv2d (reference layers/misc.py:19-203).

Conv2d / ConvTranspose2d are the drop-in boundary of the dense-conv hot path: for CUDA tensors their forward is
served by the tcgen05 engine of libmrb_b200.so (mrb_b200.engine), with ATen as the logged fallback for
geometries the engine does not cover; CPU tensors always take ATen (the reference's CPU path)."""
import math

import torch
from torch import nn
from torch.nn.modules.utils import _ntuple


class _NewEmptyTensorOp(torch.autograd.Function):
    """Differentiable `x.new_empty(shape)` for zero-element tensors (misc.py:19-27)."""

    @staticmethod
    def forward(ctx, x, new_shape):
        ctx.shape = x.shape
        return x.new_empty(new_shape)

    @staticmethod
    def backward(ctx, grad):
        return _NewEmptyTensorOp.apply(grad, ctx.shape), None


def _conv_out_hw(hw, padding, dilation, kernel_size, stride):
    return [(i + 2 * p - (di * (k - 1) + 1)) // d + 1
            for i, p, di, k, d in zip(hw, padding, dilation, kernel_size, stride)]


def _engine():
    from mrb_b200 import engine      # lazy: layers must import without the host-side package being touched on CPU
    return engine


class Conv2d(torch.nn.Conv2d):
    def forward(self, x):
        if x.numel() > 0:
            if x.is_cuda:
                y = _engine().conv2d_module(self, x)
                if y is not None:
                    return y
            return super().forward(x)
        hw = _conv_out_hw(x.shape[-2:], self.padding, self.dilation, self.kernel_size, self.stride)
        return _NewEmptyTensorOp.apply(x, [x.shape[0], self.weight.shape[0]] + hw)


class ConvTranspose2d(torch.nn.ConvTranspose2d):
    def forward(self, x):
        if x.numel() > 0:
            if x.is_cuda:
                y = _engine().conv_transpose2d_module(self, x)
                if y is not None:
                    return y
            return super().forward(x)
        hw = [(i - 1) * d - 2 * p + (di * (k - 1) + 1) + op
              for i, p, di, k, d, op in zip(x.shape[-2:], self.padding, self.dilation, self.kernel_size,
                                            self.stride, self.output_padding)]
        return _NewEmptyTensorOp.apply(x, [x.shape[0], self.bias.shape[0]] + hw)


class BatchNorm2d(torch.nn.BatchNorm2d):
    def forward(self, x):
        if x.numel() > 0:
            return super().forward(x)
        return _NewEmptyTensorOp.apply(x, x.shape)


def interpolate(input, size=None, scale_factor=None, mode="nearest", align_corners=None):
    if input.numel() > 0:
        return torch.nn.functional.interpolate(input, size, scale_factor, mode, align_corners)
    if size is None and scale_factor is None:
        raise ValueError("either size or scale_factor should be defined")
    if size is not None and scale_factor is not None:
        raise ValueError("only one of size or scale_factor should be defined")
    if scale_factor is not None and isinstance(scale_factor, tuple) and len(scale_factor) != 2:
        raise ValueError("scale_factor shape must match input shape. Input is 2D, scale_factor size is %d"
                         % len(scale_factor))
    if size is not None:
        out_hw = tuple(size) if isinstance(size, (tuple, list)) else (size, size)
    else:
        sf = _ntuple(2)(scale_factor)
        out_hw = tuple(int(math.floor(input.size(i + 2) * sf[i])) for i in range(2))
    return _NewEmptyTensorOp.apply(input, tuple(input.shape[:-2]) + out_hw)


class DFConv2d(nn.Module):
    """Deformable conv block: an ordinary conv predicts offsets (and masks), a (modulated)
    deformable conv consumes them.  Submodule names `offset` / `conv` are checkpoint keys
    (utils/c2_model_loading.py:146-170)."""

    def __init__(self, in_channels, out_channels, with_modulated_dcn=True, kernel_size=3, stride=1, groups=1,
                 dilation=1, deformable_groups=1, bias=False):
        super().__init__()
        if isinstance(kernel_size, (list, tuple)):
            assert isinstance(stride, (list, tuple)) and isinstance(dilation, (list, tuple))
            assert len(kernel_size) == 2 and len(stride) == 2 and len(dilation) == 2
            padding = (dilation[0] * (kernel_size[0] - 1) // 2, dilation[1] * (kernel_size[1] - 1) // 2)
            taps = kernel_size[0] * kernel_size[1]
        else:
            padding = dilation * (kernel_size - 1) // 2
            taps = kernel_size * kernel_size
        from maskrcnn_benchmark.layers import DeformConv, ModulatedDeformConv
        conv_block = ModulatedDeformConv if with_modulated_dcn else DeformConv
        offset_channels = taps * (3 if with_modulated_dcn else 2)
        self.offset = Conv2d(in_channels, deformable_groups * offset_channels, kernel_size=kernel_size,
                             stride=stride, padding=padding, groups=1, dilation=dilation)
        nn.init.kaiming_uniform_(self.offset.weight, a=1)
        nn.init.constant_(self.offset.bias, 0.)
        self.conv = conv_block(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                               padding=padding, dilation=dilation, groups=groups,
                               deformable_groups=deformable_groups, bias=bias)
        self.with_modulated_dcn = with_modulated_dcn
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation

    def forward(self, x):
        if x.numel() > 0:
            if not self.with_modulated_dcn:
                return self.conv(x, self.offset(x))
            om = self.offset(x)
            # misc.py:186-187: hard-wired 3x3, deformable_groups=1 split
            return self.conv(x, om[:, :18, :, :], om[:, -9:, :, :].sigmoid())
        ks, st, pd, dl = (_ntuple(2)(v) for v in (self.kernel_size, self.stride, self.padding, self.dilation))
        hw = _conv_out_hw(x.shape[-2:], pd, dl, ks, st)
        return _NewEmptyTensorOp.apply(x, [x.shape[0], self.conv.weight.shape[0]] + hw)
