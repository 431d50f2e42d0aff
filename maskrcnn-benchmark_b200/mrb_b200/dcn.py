"""Deformable convolution (v1 / v2) of the model graph on the tensor cores.

Reference: DFConv2d (layers/misc.py:114-203) = offset conv -> DeformConv / ModulatedDeformConv
(layers/dcn/deform_conv_func.py:9-258 -> csrc/cuda/deform_conv_cuda.cu: fp32 im2col `columns` + cuBLAS GEMM,
per-image loop for v2), followed in Bottleneck.forward by FrozenBatchNorm2d and ReLU (resnet.py:332-334).

Here: bilinear sampler -> bf16 columns [pixels, 9C] (csrc/dcn_nhwc.cu) -> tcgen05 GEMM with the frozen-BN scale/shift and
the ReLU in its epilogue (the 3x3 KRSC filter IS the [Cout, 9C] B operand: no copy); backward = two tcgen05 GEMMs (column
gradient, weight gradient) + one scatter/coordinate kernel.  bf16 operands, fp32 offsets / bilinear weights /
accumulation: within 1e-2 of the fp32 reference (tests/test_dcn_gpu.py); the `_C.deform_conv_*` entry points keep the fp32
path (1e-4)."""
import torch


class _DcnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, om, weight, w16, scale, shift, relu, modulated, stride, pad, be, wsink):
        from mrb_b200 import ops
        co, ci, kh, kw = w16.shape
        cols = ops.dcn_sample_nhwc(x, om, kh, stride, pad, 1, modulated)
        wv = torch.as_strided(w16, (co, kh * kw * ci, 1, 1), (kh * kw * ci, 1, 1, 1), w16.storage_offset())
        y = ops.conv2d_fwd(cols, wv, scale, shift, None, 1, 0, relu)
        ctx.cfg = (relu, modulated, stride, pad, kh, tuple(w16.shape))
        ctx.be, ctx.wsink = be, wsink
        ctx.save_for_backward(x, om, cols, wv, scale, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, g):
        from mrb_b200 import ops
        x, om, cols, wv, scale, y = ctx.saved_tensors
        relu, modulated, stride, pad, k, wshape = ctx.cfg
        co, ci = wshape[0], wshape[1]
        if relu:
            g = torch.where(y > 0, g, torch.zeros((), dtype=g.dtype, device=g.device))
        g = g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gx = gom = gw = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            gcols = ops.conv2d_dgrad(g, wv, tuple(cols.shape), scale, None, None, 1, 0)
            gx32, gom = ops.dcn_backward_nhwc(x, om, gcols, k, stride, pad, 1, modulated, need_grad_x=ctx.needs_input_grad[0])
            if gx32 is not None:
                gx = gx32.to(torch.bfloat16)
        if ctx.needs_input_grad[2]:
            if ctx.wsink is not None:
                sink = torch.as_strided(ctx.wsink, (co, k * k * ci, 1, 1), (k * k * ci, 1, 1, 1), ctx.wsink.storage_offset())
                ctx.be.side_launch((cols, g), lambda: ops.conv2d_wgrad(cols, g, (co, k * k * ci, 1, 1), 1, 0, scale, accumulate_into=sink))
            else:
                gwv = ops.conv2d_wgrad(cols, g, (co, k * k * ci, 1, 1), 1, 0, scale)      # [co, (tap, ci)] fp32
                gw = gwv.view(co, k, k, ci).permute(0, 3, 1, 2)                           # logical [co, ci, kh, kw], KRSC memory
        return gx, gom, gw, None, None, None, None, None, None, None, None, None


def deform_conv_nhwc(x, om, weight, w16, scale=None, shift=None, relu=False, modulated=False, stride=1, pad=1, be=None, wsink=None):
    """x: bf16 NHWC [N,C,H,W]; om: fp32 channels_last offsets (+ mask logits); weight: the nn.Parameter ([Cout,C,3,3]) for
    autograd; w16: its bf16 KRSC copy.  -> bf16 NHWC [N,Cout,Ho,Wo] = act(dcn(x) * scale + shift)."""
    if wsink is not None and not wsink.is_contiguous(memory_format=torch.channels_last):
        wsink = None
    return _DcnFn.apply(x, om, weight, w16, scale, shift, relu, modulated, stride, pad, be, wsink)
