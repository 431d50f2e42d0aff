"""Compute backends of the harness.

B200Backend is the product: every hot op goes through libmrb_b200.so (tcgen05 convs in bf16 NHWC,
fused multi-level ROIAlign, on-device batched NMS).  The model code is written against the small
`Backend` interface so that tests can substitute a CPU checker backend (tests/_cpu_backend.py, built
on plain PyTorch + the oracle) to validate the harness logic without a GPU; the product never does."""
import torch
import torch.nn.functional as F
from torch.autograd import Function


class Backend:
    name = "abstract"
    act_dtype = torch.float32
    channels_last = False

    def prepare_input(self, images):
        raise NotImplementedError

    def conv(self, x, weight, scale=None, shift=None, bias=None, residual=None, stride=1, pad=0, relu=False,
             out_fp32=False):
        """y = act(conv(x, weight) * scale + shift|bias + residual).  `shift` is a frozen per-channel
        constant (FrozenBatchNorm2d), `bias` a trainable conv bias; at most one of them is given."""
        raise NotImplementedError


# --------------------------------------------------------------------------------------------
class _ConvFn(Function):
    """Fused conv + per-channel affine + residual + ReLU on the tcgen05 engine, with a hand-written
    backward: ReLU mask -> dgrad on the same engine (BN scale folded into the flipped weights) -> wgrad."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, w16, scale, shift, stride, pad, relu, out_fp32, wgrad_fn):
        from mrb_b200 import ops
        add = shift if shift is not None else (bias.detach().float() if bias is not None else None)
        y = ops.conv2d_fwd(x, w16, scale, add, residual, stride, pad, relu,
                           torch.float32 if out_fp32 else torch.bfloat16)
        ctx.cfg = (stride, pad, relu, tuple(x.shape), wgrad_fn)
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, w16, scale, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        from mrb_b200 import ops
        x, w16, scale, y = ctx.saved_tensors
        stride, pad, relu, x_shape, wgrad_fn = ctx.cfg
        g = gy
        if relu:
            g = torch.where(y > 0, g, torch.zeros((), dtype=g.dtype, device=g.device))
        if g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        g = g.contiguous(memory_format=torch.channels_last)
        gx = gw = gb = gres = None
        if ctx.needs_input_grad[0]:
            gx = ops.conv2d_dgrad(g, w16, x_shape, scale, None, None, stride, pad)
        if ctx.needs_input_grad[1]:
            gw = wgrad_fn(x, g, w16, stride, pad)
            if scale is not None:
                gw = gw * scale[:, None, None, None]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g.float().sum((0, 2, 3))
        if ctx.has_res and ctx.needs_input_grad[3]:
            gres = g
        return gx, gw, gb, gres, None, None, None, None, None, None, None, None


def _wgrad_cudnn(x, g, w16, stride, pad):
    """Weight gradient.  TODO(round 2): tcgen05 MN-major split-K wgrad kernel; until then this one
    GEMM family is borrowed from the library (ATen/cuDNN) and reported as such in bench.py."""
    gw = torch.ops.aten.convolution_backward(g, x, w16, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1,
                                             [False, True, False])[1]
    return gw.float()


class B200Backend(Backend):
    name = "b200"
    act_dtype = torch.bfloat16
    channels_last = True

    def __init__(self):
        self._w16 = {}
        self.wgrad_fn = _wgrad_cudnn
        self.wgrad_impl = "aten.convolution_backward (cuDNN)"

    def _weight16(self, w):
        key = id(w)
        ent = self._w16.get(key)
        if ent is None or ent[0] != w._version or ent[1].device != w.device:
            w16 = w.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            self._w16[key] = (w._version, w16)
            return w16
        return ent[1]

    def prepare_input(self, images):
        return images.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def conv(self, x, weight, scale=None, shift=None, bias=None, residual=None, stride=1, pad=0, relu=False,
             out_fp32=False):
        if x.numel() == 0:
            n, _, h, w = x.shape
            kh, kw = weight.shape[2:]
            return x.new_zeros((n, weight.shape[0], (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1),
                               dtype=torch.float32 if out_fp32 else torch.bfloat16)
        return _ConvFn.apply(x, weight, bias, residual, self._weight16(weight), scale, shift, stride, pad, relu,
                             out_fp32, self.wgrad_fn)

    def max_pool(self, x, k, s, p):
        return F.max_pool2d(x, k, s, p)

    def upsample2x(self, x):
        return F.interpolate(x, scale_factor=2, mode="nearest")
