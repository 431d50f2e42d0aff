"""Grouped 3x3 convolution (ResNeXt / X-101-32x8d, reference modeling/backbone/resnet.py:283-296 `groups=num_groups`)
on the dense tcgen05 engine.

A group of Cg = C / groups <= 64 channels is a diagonal block of a 64 -> 64 "super-group": the layer is C / 64
independent dense 64 -> 64 convolutions whose weights are block-diagonal (MMA work x 64/Cg, no extra memory traffic:
the tensor core is not what bounds these layers).  Round 1 composed the layer from one call per super-group on strided
channel windows (kept as MRB_GROUPED=composed); the native form is one launch per direction (include/mrb_b200.h
MRB_CONV_GROUPED64).  The weight expansion is pinned on CPU (tests/test_grouped_cpu.py), the kernels against fp32
F.conv2d(groups=...) on the B200 (tests/test_zz_grouped_gpu.py)."""
import os

import torch

SG = 64      # channels of a super-group = one k-block of the engine


def check_geometry(cin, cout, groups):
    if cin != cout or cin % groups or cin % SG or SG % (cin // groups):
        raise RuntimeError("grouped conv: need Cin == Cout, Cin %% 64 == 0 and (Cin/groups) | 64 (got %d, %d, %d)"
                           % (cin, cout, groups))
    return cin // groups, cin // SG


def expand_group_weights(w, groups):
    """[C, C/groups, kh, kw] grouped weights -> [C, 64, kh, kw]: row co holds its group's filter at the position of that
    group inside co's 64-channel super-group, zeros elsewhere (block-diagonal within each super-group)."""
    c, cg, kh, kw = w.shape
    check_geometry(c, c, groups)
    assert cg == c // groups
    co = torch.arange(c, device=w.device)
    local0 = (co // cg) * cg - (co // SG) * SG                   # first local input channel of co's group
    idx = local0[:, None] + torch.arange(cg, device=w.device)[None, :]        # [C, Cg]
    out = w.new_zeros((c, SG, kh, kw))
    out.scatter_(1, idx[:, :, None, None].expand(c, cg, kh, kw), w)
    return out


def collapse_group_grads(gw_exp, groups):
    """Inverse gather for gradients: [C, 64, kh, kw] (dense super-group gradients) -> [C, C/groups, kh, kw]."""
    c, _, kh, kw = gw_exp.shape
    cg = c // groups
    co = torch.arange(c, device=gw_exp.device)
    local0 = (co // cg) * cg - (co // SG) * SG
    idx = local0[:, None] + torch.arange(cg, device=gw_exp.device)[None, :]
    return torch.gather(gw_exp, 1, idx[:, :, None, None].expand(c, cg, kh, kw))


def _window(t, sg):
    """Channels [64*sg, 64*sg + 64) of an NHWC (channels_last) tensor as a strided view (no copy)."""
    n, c, h, w = t.shape
    return torch.as_strided(t, (n, SG, h, w), (h * w * c, 1, w * c, c), t.storage_offset() + sg * SG)


class GroupedConvFn(torch.autograd.Function):
    """y = act(grouped_conv(x, w) * scale + shift) on the conv engine (bf16 NHWC), ONE launch per direction:
    MRB_CONV_GROUPED64 -- the N tile index doubles as the 64-channel offset of the A operand, the weight operand is the
    block-diagonal expansion [C, 64, kh, kw].  Stride 2 (`STRIDE_IN_1X1: False`, the first block of res3..res5 only) is
    computed at stride 1 and subsampled: 3 of ~100 layers pay 4x on a layer class that is not tensor-bound.
    MRB_GROUPED=composed selects round 1's per-super-group composition (one launch per 64 channels) for A/B runs."""

    @staticmethod
    def forward(ctx, x, weight, scale, shift, groups, pad, relu, stride=1, w16=None, wsink=None):
        from mrb_b200 import ops
        c = x.shape[1]
        _, sgs = check_geometry(c, weight.shape[0], groups)
        if x.dtype != torch.bfloat16:
            raise RuntimeError("grouped conv: bf16 NHWC input required (cast at the call site)")
        x = x.contiguous(memory_format=torch.channels_last)
        if w16 is None:
            w16 = weight.detach().to(torch.bfloat16)
        w16 = w16.contiguous(memory_format=torch.channels_last)
        composed = os.environ.get("MRB_GROUPED", "native") == "composed"
        if composed:
            w_exp = expand_group_weights(w16, groups).contiguous(memory_format=torch.channels_last)
            out = torch.empty_like(x)
            for sg in range(sgs):
                sl = slice(sg * SG, (sg + 1) * SG)
                ops.conv2d_fwd(_window(x, sg), w_exp[sl], None if scale is None else scale[sl].contiguous(),
                               None if shift is None else shift[sl].contiguous(), None, 1, pad, relu, out=_window(out, sg))
        else:
            w_exp, _ = ops.grouped_expand_weights(w16, groups, None, True, False)      # one launch (csrc/grouped_prep.cu)
            out = ops.conv2d_fwd(x, w_exp, scale, shift, None, 1, pad, relu, grouped=True)
        if stride == 2:
            out = out[:, :, ::2, ::2].contiguous(memory_format=torch.channels_last)
        elif stride != 1:
            raise RuntimeError("grouped conv: stride 1 or 2")
        ctx.cfg = (groups, pad, relu, sgs, stride, composed, tuple(x.shape), tuple(weight.shape))
        ctx.wsink = wsink
        ctx.save_for_backward(x, w16, scale, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        from mrb_b200 import ops
        x, w16, scale, y = ctx.saved_tensors
        groups, pad, relu, sgs, stride, composed, x_shape, w_shape = ctx.cfg
        if relu:
            g = torch.where(y > 0, g, torch.zeros((), dtype=g.dtype, device=g.device))
        g = g.to(torch.bfloat16)
        n, c, h, w = x_shape
        k = w_shape[2]
        if stride == 2:
            hf, wf = h + 2 * pad - k + 1, w + 2 * pad - k + 1
            gf = torch.zeros((n, c, hf, wf), dtype=torch.bfloat16, device=g.device).contiguous(memory_format=torch.channels_last)
            gf[:, :, ::2, ::2] = g
            g = gf
        g = g.contiguous(memory_format=torch.channels_last)
        gx = gw = None
        if composed:
            w_exp = expand_group_weights(w16, groups).contiguous(memory_format=torch.channels_last)
        if ctx.needs_input_grad[0]:
            if composed:
                parts = [ops.conv2d_dgrad(_window(g, sg), w_exp[sg * SG:(sg + 1) * SG], (n, SG, h, w),
                                          None if scale is None else scale[sg * SG:(sg + 1) * SG].contiguous(), None, None, 1, pad)
                         for sg in range(sgs)]
                gx = torch.cat(parts, 1).contiguous(memory_format=torch.channels_last)
            else:
                _, wd = ops.grouped_expand_weights(w16, groups, scale, False, True)
                gx = ops.conv2d_dgrad_grouped(g, wd, x_shape, 1, pad)
        if ctx.needs_input_grad[1]:
            if composed or c % 128:
                if not composed:
                    w_exp = expand_group_weights(w16, groups).contiguous(memory_format=torch.channels_last)
                parts = [ops.conv2d_wgrad(_window(x, sg), _window(g, sg), (SG, SG, k, k), 1, pad,
                                          None if scale is None else scale[sg * SG:(sg + 1) * SG].contiguous())
                         for sg in range(sgs)]
                gw = collapse_group_grads(torch.cat(parts, 0), groups)
            else:
                gw128 = ops.conv2d_wgrad_grouped(x, g, k, 1, pad, scale, raw=True)
                if ctx.wsink is not None:
                    ops.grouped_collapse_wgrad(gw128, w_shape, groups, accumulate_into=ctx.wsink)   # straight into the arena
                else:
                    gw = ops.grouped_collapse_wgrad(gw128, w_shape, groups)
        return gx, gw, None, None, None, None, None, None, None, None


def conv2d_grouped(x, weight, groups, scale=None, shift=None, pad=1, relu=False, stride=1, w16=None, wsink=None):
    """w16: a current bf16 copy of `weight` (e.g. the optimizer arena's); wsink: persistent fp32 gradient accumulator of `weight`
    (then the weight gradient is ADDED there and None is returned to autograd)."""
    return GroupedConvFn.apply(x, weight, scale, shift, groups, pad, relu, stride, w16, wsink)
