"""Compute backends of the harness.

B200Backend is the product: every hot op goes through libmrb_b200.so (tcgen05 convs in bf16 NHWC,
fused multi-level ROIAlign, on-device batched NMS).  The model code is written against the small
`Backend` interface so that tests can substitute a CPU checker backend (oracle/cpu_backend.py, built
on plain PyTorch + the oracle) to validate the harness logic without a GPU; the product never does."""
import torch
import torch.nn.functional as F
from torch.autograd import Function


class Backend:
    name = "abstract"
    act_dtype = torch.float32
    channels_last = False

    def lateral_topdown(self, feat, weight, bias, top, premask_x=True):
        """FPN inner block (fpn.py:51-64): conv1x1(feat) + nearest_upsample_2x(top) (top may be None).
        premask_x: `feat` is a ReLU output whose producer expects its gradient already masked by [feat > 0]."""
        top_down = self.upsample2x(top) if top is not None else None
        return self.conv(feat, weight, bias=bias, residual=top_down, premask_x=premask_x)

    def prepare_input(self, images):
        raise NotImplementedError

    def conv(self, x, weight, scale=None, shift=None, bias=None, residual=None, stride=1, pad=0, relu=False,
             out_fp32=False, premask_x=False, gy_premasked=False):
        """y = act(conv(x, weight) * scale + shift|bias + residual).  `shift` is a frozen per-channel
        constant (FrozenBatchNorm2d), `bias` a trainable conv bias; at most one of them is given.
        Backward-fusion hints (pure optimisation, ignored by backends that use plain autograd):
        premask_x   -- x is the ReLU output of its producer: return grad_x already multiplied by [x > 0];
        gy_premasked -- every consumer of y does that, so this op skips its own ReLU mask."""
        raise NotImplementedError


# --------------------------------------------------------------------------------------------
class _ConvFn(Function):
    """Fused conv + per-channel affine + residual + ReLU on the tcgen05 engine, with a hand-written
    backward: ReLU mask -> dgrad on the same engine (BN scale folded into the flipped weights) -> wgrad."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, w16, scale, shift, stride, pad, relu, out_fp32, wgrad_fn,
                premask_x=False, gy_premasked=False, residual_up2=False, be=None, wsink=None, bsink=None):
        from mrb_b200 import ops
        add = shift if shift is not None else (bias.detach().float() if bias is not None else None)
        y = ops.conv2d_fwd(x, w16, scale, add, residual, stride, pad, relu,
                           torch.float32 if out_fp32 else torch.bfloat16, residual_up2=residual_up2)
        ctx.residual_up2 = residual_up2
        ctx.cfg = (stride, pad, relu and not gy_premasked, tuple(x.shape), wgrad_fn)
        ctx.premask_x = premask_x
        ctx.wparam = weight if isinstance(weight, torch.nn.Parameter) else None
        ctx.be = be
        ctx.wsink, ctx.bsink = wsink, bsink      # persistent fp32 accumulators of weight / bias (ParamArena) or None
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
            # premask_x: x is the ReLU output of its producer, whose backward then skips its own mask pass
            prep = ctx.be.dgrad_weights(ctx.wparam, w16, scale) if ctx.be is not None else None
            gx = ops.conv2d_dgrad(g, w16, x_shape, scale, None, x if ctx.premask_x else None, stride, pad, prepared=prep)
        if ctx.needs_input_grad[1]:
            if ctx.wsink is not None:
                sink = ctx.wsink.view(w16.shape)
                ctx.be.side_launch((x, g), lambda: ops.conv2d_wgrad(x, g, w16.shape, stride, pad, scale, accumulate_into=sink))
            else:
                gw = wgrad_fn(x, g, w16, stride, pad, scale)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            if ctx.bsink is not None:
                bsink = ctx.bsink
                ctx.be.side_launch((g,), lambda: ops.bias_grad(g, accumulate_into=bsink))
            else:
                gb = ops.bias_grad(g)
        if ctx.has_res and ctx.needs_input_grad[3]:
            # nearest-2x upsample backward == sum over each 2x2 block
            gres = ops.sum_pool2x2_nhwc(g) if ctx.residual_up2 else g
        return (gx, gw, gb, gres) + (None,) * 14


class _BucketBoundaryFn(Function):
    """Identity whose backward tells the arena that every gradient kernel of bucket `name` (everything downstream of these
    tensors up to the next boundary) has been issued: its all-reduce starts now and overlaps the rest of backward."""

    @staticmethod
    def forward(ctx, be, name, *xs):
        ctx.be, ctx.name = be, name
        return tuple(x.view_as(x) for x in xs)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.be.arena is not None:
            ctx.be.arena.reduce_bucket(ctx.name)
        return (None, None) + grads


class _HeadsBoundaryFn(Function):
    """Identity on the FPN outputs whose backward runs exactly when every consumer downstream (RPN head, both ROI
    heads) has issued its backward -- the moment the gradients of all their parameters are complete.  The backend uses
    it to start the data-parallel all-reduce of that bucket while the backbone's backward is still to come."""

    @staticmethod
    def forward(ctx, be, *feats):
        ctx.be = be
        return tuple(f.view_as(f) for f in feats)

    @staticmethod
    def backward(ctx, *grads):
        ctx.be.heads_grads_ready()
        return (None,) + grads


class _SelectChannelFn(Function):
    """out[r, h, w] = y[r, label[r], h, w] for an NHWC tensor (the class-specific mask logit of each ROI, reference
    roi_heads/mask_head/loss.py:120-126 `mask_logits[positive_inds, labels_pos]`).  The backward writes the gradient
    straight into a zeroed bf16 NHWC tensor -- the layout the conv engine's dgrad/wgrad read -- instead of going
    through index_put + slice-pad + cast + layout copies of the full 81-channel fp32 logits."""

    @staticmethod
    def forward(ctx, y, labels):
        r, c, h, w = y.shape
        idx = labels.view(r, 1, 1, 1).expand(r, h, w, 1)
        ctx.save_for_backward(idx)
        ctx.shape = (r, c, h, w)
        return torch.gather(y.permute(0, 2, 3, 1), 3, idx).squeeze(3)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        r, c, h, w = ctx.shape
        gy = torch.empty((r, c, h, w), dtype=torch.bfloat16, device=g.device, memory_format=torch.channels_last).zero_()
        gy.permute(0, 2, 3, 1).scatter_(3, idx, g.to(torch.bfloat16).unsqueeze(3))   # g arrives in y's dtype
        return gy, None


class _Deconv2x2Fn(Function):
    """ConvTranspose2d(k=2, s=2) (+bias, +ReLU) of the mask head (reference roi_heads/mask_head/
    roi_mask_predictors.py:17-35) on the conv engine with NO pixel shuffle: out[r, 2h+i, 2w+j, :] = W[:, :, i, j]^T x[r, h, w, :]
    is, for a fixed row parity i, a 1x1 convolution with 2*Cout outputs (j, co) whose result rows are the odd/even rows
    of the NHWC output -- a strided window (row pitch 4*W*Cout, pixel pitch 2*Cout) that the epilogue writes in place.
    Backward reads the same windows of grad_out through TMA: dgrad_0 + dgrad_1 (accumulated in the epilogue, ReLU mask
    of the producer on the second launch), two weight gradients, one bias reduction."""

    @staticmethod
    def _windows(t, r, c, h, w):
        base = t.storage_offset()
        return [torch.as_strided(t, (1, 2 * c, r * h, w), (r * h * w * 4 * c, 1, 4 * w * c, 2 * c), base + i * 2 * w * c)
                for i in (0, 1)]

    @staticmethod
    def forward(ctx, x, weight, bias, relu, premask_x, gy_premasked):
        from mrb_b200 import ops
        r, cin, h, w = x.shape
        cout = weight.shape[1]
        x = x.contiguous(memory_format=torch.channels_last)
        xv = torch.as_strided(x, (1, cin, r * h, w), (r * h * w * cin, 1, w * cin, cin), x.storage_offset())
        w2 = weight.detach().permute(2, 3, 1, 0).reshape(2, 2 * cout, cin, 1, 1).to(torch.bfloat16)   # [i][(j, co)][ci]
        b2 = bias.detach().float().repeat(2) if bias is not None else None
        out = torch.empty((r, cout, 2 * h, 2 * w), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        for i, ov in enumerate(_Deconv2x2Fn._windows(out, r, cout, h, w)):
            ops.conv2d_fwd(xv, w2[i], None, b2, None, 1, 0, relu, out=ov)
        ctx.cfg = (relu and not gy_premasked, premask_x, bias is not None, (r, cin, h, w), cout)
        ctx.save_for_backward(xv, w2, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        from mrb_b200 import ops
        xv, w2, y = ctx.saved_tensors
        relu, premask_x, has_bias, (r, cin, h, w), cout = ctx.cfg
        if relu:
            g = torch.where(y > 0, g, torch.zeros((), dtype=g.dtype, device=g.device))
        g = g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gv = _Deconv2x2Fn._windows(g, r, cout, h, w)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ops.conv2d_dgrad(gv[0], w2[0], xv.shape, None, None, None, 1, 0)
            gx = ops.conv2d_dgrad(gv[1], w2[1], xv.shape, None, None, xv if premask_x else None, 1, 0, accumulate_into=gx)
            gx = torch.as_strided(gx, (r, cin, h, w), (h * w * cin, 1, w * cin, cin), gx.storage_offset())
        if ctx.needs_input_grad[1]:
            gw2 = torch.stack([ops.conv2d_wgrad(xv, gv[i], (2 * cout, cin, 1, 1), 1, 0) for i in (0, 1)])
            gw = gw2.view(2, 2, cout, cin).permute(3, 2, 0, 1)           # [i, j, co, ci] -> [ci, co, i, j]
        if has_bias and ctx.needs_input_grad[2]:
            gb = ops.bias_grad(g)
        return gx, gw, gb, None, None, None


class _BottleneckFn(Function):
    """One ResNet bottleneck (reference modeling/backbone/resnet.py:324-344) as a single autograd node:
    4 fused forward convs and a hand-scheduled backward in which every ReLU mask, the frozen-BN scale and
    the residual join ride in the epilogue of a dgrad launch -- no standalone elementwise pass:
        g  (grad of out, already masked by out > 0 by its consumers)
        g2 = dgrad3(g)   * [y2 > 0]        dW3 = wgrad(y2, g) * s3
        g1 = dgrad2(g2)  * [y1 > 0]        dW2 = wgrad(y1, g2) * s2
        gx = (dgrad1(g1) + identity-branch grad) * [x > 0]      dW1 = wgrad(x, g1) * s1, dWd = wgrad(x, g) * sd
    The returned gx is therefore pre-masked for the producer of x (the previous bottleneck)."""

    @staticmethod
    def forward(ctx, x, w1, w2, w3, wd, be, blk, g_premasked):
        from mrb_b200 import ops
        s1, s3, sd = blk.strides
        (a1, b1), (a2, b2), (a3, b3) = (a.get() for a in blk._aff)
        w16 = [be._weight16(w) for w in (w1, w2, w3)]
        y1 = ops.conv2d_fwd(x, w16[0], a1, b1, None, s1, 0, True)
        y2 = ops.conv2d_fwd(y1, w16[1], a2, b2, None, s3, 1, True)
        if wd is not None:
            ad, bd = blk._aff_d.get()
            wd16 = be._weight16(wd)
            idn = ops.conv2d_fwd(x, wd16, ad, bd, None, sd, 0, False)
        else:
            ad, wd16, idn = None, None, x
        out = ops.conv2d_fwd(y2, w16[2], a3, b3, idn, 1, 0, True)
        ctx.be, ctx.strides, ctx.g_premasked = be, (s1, s3, sd), g_premasked
        ctx.has_d = wd is not None
        ctx.wparams = (w1, w2, w3, wd)
        ctx.save_for_backward(x, y1, y2, out, w16[0], w16[1], w16[2], wd16, a1, a2, a3, ad)
        return out

    @staticmethod
    def backward(ctx, g):
        from mrb_b200 import ops
        x, y1, y2, out, w1, w2, w3, wd, a1, a2, a3, ad = ctx.saved_tensors
        s1, s3, sd = ctx.strides
        be = ctx.be
        p1, p2, p3, pd = ctx.wparams

        def wg(xin, gout, w16, stride, pad, scale, param):
            sink = be.grad_sink(param)
            if sink is None:
                return be.wgrad_fn(xin, gout, w16, stride, pad, scale)
            be.side_launch((xin, gout), lambda: ops.conv2d_wgrad(xin, gout, w16.shape, stride, pad, scale, accumulate_into=sink))
            return None
        if not ctx.g_premasked:
            g = torch.where(out > 0, g, torch.zeros((), dtype=g.dtype, device=g.device))
        g = g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        need = ctx.needs_input_grad
        gw1 = gw2 = gw3 = gwd = gx = None
        if need[3]:
            gw3 = wg(y2, g, w3, 1, 0, a3, p3)
        g2 = ops.conv2d_dgrad(g, w3, y2.shape, a3, None, y2, 1, 0, prepared=be.dgrad_weights(p3, w3, a3))
        if need[2]:
            gw2 = wg(y1, g2, w2, s3, 1, a2, p2)
        g1 = ops.conv2d_dgrad(g2, w2, y1.shape, a2, None, y1, s3, 1, prepared=be.dgrad_weights(p2, w2, a2))
        if need[1]:
            gw1 = wg(x, g1, w1, s1, 0, a1, p1)
        if ctx.has_d and need[4]:
            gwd = wg(x, g, wd, sd, 0, ad, pd)
        if need[0]:
            pw1 = be.dgrad_weights(p1, w1, a1)
            if not ctx.has_d:
                gx = ops.conv2d_dgrad(g1, w1, x.shape, a1, g, x, s1, 0, prepared=pw1)
            elif sd == 1:
                gi = ops.conv2d_dgrad(g, wd, x.shape, ad, None, None, 1, 0, prepared=be.dgrad_weights(pd, wd, ad))
                gx = ops.conv2d_dgrad(g1, w1, x.shape, a1, gi, x, s1, 0, prepared=pw1)
            else:
                gx = ops.conv2d_dgrad(g, wd, x.shape, ad, None, None, sd, 0, prepared=be.dgrad_weights(pd, wd, ad))
                gx = ops.conv2d_dgrad(g1, w1, x.shape, a1, None, x, s1, 0, accumulate_into=gx, prepared=pw1)
        return gx, gw1, gw2, gw3, gwd, None, None, None


def _wgrad_cudnn(x, g, w16, stride, pad, scale=None):
    """Library weight gradient (ATen/cuDNN): kept only as an A/B switch for bench.py --wgrad cudnn."""
    gw = torch.ops.aten.convolution_backward(g, x, w16, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0], 1,
                                             [False, True, False])[1].float()
    return gw * scale[:, None, None, None] if scale is not None else gw


def _wgrad_tc(x, g, w16, stride, pad, scale=None):
    """Weight gradient on the tcgen05 engine (MN-major operands, split-K, fp32 red.add, BN scale fused)."""
    from mrb_b200 import ops
    return ops.conv2d_wgrad(x, g, w16.shape, stride, pad, scale)


class B200Backend(Backend):
    name = "b200"
    act_dtype = torch.bfloat16
    channels_last = True
    stride2_3x3 = False       # 3x3 stride-2 convs (STRIDE_IN_1X1: False) on the engine
    # proposal selection / target assignment / sampling as the launches of csrc/detect_glue.cu (MRB_FUSED_GLUE=0: the
    # PyTorch formulation, kept as the A/B arm and as what the CPU checker backend runs)
    fused_glue = __import__("os").environ.get("MRB_FUSED_GLUE", "1") != "0"
    # the three loss stages as the fused forward / backward launches of csrc/loss_glue.cu (MRB_FUSED_LOSSES=0: PyTorch ops)
    fused_losses = __import__("os").environ.get("MRB_FUSED_LOSSES", "1") != "0"
    # RPN top-k + decode as one cluster launch per level (mrb_rpn_topk_decode) instead of torch.topk + mrb_rpn_decode_packed
    fused_topk = __import__("os").environ.get("MRB_FUSED_TOPK", "1") != "0"

    def __init__(self, wgrad="tc"):
        self._w16 = {}
        self._wd = {}     # (id(param), id(scale)) -> [param, scale, version, w16, prepared dgrad weights]
        self.arena = None
        self.side = None          # second stream for gradient-sink kernels (see side_launch)
        self.lanes = []
        self._side_busy = False
        if wgrad == "tc":
            self.wgrad_fn, self.wgrad_impl = _wgrad_tc, "mrb_conv2d_wgrad (tcgen05, in-house)"
        else:
            self.wgrad_fn, self.wgrad_impl = _wgrad_cudnn, "aten.convolution_backward (cuDNN)"

    def attach_arena(self, arena):
        """Parameters now live in a ParamArena: their bf16 operand copies are the arena's (kept current by its fused
        update kernel) and their gradients accumulate into the arena's persistent fp32 views."""
        self.arena = arena
        self._w16 = {k: v for k, v in self._w16.items() if not isinstance(k, int)}
        self._wd.clear()
        for p in arena.params:
            if p.dim() in (2, 4):
                self._w16[id(p)] = (p, p._version, arena.views16[id(p)])

    def arena_updated(self):
        """The arena's update kernel rewrote parameters and bf16 copies behind autograd's back (no version bump):
        re-derive the flipped/scaled data-gradient weights of every conv seen so far, one batched launch."""
        stale = [e for e in self._wd.values() if self.arena.grad_sink(e[0]) is not None]
        if stale:
            from mrb_b200 import ops
            ops.prepare_dgrad_weights([e[3] for e in stale], [e[1] for e in stale], [e[4] for e in stale])

    def prepare_async(self):
        """Issue arena_updated() on its own stream (start of a step, with ParamArena.defer_dgrad_prepare); join_prepare() before
        the first data-gradient launch (i.e. before backward)."""
        if self.arena is None:
            return
        if getattr(self, "_prep_stream", None) is None:
            self._prep_stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self._prep_stream.wait_stream(cur)
        with torch.cuda.stream(self._prep_stream):
            self.arena_updated()
        self._prep_pending = True

    def join_prepare(self):
        if getattr(self, "_prep_pending", False):
            torch.cuda.current_stream().wait_stream(self._prep_stream)
            self._prep_pending = False

    def enable_overlap(self, on=True):
        """Run the weight-/bias-gradient kernels that accumulate into arena sinks on a second stream: nothing in the
        backward pass reads their result, so they overlap the data-gradient chain (most layers of res4/res5 and the
        heads fill well under 148 SMs).  Captured into the CUDA graph as parallel branches; join_side() is the join."""
        import os
        prio = -1 if os.environ.get("MRB_SIDE_PRIORITY", "0") == "1" else 0        # A/B switch; measured 8.44 vs 8.29 ms/step: off
        self.side = torch.cuda.Stream(priority=prio) if on else None
        self.lanes = [torch.cuda.Stream() for _ in range(6)] if on else []      # independent glue chains (fork(lane=i))

    def side_launch(self, tensors, fn):
        if self.side is None:
            fn()
            return
        self.side.wait_stream(torch.cuda.current_stream())        # inputs were produced on the current stream
        with torch.cuda.stream(self.side):
            fn()
        for t in tensors:
            t.record_stream(self.side)                             # keep the allocator from recycling them early
        self._side_busy = True

    def fork(self, inputs, fn, lane=None):
        """Run fn() (no autograd, independent of everything launched since) on another stream -- the gradient stream, or
        one of the `lanes` for mutually independent chains (e.g. the per-level proposal pipelines); returns a handle
        for join().  Without extra streams it just runs inline."""
        st = self.side if lane is None else (self.lanes[lane % len(self.lanes)] if self.lanes else None)
        if st is None:
            return (None, fn())
        cur = torch.cuda.current_stream()
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            out = fn()
        for t in inputs:
            t.record_stream(st)
        return (st, out)

    def join(self, handle):
        st, out = handle
        if st is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(st)

            def mark(o):
                if isinstance(o, torch.Tensor):
                    o.record_stream(cur)
                elif isinstance(o, (list, tuple)):
                    for x in o:
                        mark(x)
            mark(out)
        return out

    def join_side(self):
        if self.side is not None and self._side_busy:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_busy = False

    def heads_boundary(self, feats):
        if self.arena is None or self.arena.world <= 1:
            return feats
        return list(_HeadsBoundaryFn.apply(self, *feats))

    def stage_boundary(self, xs, bucket):
        """Mark `xs` as the inputs of everything belonging to gradient bucket `bucket` (data-parallel runs only)."""
        if self.arena is None or self.arena.world <= 1 or not any(x.requires_grad for x in xs):
            return xs
        return list(_BucketBoundaryFn.apply(self, bucket, *xs))

    def heads_grads_ready(self):
        if self.arena is not None:
            self.arena.early_reduce()

    def grad_sink(self, p):
        return self.arena.grad_sink(p) if (self.arena is not None and p is not None) else None

    def _weight16(self, w):
        """bf16 KRSC copy of a weight.  Cached per nn.Parameter (refreshed when the optimizer bumps its
        version); derived tensors (views, concatenations) are temporaries whose id/address can be recycled,
        so they are converted on every call."""
        def conv(t):
            t = t.detach().to(torch.bfloat16)
            return t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t.contiguous()
        if not isinstance(w, torch.nn.Parameter):
            return conv(w)
        key = id(w)
        ent = self._w16.get(key)
        if self.arena is not None and self.arena.grad_sink(w) is not None:
            # arena-resident copy: kept current by the fused update; re-cast in place if someone else wrote the parameter
            if ent[1] != w._version:
                ent[2].copy_(w.detach())
                self._w16[key] = (w, w._version, ent[2])
                for e in self._wd.values():
                    if e[0] is w:
                        e[2] = -1
            return ent[2]
        if ent is None or ent[0] is not w or ent[1] != w._version or ent[2].device != w.device:
            self._w16[key] = (w, w._version, conv(w))   # holding `w` keeps its id from being recycled
        return self._w16[key][2]

    def prepare_input(self, images):
        return images.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def conv(self, x, weight, scale=None, shift=None, bias=None, residual=None, stride=1, pad=0, relu=False,
             out_fp32=False, w16=None, premask_x=False, gy_premasked=False, residual_up2=False, wparam=None,
             keep_padded=False):
        """`wparam`: the nn.Parameter `weight` is a plain view of (linear's [Cout, K] -> [Cout, K, 1, 1]), if any.
        keep_padded: return all round-up-to-8 output channels (the extra ones are exactly zero)."""
        wsink = self.grad_sink(weight if isinstance(weight, torch.nn.Parameter) else wparam)
        bsink = self.grad_sink(bias) if isinstance(bias, torch.nn.Parameter) else None
        if x.numel() == 0:
            n, _, h, w = x.shape
            kh, kw = weight.shape[2:]
            ph, pw = (pad, pad) if isinstance(pad, int) else pad
            return x.new_zeros((n, weight.shape[0], (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1),
                               dtype=torch.float32 if out_fp32 else torch.bfloat16)
        co = weight.shape[0]
        if co % 8:
            # the data-gradient GEMM reduces over Cout, whose row pitch must be a multiple of 16 B for TMA:
            # pad the (few, small) odd-sized heads -- RPN 3+12, predictor 81+324, mask logits 81 -- with zero
            # output channels and slice them off again
            if residual is not None:
                raise RuntimeError("conv: residual with Cout % 8 != 0 is not supported")
            padn = 8 - co % 8
            weight = torch.cat([weight, weight.new_zeros((padn,) + tuple(weight.shape[1:]))], 0)
            if bias is not None:
                bias = torch.cat([bias, bias.new_zeros(padn)])
            if shift is not None:
                shift = torch.cat([shift, shift.new_zeros(padn)])
            if scale is not None:
                scale = torch.cat([scale, scale.new_ones(padn)])
            w16 = wsink = bsink = None
        if w16 is None:
            w16 = self._weight16(weight)
        y = _ConvFn.apply(x, weight, bias, residual, w16, scale, shift, stride, pad, relu, out_fp32, self.wgrad_fn,
                          premask_x, gy_premasked, residual_up2, self, wsink, bsink)
        return y[:, :co] if (co % 8 and not keep_padded) else y

    def conv_select(self, x, weight, bias, labels, premask_x=False):
        """1x1 conv with fp32 output followed by the per-sample channel pick y[r, labels[r]] (mask logits + loss.py:120-126)."""
        # bf16 logits: the loss only reads one plane per ROI (converted to fp32 after the pick); an fp32 [R, 88, M, M]
        # tensor would be the largest write of the mask head and its gradient would travel fp32 -> bf16 again
        y = self.conv(x, weight, bias=bias, out_fp32=False, premask_x=premask_x, keep_padded=True)
        return _SelectChannelFn.apply(y, labels).float()

    def bottleneck(self, blk, x, g_premasked):
        """Whole bottleneck as one autograd node (see _BottleneckFn).  Requires STRIDE_IN_1X1 geometry."""
        wd = blk.downsample[0].weight if blk.downsample is not None else None
        return _BottleneckFn.apply(x, blk.conv1.weight, blk.conv2.weight, blk.conv3.weight, wd, self, blk, g_premasked)

    # ---------------------------------------------------------------- grouped / deformable 3x3 variants
    @staticmethod
    def _grouped_ok(c2):
        cin, cout, g = c2.in_channels, c2.out_channels, c2.groups
        return (cin == cout and g > 1 and cin % g == 0 and cin % 64 == 0 and 64 % (cin // g) == 0 and tuple(c2.kernel_size) == (3, 3)
                and tuple(c2.padding) == (1, 1) and tuple(c2.dilation) == (1, 1) and c2.stride[0] == c2.stride[1] and c2.stride[0] in (1, 2)
                and c2.padding_mode == "zeros")

    def grouped_conv(self, x, weight, bias, groups, stride, pad, out_fp32=False):
        """layers.Conv2d(groups > 1) through the engine (ResNeXt geometry only), or None."""
        from mrb_b200 import grouped
        ph, pw = (pad, pad) if isinstance(pad, int) else pad
        co, cg, kh, kw = weight.shape
        cin = x.shape[1]
        if not (cin == co and cin % 64 == 0 and cg * groups == cin and 64 % cg == 0 and (kh, kw, ph, pw) == (3, 3, 1, 1) and stride in (1, 2)):
            return None
        if bias is not None and bias.requires_grad and torch.is_grad_enabled():
            return None
        x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        y = grouped.conv2d_grouped(x16, weight, groups, None, None if bias is None else bias.detach().float(), 1, False, stride)
        return y.float() if out_fp32 else y

    def bottleneck_general_ok(self, mod):
        """Bottleneck whose 3x3 is grouped (ResNeXt, resnet.py:302-311) or deformable (DFConv2d, resnet.py:286-300)."""
        c2 = mod.conv2
        if type(c2).__name__ == "DFConv2d":
            conv, off = getattr(c2, "conv", None), getattr(c2, "offset", None)
            if conv is None or off is None or type(conv).__name__ not in ("DeformConv", "ModulatedDeformConv"):
                return False
            ks = conv.kernel_size if isinstance(conv.kernel_size, tuple) else (conv.kernel_size,) * 2
            one = lambda v: (v if isinstance(v, int) else v[0])     # noqa: E731
            return (tuple(ks) == (3, 3) and conv.groups == 1 and conv.deformable_groups == 1 and one(conv.stride) == 1
                    and one(conv.padding) == 1 and one(conv.dilation) == 1 and getattr(conv, "bias", None) is None
                    and conv.in_channels % 8 == 0 and conv.out_channels % 8 == 0 and tuple(off.kernel_size) == (3, 3)
                    and off.stride[0] == 1 and off.in_channels % 8 == 0)
        return isinstance(c2, torch.nn.Conv2d) and c2.bias is None and self._grouped_ok(c2)

    def bottleneck_general(self, blk, x):
        from mrb_b200 import dcn, grouped
        s1, s3, sd = blk.strides
        (a1, b1), (a2, b2), (a3, b3) = (a.get() for a in blk._aff)
        y = self.conv(x, blk.conv1.weight, a1, b1, stride=s1, relu=True)
        c2 = blk.conv2
        if type(c2).__name__ == "DFConv2d":
            # offsets (and mask logits) stay fp32 NHWC, all round-up-to-8 channels kept (the extra ones are zero)
            om = self.conv(y, c2.offset.weight, bias=c2.offset.bias, pad=1, out_fp32=True, keep_padded=True)
            w = c2.conv.weight
            y = dcn.deform_conv_nhwc(y, om, w, self._weight16(w), a2, b2, relu=True, modulated=bool(c2.with_modulated_dcn),
                                     stride=1, pad=1, be=self, wsink=self.grad_sink(w))
        else:
            y = grouped.conv2d_grouped(y, c2.weight, c2.groups, a2, b2, 1, True, s3, w16=self._weight16(c2.weight),
                                       wsink=self.grad_sink(c2.weight))
        if blk.downsample is not None:
            ad, bd = blk._aff_d.get()
            idn = self.conv(x, blk.downsample[0].weight, ad, bd, stride=sd)
        else:
            idn = x
        return self.conv(y, blk.conv3.weight, a3, b3, residual=idn, relu=True)

    def stem(self, images, weight, scale, shift, relu=True, bias=None, out_fp32=False):
        """7x7/2 conv on 3 channels == 4x4/1 conv on the 2x2 space-to-depth image (12 -> 16 channels):
        out(o) = sum_t w[t] in(2o-3+t); with a leading zero tap t' = t+1 the taps pair up as
        in(2(o-2+a)+i), a = 0..3, i = 0..1 -> s2d block o-2+a, phase i: a 4-tap conv with 2 blocks of
        padding on the left/top, computed only for the Ho x Wo valid outputs."""
        from mrb_b200 import ops
        n, c, h, w = images.shape
        if (h | w) & 1:
            images = F.pad(images, (0, w & 1, 0, h & 1))
        x = F.pixel_unshuffle(images, 2)                                                # [n, 4c, hs, ws]
        hs, ws = x.shape[2:]
        # 12 -> 16 channels; 2 zero columns on the left, 1 on the right: the 4 horizontal taps of output column o are
        # then the memory columns o..o+3 = 64 CONTIGUOUS bf16 of the NHWC row, i.e. one 64-channel "pixel" of an
        # overlapping-window view (pixel pitch 16).  The conv becomes 4x1 over 64 channels: K = 4 full k-blocks
        # instead of 16 quarter-full ones.
        x = F.pad(x, (2, 1, 0, 0, 0, 16 - x.shape[1])).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wp = ws + 3
        xv = torch.as_strided(x, (n, 64, hs, ws), (hs * wp * 16, 1, wp * 16, 16))
        key = ("stem", id(weight))
        ent = self._w16.get(key)
        if ent is None or ent[0] is not weight or ent[1] != weight._version or ent[2].device != weight.device:
            co = weight.shape[0]
            w8 = F.pad(weight.detach(), (1, 0, 1, 0))                                   # [co, 3, 8, 8]
            w4 = w8.view(co, c, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(co, c * 4, 4, 4)
            w4 = F.pad(w4, (0, 0, 0, 0, 0, 16 - c * 4)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            w4 = torch.as_strided(w4, (co, 64, 4, 1), (256, 1, 64, 64))                 # memory [co][kh][kw*16 + ch]
            self._w16[key] = (weight, weight._version, w4)
        else:
            w4 = ent[2]
        add = shift if shift is not None else (bias.detach().float() if bias is not None else None)
        return ops.conv2d_fwd(xv, w4, scale, add, None, 1, (2, 0), relu, torch.float32 if out_fp32 else torch.bfloat16,
                              out_hw=(hs, ws))

    def max_pool(self, x, k, s, p):
        if x.requires_grad and torch.is_grad_enabled():
            return F.max_pool2d(x, k, s, p)          # differentiated pools keep ATen's kernel (argmax + backward)
        from mrb_b200 import ops
        return ops.max_pool_nhwc(x, k, s, p)

    def upsample2x(self, x):
        return F.interpolate(x, scale_factor=2, mode="nearest")

    def lateral_topdown(self, feat, weight, bias, top, premask_x=True):
        # the 2x nearest upsample is folded into the epilogue's residual read: no upsampled map in HBM
        return self.conv(feat, weight, bias=bias, residual=top, premask_x=premask_x, residual_up2=top is not None)

    def dgrad_weights(self, wparam, w16, scale):
        """Flipped/transposed/BN-scaled bf16 weights for the data gradient of a parameter-backed conv, cached per
        parameter version.  refresh_weights() rebuilds all of them in one launch right after the optimizer step, so
        inside a step this is a pure lookup; the first use (or a stale entry) costs one small launch."""
        if wparam is None:
            return None
        from mrb_b200 import ops
        key = (id(wparam), id(scale) if scale is not None else 0)
        ent = self._wd.get(key)
        if ent is None or ent[0] is not wparam or ent[4].device != w16.device:
            ent = [wparam, scale, -1, w16, torch.empty(w16.numel(), dtype=torch.bfloat16, device=w16.device)]
            self._wd[key] = ent
        if ent[2] != wparam._version:
            ent[3] = w16
            ops.prepare_dgrad_weights([w16], [scale], [ent[4]])
            ent[2] = wparam._version
        return ent[4]

    def refresh_weights(self, params):
        """Re-derive the bf16 operand copies of all (changed) parameters with one multi-tensor copy instead of
        one cast kernel per layer; call once per step after the optimizer update."""
        src, dst = [], []
        for w in params:
            if not isinstance(w, torch.nn.Parameter) or w.dim() not in (2, 4):
                continue
            ent = self._w16.get(id(w))
            if ent is None or ent[0] is not w or ent[2].device != w.device or ent[2].stride() != w.stride():
                continue   # first use goes through _weight16
            if ent[1] != w._version:
                src.append(w.detach())
                dst.append(ent[2])
                self._w16[id(w)] = (w, w._version, ent[2])
        if src:
            torch._foreach_copy_(dst, src)
        # ... and the data-gradient copies (flipped / transposed / BN-scaled) of every conv seen so far, in one launch
        stale = [e for e in self._wd.values() if e[2] != e[0]._version]
        if stale:
            from mrb_b200 import ops
            for e in stale:
                e[3] = self._weight16(e[0])
            ops.prepare_dgrad_weights([e[3] for e in stale], [e[1] for e in stale], [e[4] for e in stale])
            for e in stale:
                e[2] = e[0]._version

    def linear(self, x, weight, bias, relu=False, out_fp32=False, premask_x=False, gy_premasked=False, keep_padded=False):
        """Fully connected layer on the conv engine: [R, K] x [Cout, K]^T as a 1x1 conv over R "pixels".
        keep_padded: return the round-up-to-8 output columns (the extra ones are exactly zero)."""
        r, k = x.shape
        co = weight.shape[0]
        w16 = self._weight16(weight).view(co, k, 1, 1)
        y = self.conv(x.to(torch.bfloat16).reshape(r, k, 1, 1), weight.view(co, k, 1, 1), bias=bias, relu=relu,
                      out_fp32=out_fp32, w16=w16, premask_x=premask_x, gy_premasked=gy_premasked,
                      wparam=weight if isinstance(weight, torch.nn.Parameter) else None, keep_padded=keep_padded)
        return y.reshape(r, y.shape[1])

    def deconv2x2(self, x, weight, bias, relu=False, premask_x=False, gy_premasked=False):
        """ConvTranspose2d(k=2, s=2), weight [Cin, Cout, 2, 2]: see _Deconv2x2Fn (two strided 1x1 convs, no shuffle)."""
        if weight.shape[1] % 8 or x.shape[1] % 8:
            raise RuntimeError("deconv2x2: channel counts must be multiples of 8")
        if x.shape[0] == 0:      # no ROIs (e.g. an image without detections at test time)
            return x.new_zeros((0, weight.shape[1], 2 * x.shape[2], 2 * x.shape[3]), dtype=torch.bfloat16)
        return _Deconv2x2Fn.apply(x.to(torch.bfloat16), weight, bias, relu, premask_x, gy_premasked)

    def roi_align_fpn(self, feats, rois, scales, pooled, sampling_ratio, nhwc):
        from mrb_b200 import ops
        return ops.roi_align_fpn(list(feats), rois, scales, pooled, sampling_ratio, out_nhwc=nhwc)

    def nms_batched(self, boxes, scores, sizes, thr, presorted=False):
        from mrb_b200 import ops
        return ops.nms_batched(boxes, scores, sizes, thr, presorted=presorted)
