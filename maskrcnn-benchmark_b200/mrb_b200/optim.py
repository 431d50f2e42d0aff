"""SGD with momentum and weight decay (reference solver/build.py:7-20 semantics: bias parameters get
lr * BIAS_LR_FACTOR and WEIGHT_DECAY_BIAS) expressed as four multi-tensor launches per step.

torch.optim.SGD(foreach=True) silently drops to one kernel per tensor per op as soon as ONE parameter and
its gradient disagree on strides (e.g. KRSC weights, sliced gradients of concatenated heads); here every
tensor is viewed as the flat run of memory it occupies, so the foreach fast path always applies."""
import torch


def _flat(t):
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


class FlatSGD:
    def __init__(self, named_params, lr=0.02, momentum=0.9, weight_decay=1e-4, bias_lr_factor=2.0, weight_decay_bias=0.0):
        self.groups = []
        w, b = [], []
        for name, p in named_params:
            if not p.requires_grad:
                continue
            if not (p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)):
                raise ValueError("FlatSGD: parameter %s is not dense" % name)
            (b if "bias" in name else w).append(p)
        for ps, glr, gwd in ((w, lr, weight_decay), (b, lr * bias_lr_factor, weight_decay_bias)):
            if ps:
                # views of p.detach() (NOT p.data): they share the parameter's version counter, so the in-place update
                # is visible to everything that caches derived copies keyed on p._version (bf16 operand copies)
                self.groups.append({"params": ps, "lr": glr, "wd": gwd, "flat": [_flat(p.detach()) for p in ps],
                                    "buf": [torch.zeros_like(_flat(p.detach())) for p in ps]})
        self.momentum = momentum

    def zero_grad(self):
        for g in self.groups:
            for p in g["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self):
        for g in self.groups:
            grads = []
            for p in g["params"]:
                gr = p.grad
                if gr is None:
                    gr = torch.zeros_like(p)
                if gr.stride() != p.stride() or gr.dtype != p.dtype:
                    gr = torch.empty_like(p).copy_(gr)       # same memory order as the parameter
                grads.append(_flat(gr))
            if g["wd"] != 0:
                torch._foreach_add_(grads, g["flat"], alpha=g["wd"])
            torch._foreach_mul_(g["buf"], self.momentum)
            torch._foreach_add_(g["buf"], grads)
            torch._foreach_add_(g["flat"], g["buf"], alpha=-g["lr"])


class ParamArena:
    """All trainable parameters of a model packed into four flat device buffers -- fp32 parameters, fp32 gradient
    accumulators, fp32 momentum, bf16 operand copies -- so that the per-step bookkeeping around the kernels is
    three launches instead of several hundred:

      * every `p` becomes a view of the parameter buffer and `p.grad` a PERSISTENT view of the gradient buffer with
        the parameter's own memory order (KRSC for conv weights).  The weight/bias-gradient kernels red.add straight
        into these views (B200Backend looks them up through `grad_sink`), so there is no per-layer zero-fill, no
        temporary and no AccumulateGrad add; gradients that still arrive through autograd (padded / concatenated
        heads) are accumulated in place by autograd itself;
      * data-parallel averaging is ONE all-reduce of the gradient buffer (no packing copy), its 1/world folded into
        the update;
      * `step()` is mrb_sgd_momentum_step once per group (weights; biases with lr x BIAS_LR_FACTOR and
        WEIGHT_DECAY_BIAS, reference solver/build.py:12-18): update, bf16 operand refresh and gradient zeroing in a
        single pass.  The whole step stays CUDA-graph capturable.

    zero_grad() is a no-op by construction (the update leaves the accumulators at zero)."""

    ALIGN = 64      # elements: every parameter starts 256 B (fp32) / 128 B (bf16) aligned -> TMA / vector friendly

    def __init__(self, named_params, backend, lr=0.02, momentum=0.9, weight_decay=1e-4, bias_lr_factor=2.0,
                 weight_decay_bias=0.0, world_size=1, group=None, late_prefix="backbone.", early_reduce=None,
                 bucket_prefixes=("backbone.body.layer1.", "backbone.body.layer2.", "backbone.body.layer3.",
                                  "backbone.body.layer4.", "backbone.fpn.")):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        for n, p in named:
            if not (p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)):
                raise ValueError("ParamArena: parameter %s is not dense" % n)
        w = [(n, p) for n, p in named if "bias" not in n]
        b = [(n, p) for n, p in named if "bias" in n]
        self.params = [p for _, p in w + b]
        dev = self.params[0].device

        def up(x):
            return (x + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        n_w = sum(up(p.numel()) for _, p in w)
        n_b = sum(up(p.numel()) for _, p in b)
        total = n_w + n_b
        self.param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(total, dtype=torch.float32, device=dev)
        self.param16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.groups = [g for g in ((0, n_w, lr, weight_decay), (n_w, total, lr * bias_lr_factor, weight_decay_bias))
                       if g[1] > g[0]]
        self.momentum, self.world, self.group, self.backend = momentum, world_size, group, backend
        # data-parallel buckets, in the order their gradients COMPLETE during backward (reverse registration order):
        #   heads (RPN + ROI heads)  -> reduced when the heads-boundary node fires (B200Backend.heads_boundary)
        #   backbone.fpn (+ ALL biases: only the FPN and the heads have any) -> when the FPN's backward has been issued
        #   backbone.body.layer4 / layer3 / ... -> when each stage's backward has been issued (stage_boundary)
        # Only the first trainable stage (its input needs no gradient, so no boundary fires) is left for sync(): the
        # exposed tail is one small bucket instead of the whole backbone (round 1: +0.36 ms at 2 GPUs, +0.53 ms at 8).
        self.n_w = n_w
        self.buckets = {}
        order, seen_other = [], False
        off_b = 0
        for n, p in w:
            if n.startswith(late_prefix):
                if seen_other:
                    # the downstream bucket must hold only parameters downstream of the backbone
                    raise ValueError("ParamArena: backbone weight %s is registered after a non-backbone weight; "
                                     "the early all-reduce bucket would include it before its gradient is complete" % n)
                key = next((bp for bp in bucket_prefixes if n.startswith(bp)), late_prefix)
            else:
                seen_other = True
                key = "heads"
            if key not in self.buckets:
                self.buckets[key] = [off_b, off_b]
                order.append(key)
            elif order[-1] != key:
                raise ValueError("ParamArena: parameters of bucket %s are not contiguous in registration order (%s)" % (key, n))
            off_b += up(p.numel())
            self.buckets[key][1] = off_b
        if n_b:
            self.buckets["bias"] = [n_w, total]
        self.split = self.buckets["heads"][0] if "heads" in self.buckets else n_w      # [0, split) = the backbone
        self._pending = {}        # bucket name -> async work handle of this step
        # True: step() leaves the re-derivation of the data-gradient weights to the caller, who issues it at the START of the
        # next step on a side stream (B200Backend.prepare_async / join_prepare): they are needed only by the backward pass, so
        # the ~150 us of transposes overlap the forward pass instead of sitting at the end of the step
        self.defer_dgrad_prepare = False
        self._comm = None
        import os
        # A/B switch.  Must be identical on every rank (collective order): pass `early_reduce` explicitly from rank-0
        # config in multi-rank programs; the environment variable is only the single-launcher default.
        self.early = (os.environ.get("MRB_EARLY_REDUCE", "1") != "0") if early_reduce is None else bool(early_reduce)
        # per-stage buckets (FPN, layer4, layer3 ...) on top of the heads bucket; MRB_BUCKETED_REDUCE=0 keeps round 1's two-way
        # split (heads early, everything else at sync) for A/B runs.  Same on every rank.
        self.bucketed = os.environ.get("MRB_BUCKETED_REDUCE", "1") != "0"
        self.sinks, self.views16 = {}, {}
        off = 0
        with torch.no_grad():
            for _, p in w + b:
                n = p.numel()
                view = torch.as_strided(self.param, p.shape, p.stride(), off)
                view.copy_(p)
                p.data = view                    # before any optimizer / graph / cache has seen the old storage
                gview = torch.as_strided(self.grad, p.shape, p.stride(), off)
                p.grad = gview
                self.sinks[id(p)] = (p, gview)
                self.views16[id(p)] = torch.as_strided(self.param16, p.shape, p.stride(), off)
                off += up(n)
            self.param16.copy_(self.param)
        if backend is not None:
            backend.attach_arena(self)

    def grad_sink(self, p):
        ent = self.sinks.get(id(p))
        return ent[1] if ent is not None and ent[0] is p else None

    def zero_grad(self):
        pass

    @torch.no_grad()
    def reduce_bucket(self, name):
        """Start the all-reduce of one bucket on the communication stream; called from backward (B200Backend's boundary
        nodes) once every gradient kernel of the bucket has been issued, so that it overlaps the rest of the backward pass.
        "backbone.fpn." also carries the bias region.  No-op for a single rank / unknown or already reduced buckets."""
        if self.world <= 1 or not self.early or name not in self.buckets or name in self._pending:
            return
        if not self.bucketed and name != "heads":
            return
        import torch.distributed as dist
        lo, hi = self.buckets[name]
        if hi <= lo:
            return
        names = [name] + (["bias"] if name == "backbone.fpn." and "bias" in self.buckets and "bias" not in self._pending else [])
        if not self.grad.is_cuda:            # host tensors (gloo, tests): no streams involved
            for nm in names:
                l, h = self.buckets[nm]
                self._pending[nm] = dist.all_reduce(self.grad[l:h], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self._comm.wait_stream(cur)
        if self.backend is not None and self.backend.side is not None:
            self._comm.wait_stream(self.backend.side)          # the weight-gradient kernels run there
        with torch.cuda.stream(self._comm):
            for nm in names:
                l, h = self.buckets[nm]
                self._pending[nm] = dist.all_reduce(self.grad[l:h], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def early_reduce(self):
        """The RPN + ROI-head bucket (kept under its round-1 name)."""
        self.reduce_bucket("heads")

    @torch.no_grad()
    def sync(self):
        """Sum the gradient accumulators over the data-parallel ranks (the mean's 1/world is applied by step()): reduce
        whatever no boundary node has reduced yet, then wait for the buckets in flight."""
        if self.backend is not None:
            self.backend.join_side()           # gradient kernels running on the backend's second stream
        if self.world > 1:
            import torch.distributed as dist
            if not self._pending:
                dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
                return
            # the not-yet-reduced ranges, merged where adjacent (one collective per run)
            todo = sorted(self.buckets[n] for n in self.buckets if n not in self._pending)
            runs = []
            for lo, hi in todo:
                if runs and runs[-1][1] == lo:
                    runs[-1][1] = hi
                else:
                    runs.append([lo, hi])
            for lo, hi in runs:
                if hi > lo:
                    dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            for wk in self._pending.values():
                wk.wait()
            if self._comm is not None:
                torch.cuda.current_stream().wait_stream(self._comm)
            self._pending = {}

    @torch.no_grad()
    def step(self):
        from mrb_b200 import ops
        if self.backend is not None:
            self.backend.join_side()
        for lo, hi, lr, wd in self.groups:
            ops.sgd_momentum_step(self.param[lo:hi], self.grad[lo:hi], self.mom[lo:hi], self.param16[lo:hi], lr,
                                  self.momentum, wd, 1.0 / self.world, True)
        if self.backend is not None and not self.defer_dgrad_prepare:
            self.backend.arena_updated()
