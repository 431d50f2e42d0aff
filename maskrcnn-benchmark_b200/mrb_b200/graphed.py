"""CUDA-graph replay of a static-shape segment of the UNMODIFIED reference module graph (forward and backward).

The eager reference-graph arm is host bound: the backbone + FPN alone are ~70 Python-dispatched launches forward and ~150
backward per step, every one a ctypes call with its argument marshalling, while the device needs ~4 ms for all of them.
`graph_module(mod, sample_args, ...)` captures the module's forward once and the backward of its outputs once (the way
torch.cuda.make_graphed_callables does) and rebinds `mod.forward` to an autograd node that copies the new inputs into the
captured ones and replays.  What differs from torch's helper, and why this is not simply a call to it:
  * the weight / bias gradients of this repository's conv engine do not come back through autograd: their kernels accumulate
    straight into the ParamArena's persistent gradient views on the backend's second stream.  The capture must join that
    stream before it ends, the warm-up passes must not leave garbage in the accumulators (they are cleared after capture), and
    parameters whose gradient never shows up in autograd are fine (allow_unused);
  * shapes other than the captured ones (or eval mode) fall back to the eager forward instead of failing.
Nothing about the numerics changes: a replay launches exactly the kernels the eager pass launches."""
import torch
from torch.utils._pytree import tree_flatten, tree_unflatten


def _sig(tensors):
    return tuple((tuple(t.shape), t.dtype, t.device) for t in tensors)


class GraphedModule:
    def __init__(self, mod, sample_args, backend=None, arena=None, warmup=3, pool=None, share_inputs=False):
        self.mod, self.be, self.arena = mod, backend, arena
        self.eager_forward = mod.forward
        self.training_state = mod.training
        flat_in, self.in_spec = tree_flatten(tuple(sample_args))
        if not all(isinstance(t, torch.Tensor) for t in flat_in):
            raise ValueError("graph_module: sample arguments must be (nested lists/tuples of) tensors")
        # share_inputs: the sample tensors ARE the static inputs (e.g. the static outputs of the segment upstream: the replay
        # then finds its inputs in place and copies nothing)
        self.static_in = [(t.detach() if share_inputs else t.detach().clone()).requires_grad_(t.requires_grad) for t in flat_in]
        self.sig = _sig(self.static_in)
        self.params = [p for p in mod.parameters() if p.requires_grad]
        surface = [t for t in self.static_in if t.requires_grad] + self.params
        self.pool = torch.cuda.graph_pool_handle() if pool is None else pool
        args = tree_unflatten(self.static_in, self.in_spec)

        def grads_of(outs, gouts):
            req = [o for o in outs if o.requires_grad]
            if not req or not surface:
                return None
            g = torch.autograd.grad(req, surface, [go for o, go in zip(outs, gouts) if o.requires_grad], allow_unused=True)
            if self.be is not None:
                self.be.join_side()           # gradient-sink kernels of this backward run on the backend's second stream
            return g

        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for _ in range(warmup):
                outs, _ = tree_flatten(self.eager_forward(*args))
                grads_of(outs, [torch.zeros_like(o) for o in outs])
            torch.cuda.synchronize()
            self.g_fwd, self.g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fwd, pool=self.pool, stream=st):
                out_tree = self.eager_forward(*args)
            self.static_out, self.out_spec = tree_flatten(out_tree)
            self.static_gout = [torch.zeros_like(o) if o.requires_grad else None for o in self.static_out]
            with torch.cuda.graph(self.g_bwd, pool=self.pool, stream=st):
                g = grads_of(self.static_out, [go if go is not None else None for go in self.static_gout])
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        self.has_bwd = g is not None
        gi = iter(g) if g is not None else iter(())
        self.static_gin = [next(gi) if t.requires_grad and g is not None else None for t in self.static_in]
        self.static_gparam = [next(gi) if g is not None else None for _ in self.params]
        if arena is not None:
            arena.grad.zero_()                # the warm-up / capture passes accumulated into the gradient sinks
        for p in self.params:
            if p.grad is not None and (arena is None or arena.grad_sink(p) is None):
                p.grad = None
        seg = self

        class _Replay(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *inputs):
                for s, t in zip(seg.static_in, inputs[:len(seg.static_in)]):
                    if s.data_ptr() != t.data_ptr():
                        s.copy_(t)
                seg.g_fwd.replay()
                return tuple(o.detach() for o in seg.static_out)

            @staticmethod
            @torch.autograd.function.once_differentiable
            def backward(ctx, *grads):
                for s, g_ in zip(seg.static_gout, grads):
                    if s is not None and g_ is not None and s.data_ptr() != g_.data_ptr():
                        s.copy_(g_)
                    elif s is not None and g_ is None:
                        s.zero_()
                if seg.pre_backward_hook is not None:
                    seg.pre_backward_hook()          # e.g. start the all-reduce of everything downstream of this segment
                seg.g_bwd.replay()
                return tuple(None if x is None else x.detach() for x in seg.static_gin + seg.static_gparam)

        self._fn = _Replay
        self.replays = 0
        self.fallbacks = 0
        self.bypass = False         # True: run the eager forward (launch accounting / A-B runs)
        self.pre_backward_hook = None

    def __call__(self, *args):
        flat, _ = tree_flatten(tuple(args))
        if self.bypass or self.mod.training != self.training_state or not torch.is_grad_enabled() or _sig(flat) != self.sig:
            self.fallbacks += 1
            return self.eager_forward(*args)
        self.replays += 1
        out = self._fn.apply(*(tuple(flat) + tuple(self.params)))
        return tree_unflatten(list(out), self.out_spec)


def graph_module(mod, sample_args, backend=None, arena=None, warmup=3, pool=None, share_inputs=False):
    """Capture `mod` (an nn.Module whose forward takes tensors / lists of tensors of FIXED shape) and rebind its forward.
    Returns the GraphedModule (counters `replays` / `fallbacks`; `.bypass = True` runs the eager forward again)."""
    seg = GraphedModule(mod, sample_args, backend, arena, warmup, pool, share_inputs)
    mod.forward = seg.__call__
    mod._mrb_graphed = seg
    return seg
