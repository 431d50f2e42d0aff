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
