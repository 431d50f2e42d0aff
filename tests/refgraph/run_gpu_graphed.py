"""Subprocess body of tests/test_refgraph_gpu.py::test_graphed_segments (B200): the fused reference graph with its
static-shape segments (backbone + FPN, RPN head) replayed as CUDA graphs (mrb_b200.graphed) against the same graph run
eagerly: identical losses and gradients (the replay launches the same kernels; only atomics' summation order differs), on
the captured batch and on a second batch (inputs are copied into the captured buffers)."""
import json
import sys

import common

common.activate()
reseed = common.deterministic_randperm()
import torch  # noqa: E402

assert torch.cuda.is_available()
from mrb_b200 import ops  # noqa: E402
from mrb_b200.fuse import fuse_model  # noqa: E402
from mrb_b200.graphed import graph_module  # noqa: E402
from mrb_b200.model.backend import B200Backend  # noqa: E402
from mrb_b200.optim import ParamArena  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "e2e_mask_rcnn_R_50_FPN_1x.yaml"
model, cfg = common.build(cfgname)
model = model.to("cuda").train()
be = B200Backend()
rep = fuse_model(model, be)
arena = ParamArena(model.named_parameters(), be, lr=1e-3, momentum=0.9, weight_decay=1e-4)
be.enable_overlap(True)
il, tg = common.inputs("cuda")
il2, tg2 = common.inputs("cuda")
il2.tensors.mul_(0.5).add_(3.0)
names = {id(p): n for n, p in model.named_parameters()}


def step(il_, tg_):
    reseed()
    torch.manual_seed(7)
    losses = model(il_, tg_)
    sum(losses.values()).backward()
    be.join_side()
    torch.cuda.synchronize()
    grads = {names[id(p)]: arena.grad_sink(p).detach().float().cpu().clone() for p in arena.params}
    arena.grad.zero_()
    return {k: float(v.detach()) for k, v in losses.items()}, grads


eager = [step(il, tg), step(il2, tg2)]
seg = graph_module(model.backbone, (il.tensors,), backend=be, arena=arena)
feats = model.backbone(il.tensors)
seg2 = graph_module(model.rpn.head, (list(feats),), backend=be, arena=arena, share_inputs=True)
assert float(arena.grad.abs().max()) == 0.0          # the capture passes left nothing in the accumulators
graphed = [step(il, tg), step(il2, tg2), step(il, tg)]


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


out = {"replays": [seg.replays, seg2.replays], "fallbacks": [seg.fallbacks, seg2.fallbacks], "losses_eager": [e[0] for e in eager],
       "losses_graphed": [g[0] for g in graphed], "n_params": len(eager[0][1])}
worst = []
for (le, ge), (lg, gg) in zip(eager + [eager[0]], graphed):
    errs = sorted(((rel(gg[n], ge[n]), n) for n in ge if float(ge[n].norm()) > 0), reverse=True)
    worst.append(errs[:3])
out["worst_grad_rel_err"] = worst
# a full optimizer step through the graphs: parameters move, the next replay sees the new weights
p0 = arena.param.clone()
losses = model(il, tg)
sum(losses.values()).backward()
arena.sync()
arena.step()
torch.cuda.synchronize()
out["param_moved"] = float((arena.param - p0).abs().max())


def fwd_losses():
    reseed()
    torch.manual_seed(7)
    with torch.enable_grad():
        return {k: float(v.detach()) for k, v in model(il, tg).items()}


l_after = fwd_losses()
seg.bypass = seg2.bypass = True
l_after_eager = fwd_losses()
out["losses_after_step"] = {"graphed": l_after, "eager": l_after_eager}
print(json.dumps(out))
