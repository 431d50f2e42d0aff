"""Subprocess body of tests/test_refgraph_gpu.py (B200): the UNMODIFIED reference GeneralizedRCNN over this
repository's `layers` / `_C`, train mode, three ways on identical weights / inputs / sampled indices:
  A  unfused on cuda  -- every layers.Conv2d forward goes to the tcgen05 engine one conv at a time
  B  fused on cuda    -- after mrb_b200.fuse.fuse_model()
  C  fp32 checker     -- the same reference graph on CPU (ATen convs, oracle ROIAlign/NMS)
Compares FPN features, losses and parameter gradients (bf16 operands vs the fp32 checker)."""
import copy
import json
import sys

import common

common.activate()
common.route_cpu_C_to_oracle()
reseed = common.deterministic_randperm()
import torch  # noqa: E402

assert torch.cuda.is_available()
cfgname = sys.argv[1] if len(sys.argv) > 1 else "e2e_mask_rcnn_R_50_FPN_1x.yaml"
# deformable convs have no CPU implementation (the reference raises NotImplementedError, deform_conv_func.py:43-44): for the
# dcn configs the checker is the UNFUSED graph on the GPU, whose DCN layers run the fp32 `_C.deform_conv_*` kernels (1e-4 vs
# the oracle, tests/test_parity_gpu.py)
gpu_checker = "dcn" in cfgname
model_cpu, cfg = common.build(cfgname)
model_cpu.train()
model = copy.deepcopy(model_cpu).to("cuda").train()
il_c, tg_c = common.inputs()
il_g, tg_g = common.inputs("cuda")


def run(m, il, tg):
    feats = {}
    h = m.backbone.register_forward_hook(lambda mod, i, o: feats.__setitem__("f", [t.detach().float().cpu() for t in o]))
    losses, grads = common.train_step(m, il, tg, reseed)
    h.remove()
    return feats["f"], losses, grads


from mrb_b200 import engine, ops  # noqa: E402
ops.STATS["launches"] = 0
fA, lA, gA = run(model, il_g, tg_g)
fC, lC, gC = (fA, lA, gA) if gpu_checker else run(model_cpu, il_c, tg_c)
engine_calls, aten_calls, launches_A = engine.STATS["engine"], engine.STATS["aten"], ops.STATS["launches"]
from mrb_b200.fuse import fuse_model  # noqa: E402
rep = fuse_model(model, sampling=False)      # the checker pins the reference's randperm stream (see fuse_model)
ops.STATS["launches"] = 0
fB, lB, gB = run(model, il_g, tg_g)
launches_B = ops.STATS["launches"]


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


out = {"config": cfgname, "checker": "unfused graph on the GPU (fp32 DCN kernels)" if gpu_checker else "fp32 CPU (ATen + oracle)",
       "report": rep, "engine_calls_unfused": engine_calls, "aten_fallbacks_unfused": aten_calls,
       "libmrb_launches": {"unfused": launches_A, "fused": launches_B},
       "losses": {"checker": lC, "unfused": lA, "fused": lB},
       "feat_rel_err": {"unfused": [rel(a, c) for a, c in zip(fA, fC)], "fused": [rel(b, c) for b, c in zip(fB, fC)]}}
gerr = {}
for tag, g in (("unfused", gA), ("fused", gB)):
    assert set(g) == set(gC), sorted(set(g) ^ set(gC))[:8]
    errs = sorted(((rel(g[n], gC[n]), n) for n in gC), reverse=True)
    gerr[tag] = {"worst": errs[:3], "median": errs[len(errs) // 2][0]}
out["grad_rel_err"] = gerr
model.eval()
with torch.no_grad():
    dets = model(il_g)
out["dets"] = [len(d) for d in dets]
print(json.dumps(out))
