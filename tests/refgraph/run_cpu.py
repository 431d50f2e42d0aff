"""Subprocess body of tests/test_fuse_reference.py: the unmodified reference Mask R-CNN (train mode, all five
losses + every parameter gradient) before and after mrb_b200.fuse.fuse_model(), CPU checker backend."""
import json
import sys

import common

common.activate()
common.route_cpu_C_to_oracle()
reseed = common.deterministic_randperm()
import torch  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
glue_bindings = "--glue-bindings" in sys.argv      # also install the detection-glue bindings (they must fall back on CPU tensors)
cfgname = args[0] if args else "e2e_mask_rcnn_R_50_FPN_1x.yaml"
model, cfg = common.build(cfgname)
model.train()
il, targets = common.inputs()
l0, g0 = common.train_step(model, il, targets, reseed)
model.eval()
with torch.no_grad():
    dets0 = model(il)
model.train()
from oracle.cpu_backend import CpuCheckerBackend  # noqa: E402
from mrb_b200.fuse import fuse_model  # noqa: E402
be = CpuCheckerBackend()
if glue_bindings:
    be.fused_glue = True
rep = fuse_model(model, be)
l1, g1 = common.train_step(model, il, targets, reseed)
worst = 0.0
for k in l0:
    assert abs(l0[k] - l1[k]) <= 1e-4 * max(1.0, abs(l0[k])), (k, l0[k], l1[k])
assert set(g0) == set(g1), sorted(set(g0) ^ set(g1))[:8]
for n in g0:
    d = (g0[n] - g1[n]).abs().max().item()
    s = g0[n].abs().max().item() + 1e-6
    worst = max(worst, d / s)
    assert d / s < 1e-3, (n, d, s)
model.eval()
with torch.no_grad():
    dets = model(il)
# eval mode: the fused graph returns the same detections (boxes, scores, labels[, masks]) as the unfused reference forward
for a, b in zip(dets0, dets):
    assert len(a) == len(b), (len(a), len(b))
    if len(a):
        torch.testing.assert_close(a.bbox, b.bbox, rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(a.get_field("scores"), b.get_field("scores"), rtol=1e-4, atol=1e-5)
        assert torch.equal(a.get_field("labels"), b.get_field("labels"))
        if a.has_field("mask"):
            torch.testing.assert_close(a.get_field("mask"), b.get_field("mask"), rtol=1e-3, atol=1e-4)
print(json.dumps({"report": rep, "losses": l1, "worst_rel_grad": worst, "n_grads": len(g0), "dets": [len(d) for d in dets]}))
