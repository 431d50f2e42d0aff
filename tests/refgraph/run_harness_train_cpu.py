"""Subprocess body of tests/test_harness_train_vs_reference.py: the harness model (mrb_b200.model, CPU checker backend)
against the UNMODIFIED reference GeneralizedRCNN in TRAIN mode -- all five losses and every parameter gradient -- on the
same weights (strict state_dict load), images, targets, with both samplers made deterministic ("first k candidates in
index order": the reference's randperm(n)[:k] with randperm := arange, the harness's smallest-k random keys with
rand := increasing keys).  Pins the harness's train-mode host logic (anchor labelling, sampling budgets, box/mask target
construction, loss normalisations, the gather-positives mask path) to the reference's."""
import json
import sys

import common

common.activate()
common.route_cpu_C_to_oracle()
import torch  # noqa: E402

orig_randperm, orig_rand = torch.randperm, torch.rand
torch.randperm = lambda n, *a, device=None, **kw: torch.arange(n, device=device if device is not None else "cpu")


def _rand(*size, device=None, generator=None, **kw):
    n = size[0] if len(size) == 1 and isinstance(size[0], int) else None
    if n is None:
        return orig_rand(*size, device=device, generator=generator, **kw)
    return (torch.arange(n, device=device, dtype=torch.float32) + 0.5) / max(n, 1)


model, cfg = common.build("e2e_mask_rcnn_R_50_FPN_1x.yaml")
model.train()
il, targets = common.inputs()
torch.manual_seed(7)
losses_ref = model(il, targets)
sum(losses_ref.values()).backward()
l_ref = {k: float(v.detach()) for k, v in losses_ref.items()}
g_ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

from oracle.cpu_backend import CpuCheckerBackend  # noqa: E402
from mrb_b200.model import GeneralizedRCNN, RCNNConfig  # noqa: E402
hcfg = RCNNConfig(stem_out=8, width_per_group=8, res2_out=32, fpn_out=32, mlp_head_dim=64, mask_conv_layers=(16, 16, 16, 16),
                  pre_nms_top_n_train=200, post_nms_top_n_train=200, fpn_post_nms_top_n_train=200,
                  pre_nms_top_n_test=100, post_nms_top_n_test=100, fpn_post_nms_top_n_test=100, roi_batch_size=64,
                  mask_rois_per_image=0)
h = GeneralizedRCNN(hcfg, CpuCheckerBackend()).train()
missing, unexpected = h.load_state_dict(model.state_dict(), strict=True)
assert not missing and not unexpected
torch.rand = _rand
ht = [{"boxes": t.bbox.clone(), "labels": t.get_field("labels").clone()} for t in targets]
losses_h = h(il.tensors, [tuple(s) for s in il.image_sizes], ht)
sum(losses_h.values()).backward()
torch.rand = orig_rand
l_h = {k: float(v.detach()) for k, v in losses_h.items()}
worst = (0.0, "")
for k in l_ref:
    assert abs(l_ref[k] - l_h[k]) <= 2e-4 * max(1.0, abs(l_ref[k])), (k, l_ref[k], l_h[k])
n_cmp = 0
for n, p in h.named_parameters():
    if n not in g_ref:
        assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
        continue
    d = float((p.grad - g_ref[n]).abs().max())
    s = float(g_ref[n].abs().max()) + 1e-6
    n_cmp += 1
    if d / s > worst[0]:
        worst = (d / s, n)
    assert d / s < 2e-3, (n, d, s)
print(json.dumps({"losses_reference": l_ref, "losses_harness": l_h, "worst_rel_grad": worst, "n_grads": n_cmp}))
