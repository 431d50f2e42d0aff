"""RPN loss of the harness (sampled-rows formulation) vs the reference's dense formulation
(modeling/rpn/loss.py:85-131: compute targets for every anchor, then index with the sampled positions),
restated here with plain tensor ops, on the same random sample."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maskrcnn-benchmark_b200"))


def test_sampled_rpn_loss_equals_dense_reference_formula():
    from mrb_b200.model import RCNNConfig, box_ops
    from mrb_b200.model.rpn import RPN
    cfg = RCNNConfig()
    rpn = RPN(cfg, 256)
    g = torch.Generator().manual_seed(5)
    n_img, a = 2, 30000                                          # > 16384: exercises the thinned nonzero_static path
    ctr = torch.rand(a, 2, generator=g) * 600
    wh = torch.rand(a, 2, generator=g) * 200 + 8
    anchors = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    vis = torch.rand(n_img, a, generator=g) > 0.1
    targets = []
    for i in range(n_img):
        c = torch.rand(6, 2, generator=g) * 500 + 50
        s = torch.rand(6, 2, generator=g) * 150 + 20
        targets.append({"boxes": torch.cat([c - s / 2, c + s / 2], 1)})
    logits = [torch.randn(n_img, a, generator=g, requires_grad=True)]
    deltas = [torch.randn(n_img, a, 4, generator=g, requires_grad=True)]
    lo, lb = rpn.loss(anchors, vis, logits, deltas, targets, generator=torch.Generator().manual_seed(9))
    (lo + lb).backward()
    got_grads = (logits[0].grad.clone(), deltas[0].grad.clone())
    logits[0].grad = deltas[0].grad = None

    # dense reference formulation on the same sample (same generator state -> same picks)
    gen = torch.Generator().manual_seed(9)
    labels, reg_t, pos_m, neg_m = [], [], [], []
    for i, t in enumerate(targets):
        midx = rpn.matcher(box_ops.box_iou(t["boxes"], anchors))
        lab = (midx >= 0).float()
        lab[midx == box_ops.Matcher.BELOW_LOW] = 0
        lab[~vis[i]] = -1
        lab[midx == box_ops.Matcher.BETWEEN] = -1
        labels.append(lab)
        reg_t.append(rpn.box_coder.encode(t["boxes"][midx.clamp(min=0)], anchors))
        p, ng = box_ops.sample_pos_neg(lab, cfg.rpn_batch_size, cfg.rpn_positive_fraction, gen)
        pos_m.append(p)
        neg_m.append(ng)
    labels, reg_t, pos_m, neg_m = torch.stack(labels), torch.stack(reg_t), torch.stack(pos_m), torch.stack(neg_m)
    sampled = pos_m | neg_m
    assert int(sampled.sum()) == n_img * cfg.rpn_batch_size and int(pos_m.sum()) > 0
    d = (deltas[0][pos_m] - reg_t[pos_m]).abs()
    beta = 1.0 / 9
    want_b = torch.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta).sum() / sampled.sum()
    want_o = F.binary_cross_entropy_with_logits(logits[0][sampled], labels[sampled])
    assert torch.allclose(lb, want_b, rtol=1e-5, atol=1e-7), (lb, want_b)
    assert torch.allclose(lo, want_o, rtol=1e-5, atol=1e-7), (lo, want_o)
    (want_o + want_b).backward()
    assert torch.allclose(got_grads[0], logits[0].grad, rtol=1e-5, atol=1e-8)
    assert torch.allclose(got_grads[1], deltas[0].grad, rtol=1e-5, atol=1e-8)
