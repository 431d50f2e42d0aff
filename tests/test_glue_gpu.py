"""Detection glue kernels (csrc/detect_glue.cu) on the B200 against
  (a) the harness's PyTorch formulation of the same stage run on the same GPU (bit for bit where the arithmetic order is
      the same; that formulation is pinned to the unmodified reference on CPU by tests/test_harness_*_vs_reference.py), and
  (b) the reference's OWN Python (RPNPostProcessor, Matcher, boxlist_iou, BoxCoder from the unmodified mirror under
      baseline/_ref) run on the same device, where a checkout is available.
Tie order of equal scores is unspecified in the reference (torch.topk); comparisons that could see such ties sort rows."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _cfg():
    from mrb_b200.model.config import RCNNConfig
    return RCNNConfig()


def _rand_boxes(g, n, w, h, min_size=8.0, max_size=300.0):
    cx = torch.rand(n, generator=g) * w
    cy = torch.rand(n, generator=g) * h
    bw = min_size + torch.rand(n, generator=g) * (max_size - min_size)
    bh = min_size + torch.rand(n, generator=g) * (max_size - min_size)
    b = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, w - 1)
    b[:, 1::2] = b[:, 1::2].clamp(0, h - 1)
    return b


def _canon(rows):
    """rows [n, c] -> rows sorted lexicographically (numpy), for order-insensitive comparison"""
    a = rows.detach().cpu().numpy()
    if a.shape[0] == 0:
        return a
    return a[np.lexsort(a.T[::-1])]


# ------------------------------------------------------------------------------------------ 1. decode
@pytest.mark.parametrize("shape", [(2, 200, 336), (2, 25, 42), (1, 13, 21)])
def test_rpn_decode_equals_torch_formulation(built_lib, shape):
    from mrb_b200 import ops
    from mrb_b200.model import box_ops
    n, gh, gw = shape
    g = torch.Generator().manual_seed(gh)
    stride = 800 // gh if gh > 13 else 64
    cell = box_ops.cell_anchors(stride, (stride * 8,), (0.5, 1.0, 2.0))
    anc = box_ops.grid_anchors(cell, stride, gh, gw, DEV)
    a = anc.shape[0]
    lg = torch.randn(n, a, generator=g).to(DEV)
    dl = (torch.randn(n, a, 4, generator=g) * 0.5).to(DEV)
    dl[0, :50, 2:] = 9.0                                   # beyond the log(1000/16) clip
    k = min(2000, a)
    idx = lg.topk(k, dim=1, sorted=True)[1]
    widths = torch.tensor([1333.0, 1201.0][:n], device=DEV)
    heights = torch.tensor([800.0, 777.0][:n], device=DEV)
    boxes = torch.empty((n, k, 4), device=DEV)
    scores = torch.empty((n, k), device=DEV)
    ops.rpn_decode(lg, dl, anc, idx, widths, heights, boxes, scores)
    coder = box_ops.BoxCoder((1.0, 1.0, 1.0, 1.0))
    d = torch.gather(dl, 1, idx[..., None].expand(-1, -1, 4))
    bx = coder.decode(d.reshape(-1, 4), anc[idx.reshape(-1)]).view(n, k, 4)
    lim = torch.stack([widths, heights, widths, heights], 1)[:, None, :] - 1
    want_b = torch.minimum(bx.clamp(min=0), lim)
    want_s = torch.gather(lg.sigmoid(), 1, idx)
    assert torch.equal(boxes, want_b), (boxes - want_b).abs().max().item()
    assert torch.equal(scores, want_s), (scores - want_s).abs().max().item()


# ------------------------------------------------------------------------------------------ 2. collect
def _nms_problem_set(g, n, ks, quantise=None):
    """random, score-sorted (image, level) problems in the layout of nms_batched -> boxes, scores, sizes"""
    bs, ss = [], []
    for k in ks:
        for _ in range(n):
            b = _rand_boxes(g, k, 1333, 800)
            s = torch.rand(k, generator=g)
            if quantise:
                s = (s * quantise).floor() / quantise
            s = s.sort(descending=True)[0]
            bs.append(b)
            ss.append(s)
    return torch.cat(bs).to(DEV), torch.cat(ss).to(DEV), [k for k in ks for _ in range(n)]


def _collect_numpy(boxes, scores, keep, counts, ks, n, post_n, fpn_post_n, per_batch, gt=None, gt_count=None):
    """plain statement of the rule (ties at the cut taken in (image, slot) order)"""
    boxes, scores, keep, counts = [t.cpu().numpy() for t in (boxes, scores, keep, counts)]
    cand = [[] for _ in range(n)]      # (score, row) per image in slot order
    row0 = 0
    for l, k in enumerate(ks):
        for i in range(n):
            base = row0 + i * k
            for j in range(min(int(counts[l * n + i]), post_n, k)):
                r = base + int(keep[base + j])
                cand[i].append((float(scores[r]), r))
        row0 += n * k
    out = []
    if per_batch:
        flat = [(-s, i, c, r) for i in range(n) for c, (s, r) in enumerate(cand[i])]
        flat.sort(key=lambda t: (t[0], t[1], t[2]))
        chosen = set((i, c) for _, i, c, _ in flat[:fpn_post_n])
        for i in range(n):
            out.append([r for c, (s, r) in enumerate(cand[i]) if (i, c) in chosen])
    else:
        for i in range(n):
            order = sorted(range(len(cand[i])), key=lambda c: (-cand[i][c][0], c))[:fpn_post_n]
            out.append([cand[i][c][1] for c in order])
    res = []
    for i in range(n):
        b = boxes[out[i]].reshape(-1, 4)
        s = scores[out[i]].reshape(-1)
        if gt is not None:
            gc = int(gt_count[i])
            b = np.concatenate([b, gt[i, :gc].cpu().numpy()], 0)
            s = np.concatenate([s, np.ones(gc, np.float32)], 0)
        res.append((b, s))
    return res


@pytest.mark.parametrize("per_batch,quantise,with_gt", [(True, None, True), (True, 64, False), (False, None, False),
                                                        (False, 32, True)])
def test_rpn_collect_equals_rule(built_lib, per_batch, quantise, with_gt):
    from mrb_b200 import ops
    g = torch.Generator().manual_seed(7 + int(per_batch))
    n, ks = 2, [2000, 2000, 1200, 600, 300]
    post_n, fpn_post_n = (2000, 2000) if per_batch else (1000, 1000)
    boxes, scores, sizes = _nms_problem_set(g, n, ks, quantise)
    keep, counts = ops.nms_batched(boxes, scores, sizes, 0.7)
    gt = gc = None
    if with_gt:
        gt = torch.zeros((n, 9, 4), device=DEV)
        gt[0, :9] = _rand_boxes(g, 9, 1333, 800).to(DEV)
        gt[1, :4] = _rand_boxes(g, 4, 1333, 800).to(DEV)
        gc = torch.tensor([9, 4], dtype=torch.int32, device=DEV)
    b, s, v = ops.rpn_collect(boxes, scores, keep, counts, ks, n, post_n, fpn_post_n, per_batch, gt, gc)
    want = _collect_numpy(boxes, scores, keep, counts, ks, n, post_n, fpn_post_n, per_batch, gt, gc)
    w = b.shape[1] - (9 if with_gt else 0)
    for i in range(n):
        wb, wsc = want[i]
        vi = v[i].cpu().numpy()
        got_b, got_s = b[i].cpu().numpy()[vi], s[i].cpu().numpy()[vi]
        assert got_b.shape == wb.shape, (got_b.shape, wb.shape)
        np.testing.assert_array_equal(got_b, wb)
        np.testing.assert_array_equal(got_s, wsc)
        assert not b[i].cpu().numpy()[~vi].any()
        # valid rows are a prefix of the first W columns and a prefix of the ground-truth columns
        first = vi[:w]
        assert first[:first.sum()].all()
        if with_gt:
            assert vi[w:].tolist() == [j < int(gc[i]) for j in range(9)]


def test_select_proposals_fused_equals_unfused(built_lib):
    """whole RPN proposal stage of the harness, fused launches vs its PyTorch formulation, train (per batch) and eval"""
    from mrb_b200 import ops
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.model.rpn import RPN
    cfg = _cfg()
    rpn = RPN(cfg, 256).to(DEV)
    g = torch.Generator().manual_seed(3)
    n = 2
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    anchors = rpn.anchor_generator.grid(grids, DEV)
    logits = [torch.randn(n, a.shape[0], generator=g).to(DEV) for a in anchors]
    deltas = [(torch.randn(n, a.shape[0], 4, generator=g) * 0.3).to(DEV) for a in anchors]
    sizes = [(800, 1333), (768, 1216)]
    targets = [{"boxes": _rand_boxes(g, 7, 1333, 800).to(DEV), "labels": torch.randint(1, 81, (7,), generator=g).to(DEV)},
               {"boxes": _rand_boxes(g, 3, 1216, 768).to(DEV), "labels": torch.randint(1, 81, (3,), generator=g).to(DEV)}]
    gtp = ops.pad_targets(targets, DEV)
    be = B200Backend()
    for training in (True, False):
        be.fused_glue = True
        fb, fs, fv = rpn.select_proposals(be, anchors, logits, deltas, sizes, targets if training else None, training,
                                          gtp=gtp if training else None)
        be.fused_glue = False
        ub, us, uv = rpn.select_proposals(be, anchors, logits, deltas, sizes, targets if training else None, training)
        assert fb.shape == ub.shape and fv.shape == uv.shape
        for i in range(n):
            f = torch.cat([fb[i][fv[i]], fs[i][fv[i]][:, None]], 1)
            u = torch.cat([ub[i][uv[i]], us[i][uv[i]][:, None]], 1)
            assert f.shape == u.shape, (training, i, f.shape, u.shape)
            if torch.equal(f, u):
                continue
            # equal sigmoid values of distinct logits may be ordered differently by the two top-k's
            cf, cu = _canon(f), _canon(u)
            same = (cf == cu).all(1).mean()
            assert same > 0.995, (training, i, same)


def test_rpn_collect_matches_reference_postprocessor(built_lib):
    """the reference's own RPNPostProcessor (unmodified mirror, its torch ops on the GPU, NMS through this repo's _C) vs
    top-k + mrb_rpn_decode + mrb_nms_batched + mrb_rpn_collect"""
    from mrb_b200 import ops, refenv
    if refenv.activate() is None:
        pytest.skip("reference checkout absent")
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.rpn.inference import RPNPostProcessor
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.model.rpn import RPN
    cfg = _cfg()
    rpn = RPN(cfg, 256).to(DEV)
    g = torch.Generator().manual_seed(11)
    n, A = 2, 3
    grids = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    anchors = rpn.anchor_generator.grid(grids, DEV)
    obj = [torch.randn(n, A, h, w, generator=g).to(DEV) for h, w in grids]
    reg = [(torch.randn(n, 4 * A, h, w, generator=g) * 0.3).to(DEV) for h, w in grids]
    sizes = [(800, 1333), (768, 1216)]
    targets_ref = [BoxList(_rand_boxes(g, 5, 1333, 800).to(DEV), (1333, 800), mode="xyxy"),
                   BoxList(_rand_boxes(g, 2, 1216, 768).to(DEV), (1216, 768), mode="xyxy")]
    targets = [{"boxes": t.bbox, "labels": torch.ones(len(t), dtype=torch.int64, device=DEV)} for t in targets_ref]
    gtp = ops.pad_targets(targets, DEV)
    logits = [o.permute(0, 2, 3, 1).reshape(n, -1) for o in obj]
    deltas = [r.view(n, A, 4, r.shape[2], r.shape[3]).permute(0, 3, 4, 1, 2).reshape(n, -1, 4) for r in reg]
    be = B200Backend()
    be.fused_glue = True
    for training in (True, False):
        pre, post, fpn_post = (2000, 2000, 2000) if training else (1000, 1000, 1000)
        pp = RPNPostProcessor(pre, post, 0.7, 0, BoxCoder((1.0, 1.0, 1.0, 1.0)), fpn_post_nms_top_n=fpn_post,
                              fpn_post_nms_per_batch=True)
        pp.train(training)
        anchor_lists = [[BoxList(a, (s[1], s[0]), mode="xyxy") for a in anchors] for s in sizes]
        ref = pp(anchor_lists, obj, reg, targets_ref if training else None)
        fb, fs, fv = rpn.select_proposals(be, anchors, logits, deltas, sizes, targets if training else None, training,
                                          gtp=gtp if training else None)
        for i in range(n):
            r = torch.cat([ref[i].bbox, ref[i].get_field("objectness")[:, None]], 1)
            f = torch.cat([fb[i][fv[i]], fs[i][fv[i]][:, None]], 1)
            assert r.shape == f.shape, (training, i, r.shape, f.shape)
            cr, cf = _canon(r), _canon(f)
            same = (np.abs(cr - cf).max(1) <= 1e-4).mean()
            assert same > 0.995, (training, i, same)


# ------------------------------------------------------------------------------------------ 3. assign + sample
def _proposal_case(g, p, gcount, near_gt=0.3, invalid_tail=50):
    gt = _rand_boxes(g, gcount, 1333, 800, 30, 400)
    b = _rand_boxes(g, p, 1333, 800)
    m = int(p * near_gt)
    src = gt[torch.randint(0, gcount, (m,), generator=g)]
    b[:m] = src + torch.randn(m, 4, generator=g) * 6
    b = b[torch.randperm(p, generator=g)]
    valid = torch.ones(p, dtype=torch.bool)
    if invalid_tail:
        valid[-invalid_tail:] = False
        b[-invalid_tail:] = 0
    return b, valid, gt


@pytest.mark.parametrize("p,gcounts,near,mask_m", [(2009, (9, 4), 0.02, 128), (2100, (100, 1), 0.5, 128), (700, (3, 5), 0.1, 0),
                                                    (4000, (20, 30), 0.9, 64)])
def test_roi_assign_sample_equals_torch_formulation(built_lib, p, gcounts, near, mask_m):
    from mrb_b200 import ops
    from mrb_b200.model.roi_heads import BoxHead
    cfg = _cfg()
    head = BoxHead(cfg, 256)
    g = torch.Generator().manual_seed(p)
    n = len(gcounts)
    bs, vs, targets = [], [], []
    for gc in gcounts:
        b, v, gt = _proposal_case(g, p, gc, near)
        bs.append(b)
        vs.append(v)
        targets.append({"boxes": gt.to(DEV), "labels": torch.randint(1, 81, (gc,), generator=g).to(DEV)})
    boxes, valid = torch.stack(bs).to(DEV), torch.stack(vs).to(DEV)
    keys = torch.rand((n, p), generator=g).to(DEV)
    keys[0, 5:40] = keys[0, 4]                         # ties
    gtp = ops.pad_targets(targets, DEV)
    S = cfg.roi_batch_size
    got = ops.roi_assign_sample(boxes, valid, keys, gtp[0], gtp[1], gtp[2], S, cfg.roi_positive_fraction, cfg.roi_fg_iou,
                                cfg.roi_bg_iou, cfg.bbox_reg_weights, mask_m)
    wb, wl, wt, wg = head.subsample((boxes, None, valid), targets, keys=keys)
    assert torch.equal(got["labels"], wl)
    assert torch.equal(got["gt_index"], wg)
    rois = got["rois"].view(n, S, 5)
    assert torch.equal(rois[..., 1:], wb)
    assert torch.equal(rois[..., 0], torch.arange(n, device=DEV, dtype=torch.float32)[:, None].expand(n, S))
    pos = wl > 0
    assert int(pos.sum(1).max()) <= int(S * cfg.roi_positive_fraction)
    torch.testing.assert_close(got["reg_targets"][pos], wt[pos], rtol=2e-5, atol=2e-6)
    fin = torch.isfinite(wt).all(-1)
    torch.testing.assert_close(got["reg_targets"][fin], wt[fin], rtol=2e-5, atol=2e-6)
    if mask_m:
        # detector._mask_loss's PyTorch formulation of the positives-first list
        order = torch.sort((~pos).to(torch.int8), dim=1, stable=True)[1][:, :mask_m]
        assert torch.equal(got["mask_weight"].view(n, mask_m), torch.gather(pos, 1, order).float())
        assert torch.equal(got["mask_labels"].view(n, mask_m), torch.gather(wl, 1, order).clamp(min=0))
        assert torch.equal(got["mask_gt_index"], torch.gather(wg, 1, order))
        assert torch.equal(got["mask_rois"].view(n, mask_m, 5), torch.gather(rois, 1, order[..., None].expand(-1, -1, 5)))


def test_roi_assign_labels_match_reference_matcher(built_lib):
    """labels / matched ground truth of the sampled rows vs the reference's boxlist_iou + Matcher (its torch ops on the GPU)"""
    from mrb_b200 import ops, refenv
    if refenv.activate() is None:
        pytest.skip("reference checkout absent")
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.matcher import Matcher
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.boxlist_ops import boxlist_iou
    cfg = _cfg()
    g = torch.Generator().manual_seed(5)
    p, gc = 1500, 12
    b, v, gt = _proposal_case(g, p, gc, 0.4, 0)
    labels_gt = torch.randint(1, 81, (gc,), generator=g)
    boxes, valid = b[None].to(DEV), v[None].to(DEV)
    targets = [{"boxes": gt.to(DEV), "labels": labels_gt.to(DEV)}]
    gtp = ops.pad_targets(targets, DEV)
    keys = torch.rand((1, p), generator=g).to(DEV)
    S = cfg.roi_batch_size
    got = ops.roi_assign_sample(boxes, valid, keys, gtp[0], gtp[1], gtp[2], S, cfg.roi_positive_fraction, cfg.roi_fg_iou,
                                cfg.roi_bg_iou, cfg.bbox_reg_weights, 0)
    # reference: box_head/loss.py:41-80
    q = boxlist_iou(BoxList(gt.to(DEV), (1333, 800)), BoxList(b.to(DEV), (1333, 800)))
    matched = Matcher(cfg.roi_fg_iou, cfg.roi_bg_iou, allow_low_quality_matches=False)(q)
    lab = labels_gt.to(DEV)[matched.clamp(min=0)]
    lab[matched == Matcher.BELOW_LOW_THRESHOLD] = 0
    lab[matched == Matcher.BETWEEN_THRESHOLDS] = -1
    reg = BoxCoder(cfg.bbox_reg_weights).encode(gt.to(DEV)[matched.clamp(min=0)], b.to(DEV))
    # every sampled row is one of the proposals: find it and compare
    rois = got["rois"][:, 1:]
    ok = got["labels"].view(-1) >= 0
    d = (rois[:, None, :] - b.to(DEV)[None, :, :]).abs().sum(-1)
    src = d.argmin(1)
    assert float(d.min(1)[0].max()) == 0.0
    assert torch.equal(got["labels"].view(-1)[ok], lab[src][ok])
    posr = got["labels"].view(-1) > 0
    assert torch.equal(got["gt_index"].view(-1)[posr], matched[src][posr])
    torch.testing.assert_close(got["reg_targets"].view(-1, 4)[posr], reg[src][posr], rtol=1e-5, atol=1e-6)
    npos = int((lab > 0).sum())
    assert int(posr.sum()) == min(npos, int(S * cfg.roi_positive_fraction))
    assert int(ok.sum()) == min(S, int(posr.sum()) + int((lab == 0).sum()))


# ------------------------------------------------------------------------------------------ 4. RPN anchor labelling
@pytest.mark.parametrize("gcounts", [(9, 4), (1, 60)])
def test_rpn_anchor_match_equals_torch_formulation(built_lib, gcounts):
    from mrb_b200 import ops
    from mrb_b200.model import box_ops
    from mrb_b200.model.rpn import RPN
    cfg = _cfg()
    rpn = RPN(cfg, 256).to(DEV)
    g = torch.Generator().manual_seed(sum(gcounts))
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    anchors_all = torch.cat(rpn.anchor_generator.grid(grids, DEV), 0)
    sizes = [(800, 1333), (768, 1216)]
    targets = []
    for gc, (h, w) in zip(gcounts, sizes):
        gt = _rand_boxes(g, gc, w, h, 20, 500)
        if gc == 60:
            gt[7] = torch.tensor([5000.0, 5000.0, 5100.0, 5100.0])      # overlaps no anchor: matcher.py's low-quality pass then
            #                                                             restores EVERY anchor with IoU 0 to it (reference quirk)
        targets.append({"boxes": gt.to(DEV), "labels": torch.ones(gc, dtype=torch.int64, device=DEV)})
    gtp = ops.pad_targets(targets, DEV)
    widths, heights = rpn._sizes(sizes, torch.device(DEV))
    labels, matched = ops.rpn_anchor_match(anchors_all, gtp[0], gtp[2], widths, heights, cfg.rpn_fg_iou, cfg.rpn_bg_iou,
                                           float(cfg.straddle_thresh))
    for i, t in enumerate(targets):
        q = box_ops.box_iou(t["boxes"], anchors_all)
        midx = rpn.matcher(q)
        lab = (midx >= 0).float()
        lab = torch.where(midx == box_ops.Matcher.BELOW_LOW, torch.zeros_like(lab), lab)
        vis = rpn.anchor_generator.visibility(anchors_all, sizes[i][1], sizes[i][0])
        lab = torch.where(~vis, -torch.ones_like(lab), lab)
        lab = torch.where(midx == box_ops.Matcher.BETWEEN, -torch.ones_like(lab), lab)
        assert torch.equal(labels[i], lab), (labels[i] != lab).sum().item()
        posm = lab > 0
        assert torch.equal(matched[i][posm].long(), midx[posm])
        assert int(posm.sum()) > 0


def test_harness_step_fused_glue_runs_and_matches_unfused_losses(built_lib):
    """one train step of the harness with the fused glue: finite losses; with the sampling keys fixed the assignment
    stages feed identical rows, so RPN / box losses equal the unfused step's to bf16 noise"""
    from mrb_b200.model import build_model
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.model.config import RCNNConfig
    torch.manual_seed(0)
    cfg = RCNNConfig()
    be = B200Backend()
    model = build_model(cfg, be, DEV).train()
    g = torch.Generator().manual_seed(1)
    images = torch.randn(2, 3, 320, 448, generator=g).to(DEV)
    sizes = [(320, 448), (300, 400)]
    targets = [{"boxes": _rand_boxes(g, 6, 448, 320, 20, 200).to(DEV), "labels": torch.randint(1, 81, (6,), generator=g).to(DEV)},
               {"boxes": _rand_boxes(g, 3, 400, 300, 20, 200).to(DEV), "labels": torch.randint(1, 81, (3,), generator=g).to(DEV)}]
    out = {}
    for fused in (True, False):
        be.fused_glue = fused
        gen = torch.Generator(device=DEV).manual_seed(5)
        losses = model(images, sizes, targets, generator=gen)
        out[fused] = {k: float(v) for k, v in losses.items()}
        assert all(np.isfinite(v) for v in out[fused].values()), out[fused]
    # the RPN losses do not depend on the ROI sampling keys' layout (same anchors, same generator draws per image)
    for k in ("loss_objectness", "loss_rpn_box_reg"):
        assert abs(out[True][k] - out[False][k]) <= 2e-2 * max(1.0, abs(out[False][k])), (k, out)


# ------------------------------------------------------------------------------------------ 5. fused losses (csrc/loss_glue.cu)
def test_rpn_loss_fused_equals_torch_formulation(built_lib):
    from mrb_b200 import ops
    from mrb_b200.model import box_ops
    from mrb_b200.model.rpn import RPN
    cfg = _cfg()
    rpn = RPN(cfg, 256).to(DEV)
    g = torch.Generator().manual_seed(21)
    n, A, ld = 2, 3, 16
    grids = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    anchors_all = torch.cat(rpn.anchor_generator.grid(grids, DEV), 0)
    outs = []
    for h, w in grids:
        o = torch.randn(n, h, w, ld, generator=g)
        o[..., 5 * A:] = 0
        outs.append(o.to(DEV).requires_grad_(True))
    sizes = [(800, 1333), (768, 1216)]
    targets = [{"boxes": _rand_boxes(g, 6, 1333, 800, 30, 400).to(DEV), "labels": torch.ones(6, dtype=torch.int64, device=DEV)},
               {"boxes": _rand_boxes(g, 2, 1216, 768, 30, 400).to(DEV), "labels": torch.ones(2, dtype=torch.int64, device=DEV)}]
    gtp = ops.pad_targets(targets, DEV)
    gen = torch.Generator(device=DEV).manual_seed(3)
    prepared = rpn.loss_targets_fused(anchors_all, sizes, targets, gtp, gen)
    assert sum(int(p[1].sum()) for p in prepared) > 0          # some positives
    # PyTorch formulation on the same numbers
    logits = [o[..., :A].reshape(n, -1) for o in outs]
    deltas = [o[..., A:5 * A].reshape(n, -1, 4) for o in outs]
    lo, lb = rpn.loss(anchors_all, None, logits, deltas, targets, None, prepared)
    (lo * 1.5 + lb * 0.7).backward()
    want_g = [o.grad.clone() for o in outs]
    for o in outs:
        o.grad = None
    st = tuple(torch.stack([p[j] for p in prepared]) for j in range(6))
    flo, flb = ops.rpn_loss(outs, A, st[3], st[4], st[5], st[0], st[1], st[2])
    (flo * 1.5 + flb * 0.7).backward()
    torch.testing.assert_close(flo, lo, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(flb, lb, rtol=1e-5, atol=1e-7)
    for o, wg in zip(outs, want_g):
        torch.testing.assert_close(o.grad, wg, rtol=1e-4, atol=1e-8)
        assert not o.grad[..., 5 * A:].any()


def test_box_head_loss_fused_equals_torch_formulation(built_lib):
    from mrb_b200 import ops
    from mrb_b200.model.roi_heads import BoxHead
    cfg = _cfg()
    head = BoxHead(cfg, 256)
    g = torch.Generator().manual_seed(22)
    r, nc, ld = 1024, cfg.num_classes, 408
    o = torch.randn(r, ld, generator=g) * 2
    o[:, 5 * nc:] = 0
    o = o.to(DEV).requires_grad_(True)
    labels = torch.randint(-1, nc, (r,), generator=g)
    labels[torch.rand(r, generator=g) < 0.6] = 0
    labels = labels.to(DEV)
    reg_t = torch.randn(r, 4, generator=g).to(DEV)
    lc, lb = head.loss(o[:, :nc], o[:, nc:5 * nc], labels, reg_t)
    (lc * 0.9 + lb * 1.3).backward()
    want = o.grad.clone()
    o.grad = None
    flc, flb = ops.box_head_loss(o, labels, reg_t, nc)
    (flc * 0.9 + flb * 1.3).backward()
    torch.testing.assert_close(flc, lc, rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(flb, lb, rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(o.grad, want, rtol=1e-4, atol=1e-9)


def test_mask_head_loss_fused_equals_torch_formulation(built_lib):
    from mrb_b200 import ops
    g = torch.Generator().manual_seed(23)
    r, c, m = 256, 88, 28
    y0 = (torch.randn(r, c, m, m, generator=g) * 2).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = y0.clone().requires_grad_(True)
    labels = torch.randint(0, 81, (r,), generator=g).to(DEV)
    tgt = (torch.rand(r, m, m, generator=g) < 0.4).float().to(DEV)
    w = (torch.rand(r, generator=g) < 0.3).float().to(DEV)
    # PyTorch formulation (detector._mask_loss): pick the class plane, per-ROI mean BCE, weighted mean
    sel = torch.gather(y.permute(0, 2, 3, 1), 3, labels.view(r, 1, 1, 1).expand(r, m, m, 1)).squeeze(3).float()
    bce = torch.nn.functional.binary_cross_entropy_with_logits(sel, tgt, reduction="none").mean((1, 2))
    want = torch.where(w > 0, bce, torch.zeros((), device=DEV)).sum() / w.sum().clamp(min=1)
    (want * 2.0).backward()
    want_g = y.grad.float().clone()
    y2 = y0.clone().requires_grad_(True)
    got = ops.mask_head_loss(y2, labels, tgt, w)
    (got * 2.0).backward()
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-7)
    assert y2.grad.dtype == torch.bfloat16 and y2.grad.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(y2.grad.float(), want_g, rtol=1.6e-2, atol=1e-9)      # bf16 rounding of the gradient


def test_harness_step_fused_losses_match_unfused(built_lib):
    """the harness train step with fused glue + fused losses vs the PyTorch loss ops on the same assignment (same keys):
    all five losses, and the gradient reaching the backbone output"""
    from mrb_b200.model import build_model
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.model.config import RCNNConfig
    torch.manual_seed(0)
    cfg = RCNNConfig(mask_rois_per_image=64)
    be = B200Backend()
    model = build_model(cfg, be, DEV).train()
    g = torch.Generator().manual_seed(1)
    images = torch.randn(2, 3, 320, 448, generator=g).to(DEV)
    sizes = [(320, 448), (300, 400)]
    targets = [{"boxes": _rand_boxes(g, 6, 448, 320, 20, 200).to(DEV), "labels": torch.randint(1, 81, (6,), generator=g).to(DEV)},
               {"boxes": _rand_boxes(g, 3, 400, 300, 20, 200).to(DEV), "labels": torch.randint(1, 81, (3,), generator=g).to(DEV)}]
    out, grads = {}, {}
    be.fused_glue = True
    for fused in (True, False):
        be.fused_losses = fused
        model.zero_grad(set_to_none=True)
        gen = torch.Generator(device=DEV).manual_seed(5)
        losses = model(images, sizes, targets, generator=gen)
        sum(losses.values()).backward()
        out[fused] = {k: float(v.detach()) for k, v in losses.items()}
        grads[fused] = {k: p.grad.detach().float().clone() for k, p in model.named_parameters() if p.grad is not None}
        assert all(np.isfinite(v) for v in out[fused].values()), out[fused]
    for k in out[True]:
        assert abs(out[True][k] - out[False][k]) <= 5e-3 * max(1.0, abs(out[False][k])), (k, out)
    assert set(grads[True]) == set(grads[False])
    worst = 0.0
    for k in grads[True]:
        a, b = grads[True][k], grads[False][k]
        den = float(b.norm()) + 1e-12
        worst = max(worst, float((a - b).norm()) / den if den > 1e-9 else 0.0)
    assert worst < 5e-2, worst      # bf16 operands on both sides; the loss gradients themselves agree to 1e-4 (tests above)


# ------------------------------------------------------------------------------------------ 6. the same launches behind the reference's own classes
def test_fused_reference_postprocessor_and_mask_targets(built_lib):
    """mrb_b200.fuse rebinds RPNPostProcessor.forward and project_masks_on_boxes of the UNMODIFIED reference: same BoxLists /
    same mask targets as the reference's Python on the same inputs"""
    from mrb_b200 import fuse, refenv
    if refenv.activate() is None:
        pytest.skip("reference checkout absent")
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.roi_heads.mask_head import loss as mask_loss_mod
    from maskrcnn_benchmark.modeling.rpn.inference import RPNPostProcessor
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.model.rpn import RPN
    cfg = _cfg()
    rpn = RPN(cfg, 256).to(DEV)
    g = torch.Generator().manual_seed(31)
    n, A = 2, 3
    grids = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    anchors = rpn.anchor_generator.grid(grids, DEV)
    obj = [torch.randn(n, A, h, w, generator=g).to(DEV) for h, w in grids]
    reg = [(torch.randn(n, 4 * A, h, w, generator=g) * 0.3).to(DEV) for h, w in grids]
    sizes = [(800, 1333), (768, 1216)]
    targets = [BoxList(_rand_boxes(g, 5, 1333, 800).to(DEV), (1333, 800), mode="xyxy"),
               BoxList(_rand_boxes(g, 2, 1216, 768).to(DEV), (1216, 768), mode="xyxy")]
    anchor_lists = [[BoxList(a, (s[1], s[0]), mode="xyxy") for a in anchors] for s in sizes]
    be = B200Backend()
    for training in (True, False):
        pre, post, fpn_post = (2000, 2000, 2000) if training else (1000, 1000, 1000)
        ref = RPNPostProcessor(pre, post, 0.7, 0, BoxCoder((1.0, 1.0, 1.0, 1.0)), fpn_post_nms_top_n=fpn_post, fpn_post_nms_per_batch=True)
        fus = RPNPostProcessor(pre, post, 0.7, 0, BoxCoder((1.0, 1.0, 1.0, 1.0)), fpn_post_nms_top_n=fpn_post, fpn_post_nms_per_batch=True)
        ref.train(training)
        fus.train(training)
        assert fuse._fuse_rpn_postprocessor(fus, be)
        want = ref(anchor_lists, obj, reg, targets if training else None)
        got = fus(anchor_lists, obj, reg, targets if training else None)
        for i in range(n):
            assert got[i].size == want[i].size and got[i].mode == want[i].mode
            r = torch.cat([want[i].bbox, want[i].get_field("objectness")[:, None]], 1)
            f = torch.cat([got[i].bbox, got[i].get_field("objectness")[:, None]], 1)
            assert r.shape == f.shape, (training, i, r.shape, f.shape)
            same = (np.abs(_canon(r) - _canon(f)).max(1) <= 1e-4).mean()
            assert same > 0.995, (training, i, same)
    # mask targets: polygons (rectangles, triangles, two-part instances) on random proposals
    rep = {"fused": {}, "skipped": []}
    orig = getattr(mask_loss_mod.project_masks_on_boxes, "_mrb_orig", mask_loss_mod.project_masks_on_boxes)
    assert fuse._fuse_mask_targets(be, rep)
    fused_fn = mask_loss_mod.project_masks_on_boxes
    try:
        polys = []
        props = _rand_boxes(g, 24, 1333, 800, 20, 300)
        for j in range(24):
            x1, y1, x2, y2 = [float(v) for v in _rand_boxes(g, 1, 1333, 800, 30, 400)[0]]
            if j % 3 == 0:
                polys.append([[x1, y1, x2, y1, x2, y2, x1, y2]])
            elif j % 3 == 1:
                polys.append([[x1, y1, x2, (y1 + y2) / 2, x1, y2]])
            else:
                polys.append([[x1, y1, (x1 + x2) / 2, y1, (x1 + x2) / 2, y2, x1, y2], [(x1 + x2) / 2 + 3, y1, x2, y1, x2, y2]])
            if j % 2:
                props[j] = torch.tensor([x1 - 7.3, y1 + 2.1, x2 + 4.9, y2 - 3.3])       # overlapping its instance
        props[:, 0::2] = props[:, 0::2].clamp(0, 1332)
        props[:, 1::2] = props[:, 1::2].clamp(0, 799)
        seg = SegmentationMask(polys, (1333, 800), mode="poly")
        pl = BoxList(props.to(DEV), (1333, 800), mode="xyxy")
        want = orig(seg, pl, 28)
        got = fused_fn(seg, pl, 28)
        assert got.shape == want.shape == (24, 28, 28) and got.device == want.device
        assert float((got != want).float().mean()) < 2e-3           # cells whose centre lies on an edge, in fp32 vs fp64
        assert float(got.mean()) > 0.02
    finally:
        mask_loss_mod.project_masks_on_boxes = orig


@pytest.mark.parametrize("grid,k", [((200, 336), 2000), ((50, 84), 2000), ((25, 42), 2000), ((13, 21), 2000), ((100, 168), 1000)])
def test_rpn_topk_decode_equals_topk_then_decode(built_lib, grid, k):
    """mrb_rpn_topk_decode (cluster radix select + sort + decode) == torch.topk(sorted) on the logits + mrb_rpn_decode_packed"""
    from mrb_b200 import ops
    from mrb_b200.model import box_ops
    gh, gw = grid
    n, apl, ld = 2, 3, 16
    g = torch.Generator().manual_seed(gh * 7 + k)
    stride = max(4, 800 // gh)
    anc = box_ops.grid_anchors(box_ops.cell_anchors(stride, (stride * 8,), (0.5, 1.0, 2.0)), stride, gh, gw, DEV)
    o = torch.randn(n, gh, gw, ld, generator=g)
    o[..., 5 * apl:] = 0
    o[0, 0, :7, 0] = 5.5                       # equal logits among the best: taken in anchor order
    o = o.to(DEV)
    a = gh * gw * apl
    kk = min(k, a)
    widths = torch.tensor([1333.0, 1201.0], device=DEV)
    heights = torch.tensor([800.0, 777.0], device=DEV)
    b1, s1 = torch.empty((n, kk, 4), device=DEV), torch.empty((n, kk), device=DEV)
    ops.rpn_topk_decode(o, apl, anc, kk, widths, heights, b1, s1)
    lg = o[..., :apl].reshape(n, -1)
    # reference order: descending logit, ascending anchor index among equal logits (a stable descending sort)
    idx = torch.sort(lg, dim=1, descending=True, stable=True)[1][:, :kk].contiguous()
    b2, s2 = torch.empty((n, kk, 4), device=DEV), torch.empty((n, kk), device=DEV)
    ops.rpn_decode_packed(o, apl, anc, idx, widths, heights, b2, s2)
    assert torch.equal(s1, s2), (s1 - s2).abs().max().item()
    assert torch.equal(b1, b2), (b1 - b2).abs().max().item()


def test_box_postprocess_fused_equals_torch_formulation(built_lib):
    """ops.box_postprocess (softmax + decode + threshold, batched NMS over image x class, top-100; no host sync) vs the harness's
    PyTorch formulation of PostProcessor.forward (box_head/inference.py:45-149)"""
    from mrb_b200 import ops
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.model.roi_heads import BoxHead
    cfg = _cfg()
    head = BoxHead(cfg, 256)
    be = B200Backend()
    g = torch.Generator().manual_seed(41)
    n, p, nc, ld = 2, 1000, cfg.num_classes, 408
    o = torch.zeros(n * p, ld)
    o[:, :nc] = torch.randn(n * p, nc, generator=g) * 2.0
    o[:, 0] += 1.0
    hot = torch.randint(1, nc, (n * p,), generator=g)
    o[torch.arange(n * p), hot] += torch.rand(n * p, generator=g) * 6.0          # a few confident classes per row
    o[:, nc:5 * nc] = torch.randn(n * p, 4 * nc, generator=g) * 0.5
    o = o.to(DEV)
    boxes = torch.stack([_rand_boxes(g, p, 1333, 800, 16, 300), _rand_boxes(g, p, 1216, 768, 16, 300)]).to(DEV)
    valid = torch.ones(n, p, dtype=torch.bool)
    valid[1, -37:] = False
    valid = valid.to(DEV)
    sizes = [(800, 1333), (768, 1216)]
    widths = torch.tensor([1333.0, 1216.0], device=DEV)
    heights = torch.tensor([800.0, 768.0], device=DEV)
    b, s, l, c = ops.box_postprocess(o, nc, boxes, valid, widths, heights, cfg.score_thresh, cfg.bbox_reg_weights, cfg.roi_nms,
                                     cfg.detections_per_img)
    want = head.postprocess(be, o[:, :nc], o[:, nc:5 * nc], (boxes, None, valid), sizes)
    for i in range(n):
        k = int(c[i])
        wi = want[i]
        assert k == min(cfg.detections_per_img, wi["scores"].numel()) or abs(k - wi["scores"].numel()) <= 1, (k, wi["scores"].numel())
        got = torch.cat([b[i, :k], s[i, :k, None], l[i, :k, None].float()], 1)
        ref = torch.cat([wi["boxes"], wi["scores"][:, None], wi["labels"][:, None].float()], 1)
        assert not b[i, k:].any() and not s[i, k:].any()
        cg_, cr = _canon(got), _canon(ref)
        if cg_.shape == cr.shape:
            same = (np.abs(cg_ - cr).max(1) <= 1e-4).mean()
            assert same > 0.97, (i, same)
        # class-major order, scores above the threshold
        assert (l[i, :k][1:] >= l[i, :k][:-1]).all() and float(s[i, :k].min()) > cfg.score_thresh


def test_fused_reference_box_postprocessor(built_lib):
    """mrb_b200.fuse rebinds the reference's box-head PostProcessor.forward: same detections as the reference's Python"""
    from mrb_b200 import fuse, refenv
    if refenv.activate() is None:
        pytest.skip("reference checkout absent")
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.roi_heads.box_head.inference import PostProcessor
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from mrb_b200.model.backend import B200Backend
    g = torch.Generator().manual_seed(43)
    nc = 81
    counts = [1000, 873]
    r = sum(counts)
    logits = torch.randn(r, nc, generator=g) * 2.0
    logits[:, 0] += 1.0
    hot = torch.randint(1, nc, (r,), generator=g)
    logits[torch.arange(r), hot] += torch.rand(r, generator=g) * 6.0
    reg = torch.randn(r, 4 * nc, generator=g) * 0.5
    boxes = [BoxList(_rand_boxes(g, counts[0], 1333, 800, 16, 300).to(DEV), (1333, 800), mode="xyxy"),
             BoxList(_rand_boxes(g, counts[1], 1216, 768, 16, 300).to(DEV), (1216, 768), mode="xyxy")]
    x = (logits.to(DEV), reg.to(DEV))
    ref = PostProcessor(0.05, 0.5, 100, BoxCoder((10.0, 10.0, 5.0, 5.0)))
    fus = PostProcessor(0.05, 0.5, 100, BoxCoder((10.0, 10.0, 5.0, 5.0)))
    assert fuse._fuse_box_postprocessor(fus, B200Backend())
    want = ref(x, boxes)
    got = fus(x, boxes)
    for w, gt in zip(want, got):
        assert gt.size == w.size and abs(len(gt) - len(w)) <= 1, (len(gt), len(w))
        a = torch.cat([w.bbox, w.get_field("scores")[:, None], w.get_field("labels")[:, None].float()], 1)
        b = torch.cat([gt.bbox, gt.get_field("scores")[:, None], gt.get_field("labels")[:, None].float()], 1)
        if a.shape == b.shape:
            same = (np.abs(_canon(a) - _canon(b)).max(1) <= 1e-4).mean()
            assert same > 0.97, same


@pytest.mark.parametrize("gcounts", [(9, 4), (60, 1)])
def test_rpn_sample_equals_torch_formulation(built_lib, gcounts):
    """mrb_rpn_sample (cluster radix select of the n smallest keys + encode) vs the harness's PyTorch sampling on the same keys:
    the same sampled sets, the same regression targets"""
    from mrb_b200 import ops
    from mrb_b200.model.rpn import RPN
    cfg = _cfg()
    rpn = RPN(cfg, 256).to(DEV)
    g = torch.Generator().manual_seed(51 + sum(gcounts))
    grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    anchors_all = torch.cat(rpn.anchor_generator.grid(grids, DEV), 0)
    sizes = [(800, 1333), (768, 1216)]
    targets = []
    for gc, (h, w) in zip(gcounts, sizes):
        targets.append({"boxes": _rand_boxes(g, gc, w, h, 20, 500).to(DEV), "labels": torch.ones(gc, dtype=torch.int64, device=DEV)})
    gtp = ops.pad_targets(targets, DEV)
    n, a = 2, anchors_all.shape[0]
    keys = torch.rand((n, a), generator=g).to(DEV)
    want = rpn.loss_targets_fused(anchors_all, sizes, targets, gtp, None, stacked=False, keys=keys)
    widths, heights = rpn._sizes(sizes, torch.device(DEV))
    labels, matched = ops.rpn_anchor_match(anchors_all, gtp[0], gtp[2], widths, heights, cfg.rpn_fg_iou, cfg.rpn_bg_iou,
                                           float(cfg.straddle_thresh))
    pos_idx, pos_ok, reg_t, sel_idx, sel_lab, sel_w = ops.rpn_sample(labels, matched, keys, anchors_all, gtp[0], cfg.rpn_batch_size,
                                                                       cfg.rpn_positive_fraction)
    P = int(cfg.rpn_batch_size * cfg.rpn_positive_fraction)
    for i in range(n):
        wp_idx, wp_ok, wreg, wsel, wlab, ww = want[i]
        got_pos = pos_idx[i][pos_ok[i]]
        ref_pos = wp_idx[wp_ok]
        assert torch.equal(torch.sort(got_pos)[0], torch.sort(ref_pos)[0]), (i, got_pos.numel(), ref_pos.numel())
        assert (got_pos[1:] > got_pos[:-1]).all()                      # emitted in anchor order
        # regression targets of the same anchors
        order_ref = torch.argsort(ref_pos)
        torch.testing.assert_close(reg_t[i][pos_ok[i]], wreg[wp_ok][order_ref], rtol=1e-6, atol=1e-7)
        got_neg = sel_idx[i, P:][sel_w[i, P:] > 0]
        ref_neg = wsel[P:][ww[P:] > 0]
        assert torch.equal(torch.sort(got_neg)[0], torch.sort(ref_neg)[0]), (i, got_neg.numel(), ref_neg.numel())
        assert got_pos.numel() + got_neg.numel() == min(cfg.rpn_batch_size, int((labels[i] >= 1).sum().clamp(max=P)) + int((labels[i] == 0).sum()))
        assert torch.equal(sel_idx[i, :P][pos_ok[i]], got_pos) and float(sel_lab[i, :P].min()) == 1.0 and float(sel_lab[i, P:].max()) == 0.0
        assert torch.equal(sel_w[i, :P] > 0, pos_ok[i])


def test_mask_targets_rect_equals_torch_formulation(built_lib):
    from mrb_b200 import ops
    from mrb_b200.model.roi_heads import MaskHead
    g = torch.Generator().manual_seed(61)
    r, m = 256, 28
    gt = _rand_boxes(g, r, 1333, 800, 30, 400).to(DEV)
    props = _rand_boxes(g, r, 1333, 800, 16, 300)
    props[::3] = gt.cpu()[::3] + torch.randn(len(props[::3]), 4, generator=g) * 5
    props[5] = torch.tensor([10.0, 10.0, 10.0, 50.0])                    # zero width
    rois = torch.cat([torch.zeros(r, 1), props], 1).to(DEV)
    got = ops.mask_targets_rect(gt, rois, m)
    want = MaskHead.mask_targets(gt, rois[:, 1:], m)
    assert torch.equal(got, want), (got != want).sum().item()
    assert 0.02 < float(got.mean()) < 0.9


def test_nms_batched_presorted_equals_sorting_form(built_lib):
    from mrb_b200 import ops
    g = torch.Generator().manual_seed(71)
    n, ks = 2, [2000, 1200, 819, 64]
    boxes, scores, sizes = _nms_problem_set(g, n, ks, quantise=128)          # sorted per problem, with equal scores
    boxes[:300] = boxes[:300] * 0.2 + 100                                    # heavy overlap in the first problem
    k1, c1 = ops.nms_batched(boxes, scores, sizes, 0.7)
    k2, c2 = ops.nms_batched(boxes, scores, sizes, 0.7, presorted=True)
    assert torch.equal(c1, c2)
    off = 0
    for sz, c in zip(sizes, c1.tolist()):
        assert torch.equal(k1[off:off + c], k2[off:off + c])
        off += sz
    assert int(c1[0]) < 1900


def test_fused_reference_rpn_loss_targets_and_box_subsample(built_lib):
    """mrb_b200.fuse binds the target side of RPNLossComputation.__call__ and FastRCNNLossComputation.subsample of the UNMODIFIED
    reference to the glue launches: the sampling-independent parts equal the reference's Python"""
    from mrb_b200 import fuse, refenv
    if refenv.activate() is None:
        pytest.skip("reference checkout absent")
    from maskrcnn_benchmark.modeling.balanced_positive_negative_sampler import BalancedPositiveNegativeSampler
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.matcher import Matcher
    from maskrcnn_benchmark.modeling.roi_heads.box_head.loss import FastRCNNLossComputation
    from maskrcnn_benchmark.modeling.rpn.loss import RPNLossComputation, generate_rpn_labels
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.model.rpn import RPN
    cfg = _cfg()
    rpn = RPN(cfg, 256).to(DEV)
    be = B200Backend()
    g = torch.Generator().manual_seed(81)
    n, A = 2, 3
    grids = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    anchors = rpn.anchor_generator.grid(grids, DEV)
    sizes = [(800, 1333), (768, 1216)]
    targets = [BoxList(_rand_boxes(g, 3, 1333, 800, 90, 400).to(DEV), (1333, 800), mode="xyxy"),
               BoxList(_rand_boxes(g, 2, 1216, 768, 90, 400).to(DEV), (1216, 768), mode="xyxy")]
    for t in targets:
        t.add_field("labels", torch.randint(1, 81, (len(t),), generator=g).to(DEV))
    anchor_lists = []
    for (h, w) in sizes:
        per = []
        for a in anchors:
            bl = BoxList(a, (w, h), mode="xyxy")
            bl.add_field("visibility", (a[:, 0] >= 0) & (a[:, 1] >= 0) & (a[:, 2] < w) & (a[:, 3] < h))
            per.append(bl)
        anchor_lists.append(per)
    obj = [torch.full((n, A, h, w), 0.3, device=DEV) for h, w in grids]                 # equal logits: the loss depends on the
    reg = [(torch.randn(n, 4 * A, h, w, generator=g) * 0.3).to(DEV) for h, w in grids]  # sample only through its counts

    class Parent:
        pass
    par = Parent()
    par.loss_evaluator = RPNLossComputation(Matcher(0.7, 0.3, allow_low_quality_matches=True), BalancedPositiveNegativeSampler(256, 0.5),
                                            BoxCoder((1.0, 1.0, 1.0, 1.0)), generate_rpn_labels)
    ref_ev = par.loss_evaluator
    assert fuse._fuse_rpn_loss(par, be)
    want = ref_ev(anchor_lists, obj, reg, targets)
    got = par.loss_evaluator(anchor_lists, obj, reg, targets)
    torch.testing.assert_close(got[0], want[0], rtol=1e-5, atol=1e-6)
    from maskrcnn_benchmark.structures.boxlist_ops import cat_boxlist
    lab_dense, _ = ref_ev.prepare_targets([cat_boxlist(per) for per in anchor_lists], targets)
    npos = [int((l >= 1).sum()) for l in lab_dense]
    assert min(npos) > 0
    if max(npos) <= 128:          # every positive is sampled by both: the box loss does not depend on the sample
        torch.testing.assert_close(got[1], want[1], rtol=1e-4, atol=1e-6)
    else:                         # the samples differ: same scale only
        assert 0.3 < float(got[1]) / float(want[1]) < 3.0, (float(got[1]), float(want[1]), npos)
    # box-head subsample
    p = 1500
    props = []
    for i, (h, w) in enumerate(sizes):
        b, _, _ = _proposal_case(g, p - 100 * i, len(targets[i]), 0.4, 0)
        b[:len(targets[i]) * 20] = targets[i].bbox.cpu().repeat_interleave(20, 0) + torch.randn(len(targets[i]) * 20, 4, generator=g) * 4
        bl = BoxList(b.to(DEV), (w, h), mode="xyxy")
        bl.add_field("objectness", torch.rand(len(b), generator=g).to(DEV))
        props.append(bl)
    ev = FastRCNNLossComputation(Matcher(0.5, 0.5, allow_low_quality_matches=False), BalancedPositiveNegativeSampler(512, 0.25),
                                 BoxCoder((10.0, 10.0, 5.0, 5.0)))
    lab_ref, reg_ref = ev.prepare_targets(props, targets)                  # the reference's dense labels / regression targets
    assert fuse._fuse_box_subsample(ev, be)
    out = ev.subsample([bl.copy_with_fields(["objectness"]) for bl in props], targets)
    assert ev._proposals is out
    for i, bl in enumerate(out):
        d = (bl.bbox[:, None, :] - props[i].bbox[None, :, :]).abs().sum(-1)
        src = d.argmin(1)
        assert float(d.min(1)[0].max()) == 0.0 and (src[1:] > src[:-1]).all()           # proposal order, no duplicates
        assert torch.equal(bl.get_field("labels"), lab_ref[i][src])
        assert torch.equal(bl.get_field("objectness"), props[i].get_field("objectness")[src])
        pos = bl.get_field("labels") > 0
        torch.testing.assert_close(bl.get_field("regression_targets")[pos], reg_ref[i][src][pos], rtol=1e-5, atol=1e-6)
        npos_all = int((lab_ref[i] > 0).sum())
        assert int(pos.sum()) == min(npos_all, 128) and len(bl) == min(512, int(pos.sum()) + int((lab_ref[i] == 0).sum()))


def test_fused_reference_mask_prepare_targets(built_lib):
    """MaskRCNNLossComputation.prepare_targets of the unmodified reference, fused (device-side matching + one rasterisation launch
    on cached polygon sets) vs the reference's Python (its BoxList / SegmentationMask indexing + host rasterisation)"""
    from mrb_b200 import fuse, refenv
    if refenv.activate() is None:
        pytest.skip("reference checkout absent")
    from maskrcnn_benchmark.modeling.matcher import Matcher
    from maskrcnn_benchmark.modeling.roi_heads.mask_head.loss import MaskRCNNLossComputation
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    from mrb_b200.model.backend import B200Backend
    g = torch.Generator().manual_seed(91)
    props, targets = [], []
    for (w, h), gc in (((1333, 800), 6), ((1216, 768), 3)):
        gt = _rand_boxes(g, gc, w, h, 60, 400)
        polys = []
        for j, b in enumerate(gt.tolist()):
            x1, y1, x2, y2 = b
            polys.append([[x1, y1, x2, y1, x2, y2, x1, y2]] if j % 2 == 0 else [[x1, y1, x2, (y1 + y2) / 2, x1, y2]])
        t = BoxList(gt.to(DEV), (w, h), mode="xyxy")
        t.add_field("labels", torch.randint(1, 81, (gc,), generator=g).to(DEV))
        t.add_field("masks", SegmentationMask(polys, (w, h), mode="poly"))
        targets.append(t)
        pb = gt.repeat_interleave(7, 0) + torch.randn(gc * 7, 4, generator=g) * 6          # positives around every instance
        pb[:, 0::2] = pb[:, 0::2].clamp(0, w - 1)
        pb[:, 1::2] = pb[:, 1::2].clamp(0, h - 1)
        props.append(BoxList(pb.to(DEV), (w, h), mode="xyxy"))
    ref = MaskRCNNLossComputation(Matcher(0.5, 0.5, allow_low_quality_matches=False), 28)
    fus = MaskRCNNLossComputation(Matcher(0.5, 0.5, allow_low_quality_matches=False), 28)
    assert fuse._fuse_mask_prepare_targets(fus, B200Backend())
    lw, mw = ref.prepare_targets(props, targets)
    lg, mg = fus.prepare_targets(props, targets)
    lg2, mg2 = fus.prepare_targets(props, targets)                       # second call: cached polygon sets
    for a, b, c, d, e in zip(lw, lg, mw, mg, mg2):
        assert torch.equal(a, b)
        assert c.shape == d.shape and float((c != d).float().mean()) < 2e-3 and torch.equal(d, e)
        assert float(d.mean()) > 0.05
