#!/usr/bin/env python
"""Launch every detection-glue / loss / NMS / fused-pooler kernel once at the step's shapes (for
`ncu --set full -k regex:'rpn_|roi_assign|nms_|_loss_|roi_align_fpn' ...`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mrb_b200 import ops  # noqa: E402
from mrb_b200.model import box_ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)


def rand_boxes(n, w=1333, h=800, lo=8.0, hi=300.0):
    cx, cy = torch.rand(n, generator=g) * w, torch.rand(n, generator=g) * h
    bw, bh = lo + torch.rand(n, generator=g) * (hi - lo), lo + torch.rand(n, generator=g) * (hi - lo)
    b = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, w - 1)
    b[:, 1::2] = b[:, 1::2].clamp(0, h - 1)
    return b


n, apl, ld = 2, 3, 16
grids = [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
strides = [4, 8, 16, 32, 64]
anchors = [box_ops.grid_anchors(box_ops.cell_anchors(s, (s * 8,), (0.5, 1.0, 2.0)), s, gh, gw, DEV) for (gh, gw), s in zip(grids, strides)]
outs = []
for gh, gw in grids:
    o = torch.randn(n, gh, gw, ld, generator=g)
    o[..., 5 * apl:] = 0
    outs.append(o.to(DEV))
widths = torch.tensor([1333.0, 1333.0], device=DEV)
heights = torch.tensor([800.0, 800.0], device=DEV)
ks = [min(2000, a.shape[0]) for a in anchors]
tot = n * sum(ks)
for rep in range(2):
    boxes = torch.empty((tot, 4), device=DEV)
    scores = torch.empty((tot,), device=DEV)
    off = 0
    for o, a, k in zip(outs, anchors, ks):
        ops.rpn_topk_decode(o, apl, a, k, widths, heights, boxes[off:off + n * k], scores[off:off + n * k])
        off += n * k
    sizes = [k for k in ks for _ in range(n)]
    keep, counts = ops.nms_batched(boxes, scores, sizes, 0.7)
    targets = [{"boxes": rand_boxes(8, lo=40, hi=400).to(DEV), "labels": torch.randint(1, 81, (8,), generator=g).to(DEV)} for _ in range(n)]
    gtp = ops.pad_targets(targets, DEV)
    b, s, v = ops.rpn_collect(boxes, scores, keep, counts, ks, n, 2000, 2000, True, gtp[0], gtp[2])
    keys = torch.rand(b.shape[:2], generator=g).to(DEV)
    sm = ops.roi_assign_sample(b, v, keys, gtp[0], gtp[1], gtp[2], 512, 0.25, 0.5, 0.5, (10.0, 10.0, 5.0, 5.0), 128)
    anchors_all = torch.cat(anchors, 0)
    labels, matched = ops.rpn_anchor_match(anchors_all, gtp[0], gtp[2], widths, heights, 0.7, 0.3, 0.0)
    # losses
    o_box = (torch.randn(1024, 408, generator=g) * 2).to(DEV).requires_grad_(True)
    lc, lb = ops.box_head_loss(o_box, sm["labels"], sm["reg_targets"], 81)
    (lc + lb).backward()
    y = (torch.randn(256, 88, 28, 28, generator=g)).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    tgt = (torch.rand(256, 28, 28, generator=g) < 0.4).float().to(DEV)
    lm = ops.mask_head_loss(y, sm["mask_labels"], tgt, sm["mask_weight"])
    lm.backward()
    # fused pooler, the step's two shapes
    feats = [torch.randn(n, 256, gh, gw, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
             for gh, gw in grids[:4]]
    x = ops.roi_align_fpn(feats, sm["rois"], (0.25, 0.125, 0.0625, 0.03125), 7, 2, out_nhwc=False)
    x.sum().backward()
    xm = ops.roi_align_fpn(feats, sm["mask_rois"], (0.25, 0.125, 0.0625, 0.03125), 14, 2, out_nhwc=True)
    xm.sum().backward()
    torch.cuda.synchronize()
print("done")
