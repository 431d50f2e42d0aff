"""Box arithmetic shared by RPN and ROI heads (legacy "+1" pixel convention throughout)."""
import math

import numpy as np
import torch


def box_area(b):
    """structures/bounding_box.py:214-226 (xyxy, TO_REMOVE = 1)"""
    wh = b[:, 2:] - b[:, :2] + 1
    return wh[:, 0] * wh[:, 1]


def box_iou(a, b):
    """structures/boxlist_ops.py:53-89: [len(a), len(b)] IoU with +1 areas."""
    area_a, area_b = box_area(a), box_area(b)
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


class BoxCoder:
    """modeling/box_coder.py:7-95"""

    def __init__(self, weights, bbox_xform_clip=math.log(1000.0 / 16)):
        self.weights = weights
        self.clip = bbox_xform_clip

    def encode(self, reference_boxes, proposals):
        # same arithmetic as modeling/box_coder.py:22-50, on (x, y) pairs at once (half the launches)
        e_wh = proposals[:, 2:] - proposals[:, :2] + 1
        e_ctr = proposals[:, :2] + 0.5 * e_wh
        g_wh = reference_boxes[:, 2:] - reference_boxes[:, :2] + 1
        g_ctr = reference_boxes[:, :2] + 0.5 * g_wh
        wx, wy, ww, wh = self.weights
        d_ctr = (g_ctr - e_ctr) / e_wh
        d_wh = torch.log(g_wh / e_wh)
        if wx == wy and ww == wh:          # every reference config: (1,1,1,1) for the RPN, (10,10,5,5) for the box head
            return torch.cat((wx * d_ctr, ww * d_wh), 1)
        return torch.stack((wx * d_ctr[:, 0], wy * d_ctr[:, 1], ww * d_wh[:, 0], wh * d_wh[:, 1]), 1)

    def decode(self, rel_codes, boxes):
        boxes = boxes.to(rel_codes.dtype)
        w = boxes[:, 2] - boxes[:, 0] + 1
        h = boxes[:, 3] - boxes[:, 1] + 1
        cx = boxes[:, 0] + 0.5 * w
        cy = boxes[:, 1] + 0.5 * h
        wx, wy, ww, wh = self.weights
        dx, dy = rel_codes[:, 0::4] / wx, rel_codes[:, 1::4] / wy
        dw = torch.clamp(rel_codes[:, 2::4] / ww, max=self.clip)
        dh = torch.clamp(rel_codes[:, 3::4] / wh, max=self.clip)
        pcx, pcy = dx * w[:, None] + cx[:, None], dy * h[:, None] + cy[:, None]
        pw, ph = torch.exp(dw) * w[:, None], torch.exp(dh) * h[:, None]
        out = torch.zeros_like(rel_codes)
        out[:, 0::4] = pcx - 0.5 * pw
        out[:, 1::4] = pcy - 0.5 * ph
        out[:, 2::4] = pcx + 0.5 * pw - 1
        out[:, 3::4] = pcy + 0.5 * ph - 1
        return out


def clip_boxes(boxes, width, height):
    """BoxList.clip_to_image(remove_empty=False), structures/bounding_box.py:198-212"""
    boxes = boxes.clone()
    boxes[..., 0].clamp_(min=0, max=width - 1)
    boxes[..., 1].clamp_(min=0, max=height - 1)
    boxes[..., 2].clamp_(min=0, max=width - 1)
    boxes[..., 3].clamp_(min=0, max=height - 1)
    return boxes


class Matcher:
    """modeling/matcher.py:5-112.  -1 = below low threshold, -2 = between thresholds."""
    BELOW_LOW, BETWEEN = -1, -2

    def __init__(self, high, low, allow_low_quality_matches=False):
        assert low <= high
        self.high, self.low, self.allow_low = high, low, allow_low_quality_matches

    def __call__(self, quality):  # [num_gt, num_pred]
        if quality.numel() == 0:
            raise ValueError("No ground-truth / proposal boxes available for one of the images during training")
        vals, matches = quality.max(dim=0)
        all_matches = matches.clone() if self.allow_low else None
        below = vals < self.low
        between = (vals >= self.low) & (vals < self.high)
        matches = torch.where(below, torch.full_like(matches, self.BELOW_LOW), matches)
        matches = torch.where(between, torch.full_like(matches, self.BETWEEN), matches)
        if self.allow_low:
            best_per_gt = quality.max(dim=1)[0]
            is_best = (quality == best_per_gt[:, None]).any(dim=0)  # predictions that are some gt's best (ties incl.)
            matches = torch.where(is_best, all_matches, matches)
        return matches


def _pick_random_idx(mask, key, cap, limit):
    """min(#mask, limit) uniformly random elements of `mask` (limit <= cap; limit may be a device scalar) as a
    FIXED-SIZE index list: returns (idx [min(cap, n)], ok) -- idx is clamped into range, ok flags the real picks.
    Every element carries an iid uniform key; the answer is the `limit` smallest keys among the masked ones.
    Sync-free and sort-free on the long vector: for large inputs (268k RPN anchors) the candidates are first thinned
    by a key threshold that keeps ~4*cap of them in expectation (all of them if there are fewer), compacted with
    nonzero_static into a fixed 16*cap buffer, and only that short buffer goes through top-k.
    Approximation of randperm[:limit], documented: the thinned set is Binomial(cnt, 4*cap/cnt) ~ Poisson(4*cap); it
    overflows the 16*cap buffer (tail dropped in index order) or falls short of `limit` <= cap with probability
    < exp(-cap) (Chernoff; cap >= 128 here => < 1e-55) -- never observed, not detected at run time."""
    n = mask.numel()
    dev = mask.device
    if n <= 16384:
        k = torch.where(mask, key, key.new_full((), 2.0))
        kk, sel = torch.topk(k, min(cap, n), largest=False, sorted=True)
    else:
        cnt = mask.sum().clamp(min=1).to(key.dtype)
        cand = mask & (key < (4.0 * cap) / cnt)
        c = min(n, 16 * cap)
        idx = torch.nonzero_static(cand, size=c, fill_value=n).squeeze(1)          # padded with n
        k = torch.where(idx < n, key[idx.clamp(max=n - 1)], key.new_full((), 2.0))
        kk, order = torch.topk(k, min(cap, c), largest=False, sorted=True)
        sel = idx[order]
    ok = (kk < 1.5) & (torch.arange(kk.numel(), device=dev) < limit)
    return sel.clamp(max=n - 1), ok


def _mask_of(idx, ok, n):
    out = torch.zeros(n + 1, dtype=torch.bool, device=idx.device)                    # slot n swallows the padding
    out[torch.where(ok, idx, torch.full_like(idx, n))] = ok
    return out[:n]


def sample_pos_neg_idx(labels, batch_size, positive_fraction, generator=None, key=None):
    """modeling/balanced_positive_negative_sampler.py:19-68 for one image, as fixed-size index lists:
    (pos_idx [P], pos_ok [P], neg_idx [B], neg_ok [B]) with P = int(batch_size * positive_fraction), B = batch_size:
    up to P random positives, the rest (up to batch_size in total) random negatives -- the distribution of
    randperm(...)[:num], without host synchronisation.  labels: -1 ignore, 0 negative, >0 positive."""
    n = labels.numel()
    num_pos_cap = min(int(batch_size * positive_fraction), n)
    if key is None:
        key = torch.rand(n, device=labels.device, generator=generator)
    pos_idx, pos_ok = _pick_random_idx(labels >= 1, key, num_pos_cap, num_pos_cap)
    neg_idx, neg_ok = _pick_random_idx(labels == 0, key, min(batch_size, n), batch_size - pos_ok.sum())
    return pos_idx, pos_ok, neg_idx, neg_ok


def sample_pos_neg(labels, batch_size, positive_fraction, generator=None, key=None):
    """The same sample as boolean masks (pos, neg) over the elements of `labels`.  `key`: the iid uniform keys to use
    (n <= 4096 only) instead of drawing them."""
    n = labels.numel()
    if n <= 4096 or key is not None:
        # few candidates (the <= ~2000 proposals of an image): rank every element among the masked ones by pairwise
        # key comparison -- two n x n passes instead of two single-block top-k launches
        cap = min(int(batch_size * positive_fraction), n)
        if key is None:
            key = torch.rand(n, device=labels.device, generator=generator)
        pos, neg = labels >= 1, labels == 0
        # lt[i, j]: element j precedes element i.  Ties (float32 rand has 2^24 values: ~10% of 2000-element calls
        # contain one) are broken by index, so the counts are exact as with the reference's randperm[:num]
        ar = torch.arange(n, device=labels.device)
        lt = (key[None, :] < key[:, None]) | ((key[None, :] == key[:, None]) & (ar[None, :] < ar[:, None]))
        pos_sel = pos & ((lt & pos[None, :]).sum(1) < cap)
        neg_sel = neg & ((lt & neg[None, :]).sum(1) < (batch_size - pos_sel.sum()))
        return pos_sel, neg_sel
    pos_idx, pos_ok, neg_idx, neg_ok = sample_pos_neg_idx(labels, batch_size, positive_fraction, generator)
    return _mask_of(pos_idx, pos_ok, n), _mask_of(neg_idx, neg_ok, n)


# ---------------------------------------------------------------------------------- anchors
def _whctrs(a):
    w, h = a[2] - a[0] + 1, a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def _mk(ws, hs, cx, cy):
    ws, hs = ws[:, None], hs[:, None]
    return np.hstack((cx - 0.5 * (ws - 1), cy - 0.5 * (hs - 1), cx + 0.5 * (ws - 1), cy + 0.5 * (hs - 1)))


def cell_anchors(stride, sizes, aspect_ratios):
    """The Detectron anchor enumeration (modeling/rpn/anchor_generator.py:211-289): ratios on the
    stride-sized base window with rounded widths/heights, then scales = size / stride."""
    base = np.array([1, 1, stride, stride], dtype=np.float64) - 1
    ratios = np.array(aspect_ratios, dtype=np.float64)
    scales = np.array(sizes, dtype=np.float64) / stride
    w, h, cx, cy = _whctrs(base)
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    ratio_anchors = _mk(ws, hs, cx, cy)
    out = []
    for a in ratio_anchors:
        w, h, cx, cy = _whctrs(a)
        out.append(_mk(w * scales, h * scales, cx, cy))
    return torch.from_numpy(np.vstack(out)).float()


def grid_anchors(cell, stride, gh, gw, device):
    """anchor_generator.py:72-98: [(gh*gw*A), 4], location-major, anchor-minor."""
    sx = torch.arange(0, gw * stride, step=stride, dtype=torch.float32, device=device)
    sy = torch.arange(0, gh * stride, step=stride, dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), 1)
    return (shifts.view(-1, 1, 4) + cell.to(device).view(1, -1, 4)).reshape(-1, 4)
