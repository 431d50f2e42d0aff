"""Seeded synthetic inputs shared by the golden generator, the CPU tests, the GPU parity tests and
bench.py.  torch CPU generators are deterministic for a given seed/version."""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def nms_boxes(n, seed, n_clusters=None, img=(1333, 800), distinct_scores=True):
    """Clustered boxes (seed boxes + jitter) so that 30-70 % survive at thr 0.7 / 0.5."""
    g = _gen(1000 + seed)
    k = n_clusters or max(1, n // 20)
    w = torch.exp(torch.rand(k, generator=g) * math.log(512 / 16) + math.log(16))
    ar = torch.tensor([0.5, 1.0, 2.0])[torch.randint(0, 3, (k,), generator=g)]
    bw, bh = w * ar.sqrt(), w / ar.sqrt()
    cx, cy = torch.rand(k, generator=g) * img[0], torch.rand(k, generator=g) * img[1]
    idx = torch.randint(0, k, (n,), generator=g)
    jit = torch.randn(n, 4, generator=g) * 0.2
    x1 = cx[idx] - bw[idx] / 2 + jit[:, 0] * bw[idx]
    y1 = cy[idx] - bh[idx] / 2 + jit[:, 1] * bh[idx]
    x2 = cx[idx] + bw[idx] / 2 + jit[:, 2] * bw[idx]
    y2 = cy[idx] + bh[idx] / 2 + jit[:, 3] * bh[idx]
    boxes = torch.stack([x1.clamp(0, img[0] - 1), y1.clamp(0, img[1] - 1),
                         torch.maximum(x2, x1 + 1).clamp(0, img[0] - 1),
                         torch.maximum(y2, y1 + 1).clamp(0, img[1] - 1)], 1).float().contiguous()
    if distinct_scores:
        scores = (torch.randperm(n, generator=g).float() + 1) / n
    else:
        scores = torch.randint(0, 8, (n,), generator=g).float() / 8
    return boxes, scores.contiguous()


def rois_for_level(r, n_img, seed, img=(1333, 800), min_size=16, max_size=512):
    """[r,5] (batch_idx, x1, y1, x2, y2): log-uniform scale, aspect in {1/2,1,2}, clipped."""
    g = _gen(2000 + seed)
    s = torch.exp(torch.rand(r, generator=g) * math.log(max_size / min_size) + math.log(min_size))
    ar = torch.tensor([0.5, 1.0, 2.0])[torch.randint(0, 3, (r,), generator=g)]
    bw, bh = s * ar.sqrt(), s / ar.sqrt()
    cx, cy = torch.rand(r, generator=g) * img[0], torch.rand(r, generator=g) * img[1]
    x1, y1 = (cx - bw / 2).clamp(0, img[0] - 1), (cy - bh / 2).clamp(0, img[1] - 1)
    x2, y2 = (cx + bw / 2).clamp(0, img[0] - 1), (cy + bh / 2).clamp(0, img[1] - 1)
    b = torch.randint(0, n_img, (r,), generator=g).float()
    return torch.stack([b, x1, y1, x2, y2], 1).float().contiguous()


def roi_align_small():
    g = _gen(7)
    feat = torch.randn(2, 8, 25, 42, generator=g)
    rois = rois_for_level(20, 2, 7, img=(168, 100), min_size=4, max_size=120)
    # adversarial rows: degenerate, out of image, negative coords
    extra = torch.tensor([[0, 5, 5, 5, 5], [1, -40, -30, 20, 10], [0, 150, 90, 400, 300], [1, 10.5, 3.25, 11.0, 90.0]])
    return feat, torch.cat([rois, extra.float()], 0).contiguous()


def roi_align_config1():
    """BASELINE.json configs[0]: 1 image, P2 level 1x256x200x336 fp32, 100 random boxes, scale 0.25."""
    g = _gen(0)
    feat = torch.randn(1, 256, 200, 336, generator=g)
    return feat, rois_for_level(100, 1, 0)


def fpn_features(n_img, seed, channels=256, dtype=torch.float32):
    """P2..P5 maps of an 800x1344 padded batch."""
    g = _gen(3000 + seed)
    return [torch.randn(n_img, channels, h, w, generator=g).to(dtype)
            for (h, w) in ((200, 336), (100, 168), (50, 84), (25, 42))]


def focal_inputs(a, num_classes, seed):
    g = _gen(4000 + seed)
    logits = torch.randn(a, num_classes, generator=g) - 4.6
    t = torch.zeros(a, dtype=torch.int32)
    u = torch.rand(a, generator=g)
    t[u < 0.02] = -1
    pos = u > 0.995
    t[pos] = torch.randint(1, num_classes + 1, (int(pos.sum()),), generator=g).int()
    return logits.contiguous(), t.contiguous()
