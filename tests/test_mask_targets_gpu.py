"""GPU mask-target rasteriser (csrc/mask_targets.cu: project_masks_on_boxes, reference mask_head/loss.py:11-42, from polygon
vertices) against a numpy statement of its rule (cell centres, even-odd crossing test, union over an instance's polygons) --
the same rule as the polygon stand-in of tests/_shims/pycocotools applied to the reference's crop + resize -- and against
exact rectangles (the harness's closed-form rectangle targets).  pycocotools' own boundary rule is not pinned (absent from
this image); cells whose centre lies within 1e-4 of an edge are excluded from the comparison."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def _np_raster(polys, roi, m):
    x1, y1, x2, y2 = roi
    out = np.zeros((m, m), bool)
    near = np.zeros((m, m), bool)
    px = x1 + (np.arange(m) + 0.5) * (x2 - x1) / m
    py = y1 + (np.arange(m) + 0.5) * (y2 - y1) / m
    PX, PY = np.meshgrid(px, py)
    for p in polys:
        p = np.asarray(p, np.float64).reshape(-1, 2)
        inside = np.zeros((m, m), bool)
        a = p[-1]
        for b in p:
            if a[1] != b[1]:
                cond = (a[1] > PY) != (b[1] > PY)
                xi = (b[0] - a[0]) * (PY - a[1]) / (b[1] - a[1]) + a[0]
                inside ^= cond & (PX < xi)
                near |= cond & (np.abs(PX - xi) < 1e-3)
            near |= (np.abs(PY - a[1]) < 1e-3) | (np.abs(PY - b[1]) < 1e-3)
            a = b
        out |= inside
    return out, near


def test_polygon_rasteriser_matches_rule_and_rectangles(built_lib):
    from mrb_b200 import ops
    from mrb_b200.model.roi_heads import MaskHead
    rng = np.random.RandomState(0)
    instances = []
    for k in range(12):
        polys = []
        for _ in range(1 + k % 3):
            n = rng.randint(3, 9)
            cx, cy, rad = rng.uniform(100, 1200), rng.uniform(100, 700), rng.uniform(20, 200)
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            rr = rad * rng.uniform(0.4, 1.0, n)
            polys.append(np.stack([cx + rr * np.cos(ang), cy + rr * np.sin(ang)], 1).reshape(-1).tolist())
        instances.append(polys)
    ps = ops.PolygonSet(instances, DEV)
    R = 200
    inst = rng.randint(0, 12, R)
    rois = []
    for r in range(R):
        p = np.asarray(instances[inst[r]][0]).reshape(-1, 2)
        x1, y1 = p.min(0) - rng.uniform(-30, 30, 2)
        x2, y2 = p.max(0) + rng.uniform(-30, 30, 2)
        rois.append([x1, y1, max(x2, x1 + 4), max(y2, y1 + 4)])
    rois = np.asarray(rois, np.float32)
    got = ops.mask_targets_polygons(ps, torch.from_numpy(rois).to(DEV), torch.from_numpy(inst).to(DEV), 28).cpu().numpy() > 0.5
    bad = 0
    for r in range(R):
        want, near = _np_raster(instances[inst[r]], rois[r].astype(np.float64), 28)
        bad += int(((got[r] != want) & ~near).sum())
    assert bad == 0
    assert 0.05 < got.mean() < 0.95
    # rectangles: equal to the closed-form targets of the harness (MaskHead.mask_targets)
    g = torch.Generator().manual_seed(1)
    gt = torch.rand(50, 2, generator=g) * 600 + 50
    gt = torch.cat([gt, gt + torch.rand(50, 2, generator=g) * 300 + 20], 1)
    prop = gt + (torch.rand(50, 4, generator=g) - 0.5) * 60
    prop[:, 2:] = torch.maximum(prop[:, 2:], prop[:, :2] + 8)
    rect = [[[float(b[0]), float(b[1]), float(b[2]), float(b[1]), float(b[2]), float(b[3]), float(b[0]), float(b[3])]] for b in gt]
    got = ops.mask_targets_polygons(ops.PolygonSet(rect, DEV), prop.to(DEV), torch.arange(50, device=DEV), 28).cpu()
    want = MaskHead.mask_targets(gt, prop, 28)
    assert float((got != want).float().mean()) < 2e-3          # cell centres that fall exactly on a rectangle edge
