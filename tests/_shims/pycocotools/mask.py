"""Stand-in for the subset of pycocotools.mask the reference's structures/segmentation_mask.py:296-305 uses to
rasterise polygons (frPyObjects -> merge -> decode).  pycocotools is not installed in this image; this is NOT
pycocotools' exact boundary rule (it samples the even-odd rule at pixel centres), which is irrelevant for the
synthetic rectangle masks of the benchmark.  Environment shim only: not part of the product."""
import numpy as np


def frPyObjects(polygons, h, w):
    return [{"size": (int(h), int(w)), "polys": [np.asarray(p, dtype=np.float64).reshape(-1, 2)]} for p in polygons]


def merge(rles, intersect=False):
    assert not intersect
    return {"size": rles[0]["size"], "polys": [p for r in rles for p in r["polys"]]}


def _inside(poly, h, w):
    ys, xs = np.mgrid[0:h, 0:w]
    px, py = xs + 0.5, ys + 0.5
    inside = np.zeros((h, w), dtype=bool)
    x0, y0 = poly[-1]
    for x1, y1 in poly:
        if y0 != y1:
            cond = (y0 > py) != (y1 > py)
            xi = (x1 - x0) * (py - y0) / (y1 - y0) + x0
            inside ^= cond & (px < xi)
        x0, y0 = x1, y1
    return inside


def decode(rle):
    if isinstance(rle, (list, tuple)):
        return np.stack([decode(r) for r in rle], axis=2)
    h, w = rle["size"]
    m = np.zeros((h, w), dtype=bool)
    for p in rle["polys"]:
        m |= _inside(p, h, w)
    return m.astype(np.uint8)


def _unavailable(*a, **kw):
    raise NotImplementedError("pycocotools is not installed in this image")


encode = area = toBbox = iou = _unavailable
