"""Import-time stand-in for pycocotools.mask (only needed so structures/segmentation_mask.py imports)."""


def _unavailable(*a, **kw):
    raise NotImplementedError("pycocotools is not installed in this image")


frPyObjects = decode = merge = encode = area = toBbox = iou = _unavailable
