"""`nms(dets, scores, threshold)` -- reference layers/nms.py:8 (amp.float_function(_C.nms))."""
from maskrcnn_benchmark import _C

from ._amp import float_function

nms = float_function(_C.nms)
