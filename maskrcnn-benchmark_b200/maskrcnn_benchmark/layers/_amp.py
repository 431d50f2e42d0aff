"""fp32-forcing decorator used where the reference used apex.amp.float_function
(layers/nms.py:8, layers/roi_align.py:57, layers/roi_pool.py:56): ROIAlign / ROIPool / NMS always
compute in fp32, whatever autocast mode the caller is in."""
import functools

import torch


def _to_f32(x):
    if isinstance(x, torch.Tensor) and x.is_floating_point() and x.dtype != torch.float32:
        return x.float()
    return x


def float_function(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        args = tuple(_to_f32(a) for a in args)
        kwargs = {k: _to_f32(v) for k, v in kwargs.items()}
        with torch.autocast(device_type="cuda", enabled=False):
            return fn(*args, **kwargs)
    return wrapper
