"""apex.amp stand-in: float_function forces fp32, everything else is a no-op."""
import contextlib
import functools

import torch


def float_function(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        cast = lambda x: x.float() if isinstance(x, torch.Tensor) and x.is_floating_point() else x  # noqa: E731
        return fn(*[cast(a) for a in args], **{k: cast(v) for k, v in kwargs.items()})
    return wrapper


def initialize(model, optimizer=None, opt_level="O0", **kw):
    return (model, optimizer) if optimizer is not None else model


@contextlib.contextmanager
def scale_loss(loss, optimizer, **kw):
    yield loss


def init(*a, **kw):
    return None
