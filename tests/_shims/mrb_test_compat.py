"""Test-only.  torch._six was removed from torch; the reference's utils/imports.py:4 still reads torch._six.PY3."""
import sys
import types

try:
    import torch
    if not hasattr(torch, "_six"):
        six = types.ModuleType("torch._six")
        six.PY3 = True
        six.string_classes = (str,)
        six.int_classes = (int,)
        torch._six = six
        sys.modules["torch._six"] = six
except Exception:  # pragma: no cover
    pass

try:  # the reference still uses the numpy<1.24 aliases (e.g. modeling/rpn/anchor_generator.py)
    import numpy as np
    for _n, _t in (("float", float), ("int", int), ("bool", bool)):
        if _n not in np.__dict__:
            setattr(np, _n, _t)
except Exception:  # pragma: no cover
    pass
