"""Drop-in `maskrcnn_benchmark` package root for the Blackwell (sm_100a) hot path.

Holds only what the path needs: `_C` (ctypes binding of libmrb_b200.so) and `layers`
(the reference's `maskrcnn_benchmark.layers` API, re-authored).  Everything else of the reference
(`modeling`, `structures`, `config`, `utils`, `engine`, `data`, `solver`) is NOT rebuilt here:
set MRB_REFERENCE_ROOT=/path/to/maskrcnn-benchmark and those sub-packages resolve, unmodified,
from the reference checkout through the extended package __path__ below, running on top of this
`_C` / `layers`.
"""
import os as _os

_ref_root = _os.environ.get("MRB_REFERENCE_ROOT")
if _ref_root:
    _ref_pkg = _os.path.join(_ref_root, "maskrcnn_benchmark")
    if _os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
        # our directory stays first: `_C` and `layers` always come from here
        __path__.append(_ref_pkg)
