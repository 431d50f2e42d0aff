"""Synthetic-step harness: a from-scratch module graph for the reference's
e2e_{faster,mask}_rcnn_R_50_FPN_1x configs, built on `maskrcnn_benchmark.layers` / the C ABI.

The reference's own modeling/ package stays the reference's (it runs unmodified on top of our
`_C` / `layers` when MRB_REFERENCE_ROOT is set -- tests/test_abi.py).  This harness exists because
the GPU box has no reference checkout: it drives the hot path end to end for bench.py with the
same architecture, state_dict key names, hyper-parameters and loss definitions
(reference modeling/detector/generalized_rcnn.py:16-65 and the files cited per module)."""
from .config import RCNNConfig  # noqa: F401
from .detector import GeneralizedRCNN, build_model  # noqa: F401
