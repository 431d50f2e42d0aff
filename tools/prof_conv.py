#!/usr/bin/env python
"""Launch a few representative tcgen05 conv shapes (for `ncu --set full -k regex:conv_`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mrb_b200 import ops  # noqa: E402

DEV = "cuda:0"
SHAPES = [  # kind, n, cin, h, w, cout, k, stride, pad
    ("fwd", 2, 64, 200, 336, 256, 1, 1, 0),
    ("fwd", 2, 256, 200, 336, 256, 3, 1, 1),
    ("fwd", 2, 256, 50, 84, 1024, 1, 1, 0),
    ("wgrad", 2, 256, 50, 84, 256, 3, 1, 1),
]
which = [int(a) for a in sys.argv[1:]] or range(len(SHAPES))
for i in which:
    kind, n, cin, h, w, cout, k, stride, pad = SHAPES[i]
    x = torch.randn(n, cin, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, k, k, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    for _ in range(2):
        if kind == "fwd":
            y = ops.conv2d_fwd(x, wt, None, None, None, stride, pad, True)
        else:
            go = torch.randn(n, cout, ho, wo, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            y = ops.conv2d_wgrad(x, go, wt.shape, stride, pad)
    torch.cuda.synchronize()
print("done")
