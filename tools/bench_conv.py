#!/usr/bin/env python
"""Micro-benchmark of single tcgen05 conv launches: cold (L2 flushed, one launch between events) and warm (the same
launch replayed back-to-back inside one CUDA graph, average) -- separates launch/prologue cost from steady state.
Usage: python tools/bench_conv.py  [kind n cin h w cout k stride pad]..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mrb_b200 import ops  # noqa: E402

DEV = "cuda:0"
SHAPES = [
    ("fwd", 2, 1024, 50, 84, 256, 1, 1, 0), ("fwd", 2, 256, 50, 84, 1024, 1, 1, 0), ("fwd", 2, 256, 50, 84, 256, 3, 1, 1),
    ("dgrad", 2, 256, 50, 84, 256, 3, 1, 1), ("wgrad", 2, 256, 50, 84, 256, 3, 1, 1), ("wgrad", 2, 1024, 50, 84, 256, 1, 1, 0),
    ("fwd", 2, 512, 25, 42, 512, 3, 1, 1), ("fwd", 2, 64, 8, 16, 64, 1, 1, 0), ("fwd", 256, 256, 14, 14, 256, 3, 1, 1),
    ("wgrad", 256, 256, 14, 14, 256, 3, 1, 1), ("fwd", 1024, 12544, 1, 1, 1024, 1, 1, 0), ("fwd", 2, 256, 200, 336, 256, 3, 1, 1),
]
if len(sys.argv) > 1:
    a = sys.argv[1:]
    SHAPES = [(a[i], *map(int, a[i + 1:i + 9])) for i in range(0, len(a), 9)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for kind, n, cin, h, w, cout, k, stride, pad in SHAPES:
    x = torch.randn(n, cin, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, k, k, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    go = torch.randn(n, cout, ho, wo, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    if kind == "fwd":
        odt = torch.float32 if os.environ.get("MRB_BENCH_FP32") == "1" else torch.bfloat16
        fn = lambda: ops.conv2d_fwd(x, wt, None, None, None, stride, pad, True, out_dtype=odt)  # noqa: E731
    elif kind == "dgrad":
        prep = ops.prepare_dgrad_weights([wt], [None])[0]
        fn = lambda: ops.conv2d_dgrad(go, wt, x.shape, None, None, None, stride, pad, prepared=prep)  # noqa: E731
    else:
        acc = torch.zeros(cout, cin, k, k, device=DEV).contiguous(memory_format=torch.channels_last)
        fn = lambda: ops.conv2d_wgrad(x, go, wt.shape, stride, pad, accumulate_into=acc)  # noqa: E731
    fn()
    cold = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        cold.append(e0.elapsed_time(e1) * 1e3)
    reps = 20
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    warm = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        warm.append(e0.elapsed_time(e1) * 1e3 / reps)
    fl = 2.0 * n * ho * wo * cout * cin * k * k
    print("%-5s n%d cin%d %dx%d cout%d k%d s%d: cold %.1f us (%.0f TF/s)   warm back-to-back %.1f us (%.0f TF/s)" % (
        kind, n, cin, h, w, cout, k, stride, sorted(cold)[2], fl / sorted(cold)[2] / 1e6, min(warm), fl / min(warm) / 1e6))
