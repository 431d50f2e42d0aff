#!/usr/bin/env python
"""Tile-plan sweep for the tcgen05 conv kernel: for each layer shape, time the planner's own choice and a list of forced
plans (MRB_CONV_TILE = th,tw,bn,epi[,sets]) -- cold (L2 flushed, CUDA events, median of 5) and warm (20 launches
back-to-back in one CUDA graph).  Output: JSON lines; the planner's cost constants in csrc/conv_tc.cu are fitted to it.
Usage: python tools/sweep_conv.py [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mrb_b200 import ops  # noqa: E402

DEV = "cuda:0"
# (kind, n, cin, h, w, cout, k, res, [plans])
S33 = ["8,16,256,0", "8,16,128,1,2", "8,16,128,1,3", "4,32,256,0", "4,32,128,1,2"]
SHAPES = [
    ("fwd", 2, 256, 50, 84, 256, 3, False, ["4,32,256,0", "4,32,128,1,2", "10,12,128,1,2", "10,12,256,0", "10,12,128,0", "10,12,128,1,1", "10,12,64,1,1", "10,12,64,0"]),
    ("dgrad", 2, 256, 50, 84, 256, 3, False, ["4,32,128,1,2", "10,12,128,1,2", "10,12,128,1,3", "10,12,256,0"]),
    ("fwd", 2, 512, 25, 42, 512, 3, False, ["9,14,128,1,2", "9,14,64,1,2", "9,14,256,0", "9,14,128,0", "9,14,128,1,1", "9,14,64,0", "9,14,64,1,1"]),
    ("fwd", 2, 128, 100, 168, 128, 3, False, ["8,16,128,1,2", "8,16,128,0", "8,16,128,1,1"]),
    ("fwd", 2, 64, 200, 336, 64, 3, False, ["8,16,64,1,2", "10,12,64,1,2", "8,16,64,0"]),
    ("fwd", 2, 256, 200, 336, 256, 3, False, S33),
    ("fwd", 2, 64, 200, 336, 256, 1, False, ["1,128,128,1,2", "1,128,128,1,3", "1,128,256,0", "1,128,64,1,2", "1,128,64,1,3"]),
    ("fwd", 2, 64, 200, 336, 256, 1, True, ["1,128,128,1,2", "1,128,128,1,3", "1,128,64,1,3"]),
    ("fwd", 2, 256, 200, 336, 64, 1, False, ["1,128,64,1,2", "1,128,64,1,3", "1,128,64,0"]),
    ("fwd", 2, 256, 50, 84, 1024, 1, True, ["1,128,128,1,2", "1,128,128,1,3", "1,128,256,0", "1,114,128,1,3"]),
    ("fwd", 2, 1024, 50, 84, 256, 1, False, ["1,128,128,1,2", "1,128,256,0", "1,128,128,0", "1,128,128,1,1", "1,128,64,0"]),
    ("fwd", 2, 512, 25, 42, 2048, 1, True, ["1,128,128,1,2", "1,128,128,1,3", "1,128,256,0"]),
    ("fwd", 2, 2048, 25, 42, 512, 1, False, ["1,128,128,1,2", "1,128,256,0", "1,128,64,1,2"]),
    ("fwd", 256, 256, 14, 14, 256, 3, False, ["8,16,256,0", "9,14,256,0", "7,14,256,0", "9,14,128,1,2"]),
    ("fwd", 1024, 12544, 1, 1, 1024, 1, False, ["1,128,128,1,2", "1,128,256,0", "1,128,64,1,2"]),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
out = []
for kind, n, cin, h, w, cout, k, res, plans in SHAPES:
    pad = k // 2
    x = torch.randn(n, cin, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, k, k, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    go = torch.randn(n, cout, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    sc, bi = torch.rand(cout, device=DEV) + 0.5, torch.randn(cout, device=DEV)
    rs = torch.randn(n, cout, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
    if kind == "fwd":
        fn = lambda: ops.conv2d_fwd(x, wt, sc, bi, rs, 1, pad, True)  # noqa: E731
    else:
        prep = ops.prepare_dgrad_weights([wt], [None])[0]
        mk = torch.randn(n, cin, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        fn = lambda: ops.conv2d_dgrad(go, wt, x.shape, None, None, mk, 1, pad, prepared=prep)  # noqa: E731
    fl = 2.0 * n * h * w * cout * cin * k * k
    for plan in ["auto"] + plans:
        if plan == "auto":
            os.environ.pop("MRB_CONV_TILE", None)
        else:
            os.environ["MRB_CONV_TILE"] = plan
        try:
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
            rec = {"kind": kind, "n": n, "cin": cin, "h": h, "w": w, "cout": cout, "k": k, "res": res, "plan": plan,
                   "cold_us": round(sorted(cold)[2], 1), "warm_us": round(min(warm), 1), "cold_tflops": round(fl / sorted(cold)[2] / 1e6, 1),
                   "warm_tflops": round(fl / min(warm) / 1e6, 1)}
        except Exception as e:
            rec = {"kind": kind, "n": n, "cin": cin, "h": h, "w": w, "cout": cout, "k": k, "res": res, "plan": plan, "error": repr(e)[:200]}
        out.append(rec)
        print(json.dumps(rec), flush=True)
os.environ.pop("MRB_CONV_TILE", None)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=0)
