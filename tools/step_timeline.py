"""Kernel timeline of ONE eager train step (the bench.py workload) from the CUPTI activity trace (torch.profiler):
per-kernel device time in normal back-to-back execution (warm caches, not serialised), aggregated by kernel, plus
the idle time between kernels.  Cheaper than an ncu launch list (seconds instead of minutes); the committed ncu list
stays the reference evidence.  Usage: python tools/step_timeline.py [--top 45] [--json out.json]"""
import argparse
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")]
import torch  # noqa: E402
import bench  # noqa: E402


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.search(r"(mrb::\w+)", name)
    if m:
        return m.group(1)
    m = re.search(r"(\w+Functor\w*|\w+_kernel_cuda\w*|\w+_kernel_impl\w*|launch_clamp_scalar|compare_scalar_kernel|"
                  r"multi_tensor_apply_kernel|\w+topk::\w+|radixSort\w+|DeviceRadixSort\w+|DeviceScan\w+|DeviceSelect\w+|"
                  r"DeviceReduce\w+|reduce_kernel|index_elementwise_kernel|_scatter_gather_elementwise_kernel|"
                  r"vectorized_gather_kernel|nccl\w+|max_pool\w+|upsample\w+|cat\w*Kernel\w*|CatArrayBatchedCopy\w*|"
                  r"distribution_elementwise\w+|arange\w+|nll_loss\w+|smooth_l1\w+|softmax\w+|Memset|Memcpy \w+)", name)
    base = name.split("<")[0].split("(")[0][-40:]
    return (base + "|" + m.group(1)) if m else name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--json", default=None)
    ap.add_argument("--convs", default=None, help="write the in-situ duration of every tcgen05 conv launch of the step, "
                    "joined with its geometry (launch order == call order), aggregated per shape, to this JSON")
    ap.add_argument("--by-op", action="store_true", help="also attribute device time to the launching aten op, its "
                    "input shapes and the innermost mrb_b200/ source line (PyTorch glue only)")
    args = ap.parse_args()
    from mrb_b200.model import RCNNConfig, build_model
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.optim import ParamArena
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = build_model(RCNNConfig(mask_rois_per_image=128), backend=B200Backend(), device=dev).train()
    opt = ParamArena(model.named_parameters(), model.be, lr=1e-4)
    sizes = [(bench.IMG_H, bench.IMG_W)] * bench.IMGS_PER_GPU
    batches = [tuple(t.to(dev) for t in bench.synth_batch(bench.IMGS_PER_GPU, i)) for i in range(2)]

    def step(b):
        images, boxes, labels = b
        loss = sum(model(images, sizes, bench.targets_of(boxes, labels)).values())
        loss.backward()
        opt.sync()
        opt.step()
        return loss

    for i in range(3):
        step(batches[i % 2])
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    from mrb_b200 import ops
    ops.STATS["conv_calls"] = []
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=args.by_op,
                 with_stack=args.by_op) as prof:
        step(batches[1])
        torch.cuda.synchronize()
    conv_calls, ops.STATS["conv_calls"] = ops.STATS["conv_calls"], None
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    agg = collections.defaultdict(lambda: [0, 0.0])
    busy, last_end = 0.0, None
    for s, e, n in ks:
        k = short(n)
        agg[k][0] += 1
        agg[k][1] += e - s
        busy += e - s
    span = ks[-1][1] - ks[0][0]
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    mine = sum(v[1] for k, v in rows if k.startswith("mrb::"))
    print("device activities: %d   span %.2f ms   sum of durations %.2f ms   libmrb share of busy time %.1f%%"
          % (len(ks), span / 1e3, busy / 1e3, 100 * mine / busy))
    # exposure: wall time of the step during which NO libmrb kernel is running on any stream (pure glue), and during
    # which nothing at all is running (launch gaps; large in this eager trace, absent under the CUDA graph)
    def union(iv):
        iv = sorted(iv)
        out, cs, ce = 0.0, None, None
        for a, b in iv:
            if cs is None:
                cs, ce = a, b
            elif a <= ce:
                ce = max(ce, b)
            else:
                out += ce - cs
                cs, ce = a, b
        return out + (ce - cs if cs is not None else 0.0)
    u_all = union([(a, b) for a, b, n in ks])
    u_mrb = union([(a, b) for a, b, n in ks if "mrb::" in n or "bias_grad" in n])
    print("wall time with any kernel running %.2f ms; with a libmrb kernel running %.2f ms; glue-only %.2f ms; idle %.2f ms"
          % (u_all / 1e3, u_mrb / 1e3, (u_all - u_mrb) / 1e3, (span - u_all) / 1e3))
    glue = collections.defaultdict(float)
    mrb_iv = sorted((a, b) for a, b, n in ks if "mrb::" in n or "bias_grad" in n)
    import bisect
    starts = [a for a, _ in mrb_iv]
    for a, b, n in ks:
        if "mrb::" in n or "bias_grad" in n:
            continue
        i = bisect.bisect_right(starts, a) - 1
        covered = i >= 0 and mrb_iv[i][1] >= b       # fully under one libmrb kernel (approximation)
        if not covered:
            glue[short(n)] += b - a
    print("exposed glue kernels (not running under a libmrb kernel), top 15 by time:")
    for k, t in sorted(glue.items(), key=lambda kv: -kv[1])[:15]:
        print("  %8.1f us  %s" % (t, k))
    print("| kernel | launches | total us | share of busy |")
    print("|---|---|---|---|")
    for k, (n, t) in rows[:args.top]:
        print("| %s | %d | %.1f | %.1f%% |" % (k, n, t, 100 * t / busy))
    rest = rows[args.top:]
    if rest:
        print("| (%d more) | %d | %.1f | %.1f%% |" % (len(rest), sum(v[0] for _, v in rest), sum(v[1] for _, v in rest),
                                                     100 * sum(v[1] for _, v in rest) / busy))
    if args.convs:
        tc = [(s_, e - s_) for s_, e, n in ks if "conv_tc_kernel" in n]
        wg = [(s_, e - s_) for s_, e, n in ks if "conv_wgrad_tc_kernel" in n]
        c_tc = [c for c in conv_calls if c[0] != "wgrad"]
        c_wg = [c for c in conv_calls if c[0] == "wgrad"]
        assert len(tc) == len(c_tc) and len(wg) == len(c_wg), (len(tc), len(c_tc), len(wg), len(c_wg))
        per = collections.defaultdict(list)
        for (st, d), c in list(zip(tc, c_tc)) + list(zip(wg, c_wg)):
            per[c].append(d)
        rows_c = []
        for c, ds in per.items():
            kind, n, cin, h, w, cout, k, stride, pad = c
            kh, kw = (k, k) if isinstance(k, int) else k
            ph, pw = (pad, pad) if isinstance(pad, int) else pad
            ho, wo = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
            fl = 2.0 * n * ho * wo * cout * cin * kh * kw
            rows_c.append({"kind": kind, "n": n, "cin": cin, "h": h, "w": w, "cout": cout, "k": k, "stride": stride,
                           "count": len(ds), "us": round(sum(ds) / len(ds), 1), "us_min": round(min(ds), 1),
                           "tflops": round(fl / (sum(ds) / len(ds)) / 1e6, 1)})
        rows_c.sort(key=lambda r: -r["us"] * r["count"])
        json.dump(rows_c, open(args.convs, "w"), indent=1)
        print("\nin-situ conv time: fwd+dgrad %.1f us, wgrad %.1f us" % (sum(d for _, d in tc), sum(d for _, d in wg)))
    if args.by_op:
        ops_ = collections.defaultdict(lambda: [0, 0.0])
        for e in prof.events():
            if e.device_type != torch.autograd.DeviceType.CPU or not e.kernels:
                continue
            if any(c.kernels for c in (e.cpu_children or [])):
                continue                                   # attribute to the innermost op that owns the launch
            t = sum(k.duration for k in e.kernels)
            if any("mrb::" in k.name for k in e.kernels):
                continue
            where = next((f for f in (e.stack or []) if "mrb_b200/" in f or "bench.py" in f or "step_timeline" in f), "?")
            where = where.split("mrb_b200/")[-1][:60]
            shp = str([list(x) for x in (e.input_shapes or []) if x])[:70]
            ops_[(e.name, shp, where)][0] += len(e.kernels)
            ops_[(e.name, shp, where)][1] += t
        print("\n| aten op | input shapes | source | kernels | total us |")
        print("|---|---|---|---|---|")
        for (n, shp, where), (c, t) in sorted(ops_.items(), key=lambda kv: -kv[1][1])[:args.top]:
            print("| %s | %s | %s | %d | %.1f |" % (n, shp, where, c, t))
    if args.json:
        json.dump({"span_us": span, "busy_us": busy, "kernels": [{"name": k, "launches": n, "us": t} for k, (n, t) in rows]},
                  open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
