"""Kernel timeline of ONE replay of the captured train step (the harness arm of bench.py) from the CUPTI activity trace:
span of the replay, time with any kernel running, with a libmrb kernel running, idle gaps, and the largest exposed
non-libmrb kernels / gaps in stream order on the critical stream.  Usage: python tools/graph_timeline.py [--top 30]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200"), os.path.join(ROOT, "tools")]
import torch  # noqa: E402
import bench  # noqa: E402
from step_timeline import short  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=30)
    args = ap.parse_args()
    from mrb_b200.model import RCNNConfig, build_model
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.optim import ParamArena
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    be = B200Backend()
    model = build_model(RCNNConfig(mask_rois_per_image=128, parallel_heads=True), backend=be, device=dev).train()
    opt = ParamArena(model.named_parameters(), be, lr=1e-4, momentum=0.9, weight_decay=1e-4)
    be.enable_overlap(True)
    sizes = [(bench.IMG_H, bench.IMG_W)] * 2
    batches = [tuple(t.to(dev) for t in bench.synth_batch(2, i)) for i in range(2)]
    static = tuple(torch.empty_like(t) for t in batches[0])

    def step(b):
        images, boxes, labels = b
        loss = sum(model(images, sizes, bench.targets_of(boxes, labels)).values())
        loss.backward()
        opt.sync()
        opt.step()
        return loss
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            for a, b in zip(static, batches[i % 2]):
                a.copy_(b)
            step(static)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=torch.cuda.Stream()):
        step(static)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    span = ks[-1][1] - ks[0][0]

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
    is_mrb = lambda n: "mrb::" in n or "bias_grad" in n  # noqa: E731
    u_all = union([(a, b) for a, b, n in ks])
    u_mrb = union([(a, b) for a, b, n in ks if is_mrb(n)])
    u_conv = union([(a, b) for a, b, n in ks if "conv_tc_kernel" in n or "conv_wgrad" in n])
    busy = sum(b - a for a, b, n in ks)
    print("graph replay: %d kernels, span %.2f ms, sum of durations %.2f ms" % (len(ks), span / 1e3, busy / 1e3))
    print("wall time with any kernel running %.2f ms | a libmrb kernel %.2f ms | a tcgen05 conv kernel %.2f ms | glue only %.2f ms | idle %.2f ms"
          % (u_all / 1e3, u_mrb / 1e3, u_conv / 1e3, (u_all - u_mrb) / 1e3, (span - u_all) / 1e3))
    # time segments during which NO libmrb kernel runs: what fills them?
    mrb_iv = sorted((a, b) for a, b, n in ks if is_mrb(n))
    merged = []
    for a, b in mrb_iv:
        if merged and a <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], b)
        else:
            merged.append([a, b])
    gaps = [(merged[i][1], merged[i + 1][0]) for i in range(len(merged) - 1) if merged[i + 1][0] - merged[i][1] > 2.0]
    gaps.sort(key=lambda g: g[0] - g[1])
    print("largest intervals without any libmrb kernel (us from start, length, kernels running inside):")
    t0 = ks[0][0]
    for a, b in gaps[:args.top]:
        inside = collections.Counter(short(n) for s, e, n in ks if s < b and e > a and not is_mrb(n))
        before = next((short(n) for s, e, n in reversed(ks) if is_mrb(n) and e <= a + 0.01), "?")
        after = next((short(n) for s, e, n in ks if is_mrb(n) and s >= b - 0.01), "?")
        print("  at %8.1f  len %7.1f us  after %-28s before %-28s : %s" % (a - t0, b - a, before[-28:], after[-28:],
                                                                          ", ".join("%s x%d" % (k[-40:], c) for k, c in inside.most_common(4))))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for a, b, n in ks:
        agg[short(n)][0] += 1
        agg[short(n)][1] += b - a
    print("| kernel | launches | total us | share of busy |\n|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print("| %s | %d | %.1f | %.1f%% |" % (k, c, t, 100 * t / busy))


if __name__ == "__main__":
    main()
