#!/usr/bin/env python
"""bench.py -- images/sec of the Mask R-CNN R-50-FPN training step (forward + backward + SGD update) on
synthetic 800x1333 (padded to 800x1344) batches, 2 images per GPU, through the Blackwell hot path.

  python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's sm_100a path (both arms, see below)
  python bench.py --impl reference --gpus 1 --steps 1       # the reference's CPU path timed on the host cores
  python bench.py --impl aten                               # the same reference graph on ATen/cuDNN (GPU A/B arm)
  python bench.py --config faster_fwd|x101|dcn              # the other BASELINE.json configs (own JSON line each)
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU, NCCL)

Two arms of the product are measured in one run (`--model auto`):
  * "reference_graph" (headline when the reference mirror baseline/_ref is present): the UNMODIFIED reference
    `build_detection_model(cfg)` -- its GeneralizedRCNN, RPN, ROI heads, losses, samplers -- running over this
    repository's `maskrcnn_benchmark.layers` / `_C`, after mrb_b200.fuse.fuse_model(); eager (the reference's host
    code synchronises), optimizer = mrb_b200.optim.ParamArena (fused SGD, one NCCL all-reduce).
  * "harness": mrb_b200.model (from-scratch, fixed-shape, sync-free host code), whole step in ONE CUDA graph.
Prints ONE JSON line (rank 0).  `value` = whole-job images/s of the headline arm with inputs resident in HBM; `e2e`
= the same with pinned-host inputs (H2D of images + targets, D2H of the loss, inside the timed region); `roofline` =
the tcgen05 conv kernels, algorithmic FLOPs / CUDA-event time per launch, summed over the launches of one step;
`ops` = ROIAlign / NMS / focal-loss achieved GB/s measured in this run; `cpu_baseline` = the reference's own
GeneralizedRCNN train step on the host cores (bounded to one image)."""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

IMG_H, IMG_W, PAD_W = 800, 1333, 1344
GT_PER_IMAGE = 8

# SGD learning rate of the synthetic step.  The nets are randomly initialised (no checkpoint can be fetched), their first losses
# are O(100), and the deformable / ResNeXt variants diverge within a few dozen steps at 1e-4 (offsets are predicted by a
# kaiming-initialised conv, layers/misc.py:137-148); the update kernel does the same work at any rate.
BENCH_LR = {"mask_r50": 1e-4, "faster_fwd": 1e-4, "x101": 1e-5, "dcn": 1e-6}

CONFIGS = {
    # name: (reference yaml, images per GPU, train?, metric, workload text)
    "mask_r50": ("e2e_mask_rcnn_R_50_FPN_1x.yaml", 2, True, "images/sec Mask R-CNN R-50-FPN fwd+bwd @1333x800",
                 "e2e_mask_rcnn_R_50_FPN_1x train step (fwd + bwd + SGD momentum/wd update)"),
    "faster_fwd": ("e2e_faster_rcnn_R_50_FPN_1x.yaml", 2, False, "images/sec Faster R-CNN R-50-FPN forward-only @1333x800",
                   "e2e_faster_rcnn_R_50_FPN_1x forward-only (eval: proposals + detections)"),
    "x101": ("e2e_mask_rcnn_X_101_32x8d_FPN_1x.yaml", 1, True, "images/sec Mask R-CNN X-101-32x8d-FPN fwd+bwd @1333x800",
             "e2e_mask_rcnn_X_101_32x8d_FPN_1x train step (grouped 3x3 convs, stride in the 3x3)"),
    "dcn": ("dcn/e2e_mask_rcnn_dconv_R_50_FPN_1x.yaml", 2, True, "images/sec Mask R-CNN R-50-FPN-dconv fwd+bwd @1333x800",
            "configs/dcn/e2e_mask_rcnn_dconv_R_50_FPN_1x train step (deformable 3x3 in res3-res5)"),
}


HARNESS_CONFIGS = ("mask_r50", "faster_fwd", "x101", "dcn")     # configs the from-scratch harness model covers


def synth_batch(n, seed, device="cpu", pin=False):
    """images ~ N(0,1)*50 at 800x1333 zero-padded to 800x1344; 8 boxes per image with
    w,h ~ U[32,400], labels 1..80; masks are the box rectangles."""
    g = torch.Generator().manual_seed(seed)
    images = torch.zeros(n, 3, IMG_H, PAD_W)
    images[..., :IMG_W] = torch.randn(n, 3, IMG_H, IMG_W, generator=g) * 50
    boxes, labels = [], []
    for i in range(n):
        wh = torch.rand(GT_PER_IMAGE, 2, generator=g) * (400 - 32) + 32
        x1 = torch.rand(GT_PER_IMAGE, generator=g) * (IMG_W - 33)
        y1 = torch.rand(GT_PER_IMAGE, generator=g) * (IMG_H - 33)
        x2 = (x1 + wh[:, 0]).clamp(max=IMG_W - 1)
        y2 = (y1 + wh[:, 1]).clamp(max=IMG_H - 1)
        boxes.append(torch.stack([x1, y1, x2, y2], 1))
        labels.append(torch.randint(1, 81, (GT_PER_IMAGE,), generator=g))
    boxes, labels = torch.stack(boxes), torch.stack(labels)
    if pin:
        images, boxes, labels = images.pin_memory(), boxes.pin_memory(), labels.pin_memory()
    return images.to(device), boxes.to(device), labels.to(device)


def targets_of(boxes, labels):
    return [{"boxes": boxes[i], "labels": labels[i]} for i in range(boxes.shape[0])]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = str(index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", self.index], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.time()])

    def summary(self, windows):
        """Samples whose host timestamp lies in one of the [t0, t1] windows (the timed regions); if none landed there,
        all samples since start (warm-up included) are used and `window` says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for r in self.rows if any(t0 <= r[-1] <= t1 for t0, t1 in windows)]
        window = "timed regions"
        if not rows:
            rows, window = list(self.rows), "warm-up + timed regions (no sample landed inside the timed window)"
        sm = sorted(float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            self.t.join(timeout=2)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "_source": "fallback (B200_PROFILING.md)"}


# =========================================================================================== roofline (conv engine)
def conv_roofline(calls, peaks, device):
    """Time every distinct tcgen05 conv launch of one step in isolation (CUDA events on the launching stream, L2
    flushed between launches, median of 3) and aggregate: achieved = sum(algorithmic FLOPs) / sum(time)."""
    from mrb_b200 import ops
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    shapes = {}
    for c in calls:
        shapes[c] = shapes.get(c, 0) + 1
    tot_flops = tot_time = 0.0
    rows = []
    for key, count in sorted(shapes.items(), key=lambda kv: str(kv[0])):
        kind, n, cin, h, w, cout, k, stride, pad = key[:9]
        groups = key[9] if len(key) > 9 else 1
        x = torch.randn(n, cin, h, w, device=device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        kh, kw = (k, k) if isinstance(k, int) else k
        ph, pw = (pad, pad) if isinstance(pad, int) else pad
        ho, wo = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
        go = torch.randn(n, cout, ho, wo, device=device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        if groups > 1:
            wt = torch.randn(cout, 64, kh, kw, device=device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            if kind == "dgrad":
                fn = lambda: ops.conv2d_dgrad_grouped(go, wt.reshape(-1), (n, cin, h, w), stride, pad)  # noqa: E731
            elif kind == "wgrad":
                fn = lambda: ops.conv2d_wgrad_grouped(x, go, kh, stride, pad)  # noqa: E731
            else:
                fn = lambda: ops.conv2d_fwd(x, wt, None, None, None, stride, pad, True, grouped=True)  # noqa: E731
        else:
            wt = torch.randn(cout, cin, kh, kw, device=device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            if kind == "dgrad":
                fn = lambda: ops.conv2d_dgrad(go, wt, (n, cin, h, w), None, None, None, stride, pad)  # noqa: E731
            elif kind == "wgrad":
                fn = lambda: ops.conv2d_wgrad(x, go, wt.shape, stride, pad)  # noqa: E731
            else:
                fn = lambda: ops.conv2d_fwd(x, wt, None, None, None, stride, pad, True)  # noqa: E731
        fn()
        ts = []
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = sorted(ts)[1]
        flops = 2.0 * n * ho * wo * cout * (cin // groups) * kh * kw
        tot_flops += flops * count
        tot_time += t * count
        rows.append({"kind": kind, "n": n, "cin": cin, "h": h, "w": w, "cout": cout, "k": k if isinstance(k, int) else list(k),
                     "stride": stride, "groups": groups, "count": count, "us": round(t * 1e6, 1), "tflops": round(flops / t / 1e12, 1)})
    peak = peaks.get("bf16_tflops", 1590.0)
    ach = tot_flops / max(tot_time, 1e-12) / 1e12
    # the dominant shape of the step: its algorithmic HBM bytes (x + y + w of one launch) for the traffic comparison
    top = max(rows, key=lambda r: r["us"] * r["count"]) if rows else None
    rf = {"bound": "tensor",
          "kernel": "conv_tc_kernel + conv_wgrad_tc_kernel (tcgen05 implicit GEMM: the fwd, dgrad AND wgrad launches of one step)",
          "achieved": round(ach, 1), "peak": peak, "peak_source": peaks.get("_source", "fallback"), "unit": "TFLOP/s",
          "frac": round(ach / peak, 4), "traffic": None,
          "traffic_note": "not measured in this run (ncu --set full captures of the top shapes: profiles/ncu_conv_*_summary.json)",
          "launches_per_step": len(calls), "algorithmic_gflop_per_step": round(tot_flops / 1e9, 1),
          "kernel_ms_per_step": round(tot_time * 1e3, 2)}
    if top is not None:
        kh, kw = (top["k"], top["k"]) if isinstance(top["k"], int) else top["k"]
        ho, wo = top["h"] // top["stride"], top["w"] // top["stride"]
        rf["top_shape"] = {k: top[k] for k in ("kind", "n", "cin", "h", "w", "cout", "k", "stride", "count", "us", "tflops")}
        rf["top_shape"]["algorithmic_MB"] = round(2 * (top["n"] * top["cin"] * top["h"] * top["w"] + top["n"] * top["cout"] * ho * wo
                                                       + top["cout"] * top["cin"] * kh * kw) / 1e6, 1)
    return rf, rows


# =========================================================================================== op-level metrics
def ops_metrics(device, peaks):
    """ROIAlign / NMS / focal-loss numbers of BASELINE.json's metric, measured in this run (CUDA events, L2 flushed,
    median of 5; algorithmic bytes per SURVEY 8d)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _inputs
    from maskrcnn_benchmark import _C
    from mrb_b200 import ops
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_ops import roi_footprint_elems, timed
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    out = []
    # the reference's own CUDA kernels compiled for this GPU (oracle/_ref/cuda, checker build): timed beside ours
    try:
        import oracle
        R = oracle.ref_cuda()
    except Exception:
        R = None

    def rec(name, alg, t, **kw):
        d = {"op": name, "us": round(t * 1e6, 1), "algorithmic_MB": round(alg / 1e6, 2), "achieved_gbs": round(alg / t / 1e9, 1),
             "frac": round(alg / t / 1e9 / hbm, 4)}
        d.update(kw)
        out.append(d)

    # BASELINE config 1: 1x256x200x336 fp32, 100 boxes, 7x7, through the reference-facing `_C` call
    feat, rois = _inputs.roi_align_config1()
    n, c, h, w = feat.shape
    r = rois.shape[0]
    foot = min(roi_footprint_elems(rois, 0.25, h, w, c), n * c * h * w)
    rd = rois.to(device)
    xd = feat.to(device)
    kw = {}
    if R is not None:
        kw["reference_cuda_us"] = round(timed(lambda: R.roi_align_forward(xd, rd, 0.25, 7, 7, 2), flush, reps=5) * 1e6, 1)
    for layout, x in (("NCHW", xd), ("NHWC", feat.to(device).contiguous(memory_format=torch.channels_last))):
        t = timed(lambda: _C.roi_align_forward(x, rd, 0.25, 7, 7, 2), flush, reps=5)
        rec("roi_align_forward config1 (1x256x200x336 f32, R=100, 7x7)", 4 * (r * c * 49 + foot + 5 * r), t, layout=layout, **kw)
    g = torch.randn(r, c, 7, 7, device=device)
    kw = {}
    if R is not None:
        kw["reference_cuda_us"] = round(timed(lambda: R.roi_align_backward(g, rd, 0.25, 7, 7, n, c, h, w, 2), flush, reps=5) * 1e6, 1)
    t = timed(lambda: _C.roi_align_backward(g, rd, 0.25, 7, 7, n, c, h, w, 2), flush, reps=5)
    rec("roi_align_backward config1", 4 * (r * c * 49 + 2 * foot + n * c * h * w + 5 * r), t, layout="NCHW", **kw)
    # box-head sized single-level call (P2 of 2 images, 1024 boxes): ours vs the reference kernel
    featp2 = _inputs.fpn_features(2, 1)[0].to(device)
    roisp2 = _inputs.rois_for_level(1024, 2, 31, max_size=128)
    footp2 = min(roi_footprint_elems(roisp2, 0.25, 200, 336, 256), 2 * 256 * 200 * 336)
    rp2 = roisp2.to(device)
    kw = {}
    if R is not None:
        kw["reference_cuda_us"] = round(timed(lambda: R.roi_align_forward(featp2, rp2, 0.25, 7, 7, 2), flush, reps=5) * 1e6, 1)
    t = timed(lambda: _C.roi_align_forward(featp2, rp2, 0.25, 7, 7, 2), flush, reps=5)
    rec("roi_align_forward P2 (2x256x200x336 f32, R=1024, 7x7)", 4 * (1024 * 256 * 49 + footp2 + 5 * 1024), t, layout="NCHW", **kw)
    # the shapes the train step runs: fused multi-level ROIAlign on bf16 NHWC P2..P5 of 2 images
    feats = [f.to(torch.bfloat16).to(device).contiguous(memory_format=torch.channels_last) for f in _inputs.fpn_features(2, 2)]
    scales = (0.25, 0.125, 0.0625, 0.03125)
    for r, p, nhwc in ((1024, 7, False), (256, 14, True)):
        rois = _inputs.rois_for_level(r, 2, 40 + r).to(device)
        lv = torch.floor(4 + torch.log2(torch.sqrt((rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1)) / 224 + 1e-6)).clamp(2, 5).long() - 2
        foot = 0.0
        for l in range(4):
            m = lv == l
            if m.any():
                hh, ww = feats[l].shape[2:]
                foot += min(roi_footprint_elems(rois[m].cpu(), scales[l], hh, ww, 256), 2 * 256 * hh * ww)
        t = timed(lambda: ops.roi_align_fpn(feats, rois, scales, p, 2, out_nhwc=nhwc), flush, reps=5)
        rec("roi_align_fpn_fwd bf16 (in-step: R=%d, %dx%d)" % (r, p, p), 2 * (r * 256 * p * p + foot) + 20 * r, t)
        fr = [f.detach().clone().requires_grad_(True) for f in feats]
        y = ops.roi_align_fpn(fr, rois, scales, p, 2, out_nhwc=nhwc)
        go = torch.randn_like(y)
        t = timed(lambda: torch.autograd.grad(y, fr, go, retain_graph=True), flush, reps=5)
        maps = sum(f.numel() for f in feats)
        rec("roi_align_fpn_bwd bf16 (in-step: R=%d, %dx%d; incl. zero-fill + cast of the 4 level maps)" % (r, p, p),
            2 * r * 256 * p * p + 4 * 2 * foot + 4 * maps + 20 * r, t)
    # NMS: the 10 (image, level) problems of one RPN train step in one launch sequence; one 2000-box call through _C
    sizes = [2000] * 8 + [819] * 2
    bs = [_inputs.nms_boxes(s, 500 + i) for i, s in enumerate(sizes)]
    bd = torch.cat([b for b, _ in bs]).to(device)
    sd = torch.cat([s for _, s in bs]).to(device)
    t = timed(lambda: ops.nms_batched(bd, sd, sizes, 0.7), flush, reps=5)
    alg = sum(20 * s + 16 * s * math.ceil(s / 64) for s in sizes)
    rec("nms_batched (10 RPN problems, thr 0.7)", alg, t, us_per_problem=round(t * 1e5, 1), note="latency bound: bytes are tiny")
    b1, s1 = bs[0][0].to(device), bs[0][1].to(device)
    kw = {}
    if R is not None:
        kw["reference_cuda_us"] = round(timed(lambda: R.nms(torch.cat([b1, s1[:, None]], 1), 0.7), flush, reps=5) * 1e6, 1)
    t = timed(lambda: _C.nms(b1, s1, 0.7), flush, reps=5)
    rec("nms N=2000 through _C (incl. the 4-byte D2H sizing the result)", 20 * 2000 + 16 * 2000 * 32, t, note="latency bound", **kw)
    # focal loss, RetinaNet 800x1344: 201600 anchors x 80
    logits, targets = _inputs.focal_inputs(201600, 80, 0)
    ld, td = logits.to(device), targets.to(device)
    kw = {}
    if R is not None:
        kw["reference_cuda_us"] = round(timed(lambda: R.sigmoid_focalloss_forward(ld, td, 80, 2.0, 0.25), flush, reps=5) * 1e6, 1)
    t = timed(lambda: _C.sigmoid_focalloss_forward(ld, td, 80, 2.0, 0.25), flush, reps=5)
    rec("sigmoid_focalloss_forward (201600x80)", 4 * 201600 * 80 * 2 + 4 * 201600, t, **kw)
    dl = torch.rand_like(ld)
    kw = {}
    if R is not None:
        kw["reference_cuda_us"] = round(timed(lambda: R.sigmoid_focalloss_backward(ld, td, dl, 80, 2.0, 0.25), flush, reps=5) * 1e6, 1)
    t = timed(lambda: _C.sigmoid_focalloss_backward(ld, td, dl, 80, 2.0, 0.25), flush, reps=5)
    rec("sigmoid_focalloss_backward (201600x80)", 4 * 201600 * 80 * 3 + 4 * 201600, t, **kw)
    # deformable conv v2 3x3 on the three BASELINE config-5 layer shapes: tensor-core path of the model graph (mrb_b200.dcn) vs the
    # fp32 `_C` path vs the reference's own CUDA implementation (im2col + cuBLAS)
    try:
        from maskrcnn_benchmark import layers
        from mrb_b200 import dcn
        for (c, h, w) in ((128, 100, 168), (256, 50, 84), (512, 25, 42)):
            g = torch.Generator().manual_seed(c)
            x = torch.randn(2, c, h, w, generator=g).to(device)
            wt = (torch.randn(c, c, 3, 3, generator=g) / (3 * c ** 0.5)).to(device)
            om = torch.zeros(2, 32, h, w)
            om[:, :18] = torch.randn(2, 18, h, w, generator=g) * 2
            om[:, 18:27] = torch.randn(2, 9, h, w, generator=g)
            omd = om.to(device).contiguous(memory_format=torch.channels_last)
            x16 = x.bfloat16().contiguous(memory_format=torch.channels_last)
            w16 = wt.bfloat16().contiguous(memory_format=torch.channels_last)
            flops = 2.0 * 2 * h * w * c * c * 9
            t_tc = timed(lambda: dcn.deform_conv_nhwc(x16, omd, wt, w16, None, None, relu=True, modulated=True), flush, reps=5)
            off, msk = omd[:, :18].contiguous(), omd[:, 18:27].sigmoid().contiguous()
            t_f32 = timed(lambda: layers.modulated_deform_conv(x, off, msk, wt, None, 1, 1, 1, 1, 1), flush, reps=3)
            d = {"op": "modulated deform conv 3x3 fwd, 2x%dx%dx%d" % (c, h, w), "us": round(t_tc * 1e6, 1), "path": "tensor-core (sampler + tcgen05 GEMM)",
                 "tflops": round(flops / t_tc / 1e12, 1), "fp32_C_path_us": round(t_f32 * 1e6, 1)}
            if R is not None:
                out_r, bias0 = x.new_empty(2, c, h, w), x.new_zeros(c)
                d["reference_cuda_us"] = round(timed(lambda: R.modulated_deform_conv_forward(
                    x, wt, bias0, x.new_empty(0), off, msk, out_r, x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, False), flush, reps=3) * 1e6, 1)
            out.append(d)
    except Exception as e:
        out.append({"op": "modulated deform conv", "error": repr(e)[:200]})
    return {"peak_hbm_gbs": hbm, "timing": "CUDA events, L2 flushed between launches, median of 5",
            "reference_cuda": "the reference's csrc/cuda kernels compiled unmodified for sm_100a (oracle/build_ref.py::build_cuda), same "
                              "inputs, same timing" if R is not None else "not built on this box", "rows": out}


# =========================================================================================== CPU reference arm
def _pure_reference_package():
    """A `maskrcnn_benchmark` package made ONLY of the reference mirror (its own layers/ included) with `_C` = the
    reference's csrc/cpu kernels compiled by oracle/build_ref.py (+ the oracle's ROIAlign backward, which the
    reference does not have on the CPU: csrc/ROIAlign.h:44).  The product package is never imported here."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests", "_shims"))
    import mrb_test_compat  # noqa: F401
    import oracle
    root = None
    for cand in (os.environ.get("MRB_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "maskrcnn_benchmark", "modeling")):
            root = cand
            break
    if root is None:
        return None, None
    oracle.build()
    refc = oracle.ref()
    pkg = types.ModuleType("maskrcnn_benchmark")
    pkg.__path__ = [os.path.join(root, "maskrcnn_benchmark")]
    sys.modules["maskrcnn_benchmark"] = pkg
    c = types.ModuleType("maskrcnn_benchmark._C")
    if refc is not None:
        c.nms, c.roi_align_forward = refc.nms, refc.roi_align_forward
        kind = "reference"
    else:
        c.nms = lambda d, s, t: oracle.nms(d.contiguous(), s.contiguous(), t)
        c.roi_align_forward = lambda x, r, sc, ph, pw, s: oracle.roi_align_forward(x.contiguous(), r.contiguous(), sc, ph, pw, s)
        kind = "port"
    c.roi_align_backward = lambda g, r, sc, ph, pw, b, ch, h, w, s: oracle.roi_align_backward(g.contiguous(), r.contiguous(), sc, ph, pw, b, ch, h, w, s)
    pkg._C = c
    sys.modules["maskrcnn_benchmark._C"] = c
    return root, kind


def _ref_inputs(n, seed, device, with_masks=True):
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.image_list import ImageList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    images, boxes, labels = synth_batch(n, seed)
    il = ImageList(images.to(device), [(IMG_H, IMG_W)] * n)
    targets = []
    for i in range(n):
        t = BoxList(boxes[i], (IMG_W, IMG_H), mode="xyxy")
        t.add_field("labels", labels[i])
        if with_masks:
            polys = [[[float(b[0]), float(b[1]), float(b[2]), float(b[1]), float(b[2]), float(b[3]), float(b[0]), float(b[3])]]
                     for b in boxes[i]]
            t.add_field("masks", SegmentationMask(polys, (IMG_W, IMG_H), mode="poly"))
        targets.append(t.to(device))
    return il, targets


def cpu_reference_steps(cfg_name, n_images, threads, warmup, steps):
    """The reference's own GeneralizedRCNN on the host cores (fp32), model built once, `steps` timed steps after `warmup`:
    -> ([(images/s, seconds)], kind) or None without a reference mirror."""
    root, kind = _pure_reference_package()
    if root is None:
        return None
    torch.set_num_threads(threads)
    yaml, _, train, _, _ = CONFIGS[cfg_name]
    from maskrcnn_benchmark.config import cfg as _cfg
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    cfg = _cfg.clone()
    cfg.merge_from_file(os.path.join(root, "configs", yaml))
    cfg.merge_from_list(["MODEL.DEVICE", "cpu"])
    cfg.freeze()
    torch.manual_seed(0)
    model = build_detection_model(cfg)
    model.train(train)
    out = []
    for i in range(warmup + steps):
        il, targets = _ref_inputs(n_images, i, "cpu", with_masks=cfg.MODEL.MASK_ON)
        t0 = time.perf_counter()
        if train:
            model.zero_grad()
            losses = model(il, targets)
            sum(losses.values()).backward()
        else:
            with torch.no_grad():
                model(il)
        dt = time.perf_counter() - t0
        if i >= warmup:
            out.append((n_images / dt, dt))
    return out, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    yaml, per_gpu, train, metric, workload = CONFIGS[args.config]
    # torchrun exports OMP_NUM_THREADS=1, so the thread count is set explicitly: one thread per physical core of the CPUs this
    # process may use (all 128 hardware threads of the GPU box oversubscribe the ATen CPU convs: measured 88 s/image vs 6.5 s)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = int(os.environ.get("MRB_CPU_THREADS", 0)) or (max(1, avail // 2) if avail >= 16 else avail)
    r = cpu_reference_steps(args.config, 1, cores, args.warmup, args.steps)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "no reference mirror (baseline/_ref) on this box"}))
        return
    vals, kind = r
    v = sum(x for x, _ in vals) / len(vals)
    ms = sum(d for _, d in vals) / len(vals) * 1e3
    sample = ("1 image (800x1333 padded to 800x1344) per step: the reference's own GeneralizedRCNN (%s), fp32 ATen CPU convs, "
              "%s; ROIAlign backward from the oracle port (the reference has no CPU backward, csrc/ROIAlign.h:44)"
              % (workload, "reference csrc/cpu ROIAlign/NMS kernels compiled by oracle/build_ref.py" if kind == "reference"
                 else "oracle C port of the csrc/cpu kernels"))
    out = {"impl": "reference", "metric": metric, "value": round(v, 4), "unit": "images/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.config, args.gpus),
           "cpu_baseline": {"value": round(v, 4), "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": round(v, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def workload_config(cfg_name, n_gpus):
    yaml, per_gpu, train, metric, workload = CONFIGS[cfg_name]
    return {"workload": workload + ", synthetic 800x1333 images zero-padded to 800x1344 NCHW, 8 GT boxes/image, random-init weights",
            "reference_config": "configs/" + yaml,
            "global_batch": per_gpu * n_gpus, "images_per_gpu": per_gpu, "parallelism": "dp%d" % n_gpus,
            "l2": "per-step working set (activations + gradients, several GB) exceeds the 126 MB L2; no explicit flush"}


def _finish(world):
    """Leave a multi-rank run without tearing NCCL down: the communicator is referenced by the captured CUDA graph,
    and destroy_process_group() with ranks arriving minutes apart (rank 0 still measures the roofline and the CPU
    baseline) has been seen to block until the launcher's timeout.  All results are already printed."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


class Timer:
    """K steps bracketed by barrier + synchronize on both sides, CUDA events on the current stream."""

    def __init__(self, world):
        self.world = world
        self.windows = []

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def run(self, steps, fn):
        self.barrier()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for i in range(steps):
            out = fn(i)
        e1.record()
        self.barrier()
        self.windows.append((t0, time.time()))
        t = e0.elapsed_time(e1) * 1e-3
        if self.world > 1:
            import torch.distributed as dist
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt[0])
        return t, out


# =========================================================================================== arm: reference graph
def arm_reference_graph(args, cfg_name, device, rank, world, timer, aten=False):
    """The unmodified reference model over this repo's layers/_C (fused by mrb_b200.fuse), eager."""
    from mrb_b200 import refenv
    root = refenv.activate()
    if root is None:
        return None
    yaml, per_gpu, train, metric, workload = CONFIGS[cfg_name]
    if aten:
        os.environ["MRB_CONV_ENGINE"] = "0"
    from maskrcnn_benchmark.config import cfg as _cfg
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    from maskrcnn_benchmark.structures.image_list import ImageList
    from mrb_b200 import engine, ops
    cfg = _cfg.clone()
    cfg.merge_from_file(refenv.config_path(yaml))
    cfg.merge_from_list(["MODEL.DEVICE", "cuda"])
    cfg.freeze()
    torch.manual_seed(0)
    model = build_detection_model(cfg).to(device)
    model.train(train)
    report = None
    if world > 1:
        import torch.distributed as dist
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    opt = None
    if aten:
        model = model.to(memory_format=torch.channels_last)
        if train:
            # reference solver semantics (solver/build.py:7-20): bias lr x2, bias weight decay 0
            lr = BENCH_LR[cfg_name]
            groups = [{"params": [p], "lr": lr * (2 if "bias" in n else 1), "weight_decay": 0.0 if "bias" in n else 1e-4}
                      for n, p in model.named_parameters() if p.requires_grad]
            opt = torch.optim.SGD(groups, lr=lr, momentum=0.9)
    else:
        from mrb_b200.fuse import fuse_model
        from mrb_b200.model.backend import B200Backend
        be = B200Backend()
        engine.set_default_backend(be)
        report = fuse_model(model, be)
        if train:
            from mrb_b200.optim import ParamArena
            opt = ParamArena(model.named_parameters(), be, lr=BENCH_LR[cfg_name], momentum=0.9, weight_decay=1e-4, world_size=world)
            be.enable_overlap(True)
    n_batches = 4
    host = [synth_batch(per_gpu, 100 * rank + i, pin=True) for i in range(n_batches)]
    mask_on = bool(cfg.MODEL.MASK_ON)
    dev_batches = [_ref_inputs(per_gpu, 100 * rank + i, device, with_masks=mask_on) for i in range(n_batches)]
    graphed = {"enabled": False, "why": "the reference's host code synchronises (nonzero, per-image NMS sizing, CPU mask targets)"}
    if not aten and train and os.environ.get("MRB_REFGRAPH_SEGMENTS", "1") != "0":
        # the static-shape segment of the reference graph (backbone + FPN: ~70 launches forward, ~150 backward) replays as two
        # CUDA graphs; proposal selection, target assignment, ROI heads and losses stay the reference's eager Python
        try:
            from mrb_b200.graphed import graph_module
            seg = graph_module(model.backbone, (dev_batches[0][0].tensors,), backend=be, arena=opt)
            graphed = {"enabled": "segment", "segment": "model.backbone (body + fpn), forward and backward", "_seg": seg}
            if world > 1:
                # every gradient downstream of the backbone is complete when its backward starts: reduce that bucket while the
                # backbone's backward graph replays (ParamArena.sync() reduces the rest and waits)
                seg.pre_backward_hook = lambda: opt.reduce_bucket("heads")
            head = getattr(getattr(model, "rpn", None), "head", None)
            if head is not None and os.environ.get("MRB_REFGRAPH_SEGMENTS", "1") == "1":
                feats = model.backbone(dev_batches[0][0].tensors)        # a replay: the segment's static outputs
                seg2 = graph_module(head, (list(feats),), backend=be, arena=opt, share_inputs=True)
                graphed["segment"] += "; model.rpn.head (reads the first segment's outputs in place)"
                graphed["_seg2"] = seg2
        except Exception as e:  # noqa: BLE001  (any capture failure: stay eager, say why)
            graphed = {"enabled": False, "why": "segment capture failed: %s: %s" % (type(e).__name__, str(e)[:300])}
    host_targets = [_ref_inputs(per_gpu, 100 * rank + i, "cpu", with_masks=mask_on)[1] for i in range(n_batches)]

    def fwd_bwd(il, targets):
        if not train:
            with torch.no_grad():
                dets = model(il)
            return sum(d.bbox.sum() for d in dets)          # a scalar that depends on the detections
        if aten:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                losses = model(il, targets)
            loss = sum(losses.values())
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if world > 1:
                import torch.distributed as dist
                for p in model.parameters():
                    if p.grad is not None:
                        dist.all_reduce(p.grad)
                        p.grad /= world
            opt.step()
            return loss
        losses = model(il, targets)
        loss = sum(losses.values())
        loss.backward()
        opt.sync()
        opt.step()
        return loss

    def step_dev(i):
        il, targets = dev_batches[i % n_batches]
        return fwd_bwd(il, targets)

    # end to end: the images of step i+1 travel host -> device (pinned memory, copy stream) while step i computes -- what a
    # DataLoader with pin_memory and `images.to(device, non_blocking=True)` gives the reference's trainer (engine/trainer.py:64-68);
    # the boxes / labels of the targets are copied inside the step, the loss is read back (blocking) every step
    copy_stream = torch.cuda.Stream()
    staged = {}

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            buf = host[i % n_batches][0].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (buf, ev)

    def step_e2e(i):
        if i not in staged:
            prefetch(i)
        images, ev = staged.pop(i)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        images.record_stream(cur)
        prefetch(i + 1)
        il = ImageList(images, [(IMG_H, IMG_W)] * per_gpu)
        targets = [t.to(device) for t in host_targets[i % n_batches]]            # H2D of the boxes / labels
        return fwd_bwd(il, targets).detach().float().cpu()                       # D2H of the step's result

    first = None
    for i in range(args.warmup):
        r0 = step_dev(i)
        if first is None:
            first = round(float(r0.detach()), 4)
    ops.STATS["launches"] = 0
    ops.STATS["conv_calls"] = []
    for k in ("_seg", "_seg2"):
        if graphed.get(k) is not None:
            graphed[k].bypass = True                 # the counted step runs the segments eagerly: same kernels, visible to the counter
    step_dev(0)                                      # one counted step: libmrb launches + conv geometry per step
    torch.cuda.synchronize()
    for k in ("_seg", "_seg2"):
        if graphed.get(k) is not None:
            graphed[k].bypass = False
    launches_per_step = ops.STATS["launches"]
    conv_calls = list(ops.STATS["conv_calls"])
    ops.STATS["conv_calls"] = None
    t_dev, _ = timer.run(args.steps, step_dev)
    step_e2e(0)
    t_e2e, last = timer.run(args.steps, step_e2e)
    imgs = per_gpu * world * args.steps
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    return {"arm": "aten_reference_graph" if aten else "reference_graph",
            "model_path": ("reference GeneralizedRCNN (baseline/_ref mirror, unmodified) on ATen/cuDNN: bf16 autocast, channels_last, "
                           "torch.optim.SGD; ROIAlign/NMS from this repo's _C (the reference's CUDA sources need THC)") if aten else
                          "reference GeneralizedRCNN (baseline/_ref mirror, unmodified) over maskrcnn_benchmark.layers/_C of this "
                          "repo + mrb_b200.fuse.fuse_model (conv chains, poolers, heads; RPN / box post-processors, loss-target stages "
                          "and mask targets bound to the glue launches) + CUDA-graph replay of the static segments "
                          "(mrb_b200.graphed); the rest eager; ParamArena fused SGD",
            "value": round(imgs / t_dev, 3), "ms_per_step": round(t_dev / args.steps * 1e3, 2),
            "e2e": {"value": round(imgs / t_e2e, 3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "mode": "images of step i+1 copied host -> device on a copy stream during step i (pinned memory); targets' boxes / "
                            "labels copied and the loss read back (blocking) inside every step"},
            "gpu_launches": launches_per_step * args.steps, "libmrb_launches_per_step": launches_per_step,
            "conv_calls": conv_calls, "fuse_report": report, "result_first_step": first, "result_last_step": round(float(last), 4),
            "cuda_graph": _graphed_report(graphed),
            "engine_convs": engine.STATS["engine"], "aten_fallbacks": engine.STATS["aten"]}


def conv_flops(calls):
    tot = 0.0
    for key in calls:
        kind, n, cin, h, w, cout, k, stride, pad = key[:9]
        groups = key[9] if len(key) > 9 else 1
        kh, kw = (k, k) if isinstance(k, int) else k
        ph, pw = (pad, pad) if isinstance(pad, int) else pad
        ho, wo = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
        tot += 2.0 * n * ho * wo * cout * (cin // groups) * kh * kw
    return tot


def _in_step_profile(graph, conv_calls):
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        graph.replay()
        torch.cuda.synchronize()
    ks = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                if e.device_type == torch.autograd.DeviceType.CUDA)

    def union(iv):
        out, cs, ce = 0.0, None, None
        for a, b in sorted(iv):
            if cs is None:
                cs, ce = a, b
            elif a <= ce:
                ce = max(ce, b)
            else:
                out += ce - cs
                cs, ce = a, b
        return out + (ce - cs if cs is not None else 0.0)
    is_mrb = lambda n: "mrb::" in n or "bias_grad" in n  # noqa: E731
    conv = [(a, b) for a, b, n in ks if "conv_tc_kernel" in n or "conv_wgrad" in n]
    u_all, u_mrb, u_conv = union([(a, b) for a, b, _ in ks]), union([(a, b) for a, b, n in ks if is_mrb(n)]), union(conv)
    fl = conv_flops(conv_calls)
    return {"kernels": len(ks), "libmrb_kernels": sum(1 for _, _, n in ks if is_mrb(n)), "conv_launches": len(conv),
            "span_ms": round((ks[-1][1] - ks[0][0]) / 1e3, 3), "any_kernel_ms": round(u_all / 1e3, 3),
            "libmrb_kernel_ms": round(u_mrb / 1e3, 3), "conv_kernel_ms": round(u_conv / 1e3, 3),
            "glue_only_ms": round((u_all - u_mrb) / 1e3, 3), "conv_gflop": round(fl / 1e9, 1),
            "conv_tflops_in_step": round(fl / max(u_conv, 1e-9) / 1e6, 1),
            "note": "one CUDA-graph replay of the harness step under the CUPTI activity tracer, after the timed regions; "
                    "conv_kernel_ms = wall time with at least one tcgen05 conv kernel running"}


def _graphed_report(g):
    g = dict(g)
    seg, seg2 = g.pop("_seg", None), g.pop("_seg2", None)
    if seg is not None:
        g["replays"], g["eager_fallbacks"] = seg.replays, seg.fallbacks
    if seg2 is not None:
        g["replays_rpn_head"], g["eager_fallbacks_rpn_head"] = seg2.replays, seg2.fallbacks
    return g


# =========================================================================================== arm: harness
def arm_harness(args, cfg_name, device, rank, world, timer, sustained_s=0.0):
    import torch.distributed as dist
    from mrb_b200 import ops
    from mrb_b200.model import RCNNConfig, build_model
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.optim import FlatSGD, ParamArena
    yaml, per_gpu, train, metric, workload = CONFIGS[cfg_name]
    torch.manual_seed(0)
    # eval: the fused glue's post-processing is fixed-shape and sync-free, so the forward pass is capturable too
    use_graph = args.graph in ("on", "auto") and (train or os.environ.get("MRB_FUSED_GLUE", "1") != "0")
    kw = {}
    if cfg_name == "faster_fwd":
        kw = dict(mask_on=False)
    elif cfg_name == "x101":
        kw = dict(stage_blocks=(3, 4, 23, 3), num_groups=32, width_per_group=8, stride_in_1x1=False)
    elif cfg_name == "dcn":
        kw = dict(stage_with_dcn=(False, True, True, True))
    # graph capture needs a step without host synchronisation: fixed-shape mask head (see RCNNConfig)
    cfg = RCNNConfig(mask_rois_per_image=128 if (use_graph and train) else 0,
                     parallel_heads=(args.parallel_heads == "on" and args.overlap == "on" and args.optim == "arena"), **kw)
    be = B200Backend(wgrad=args.wgrad)
    model = build_model(cfg, backend=be, device=device)
    model.train(train)
    params = [p for p in model.parameters() if p.requires_grad]
    grad_sync = opt = None
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    if train:
        # SOLVER defaults of the reference (config/defaults.py:383-401): momentum 0.9, wd 1e-4, bias lr x2, bias wd 0
        if args.optim == "arena":
            opt = grad_sync = ParamArena(model.named_parameters(), model.be, lr=BENCH_LR[cfg_name], momentum=0.9, weight_decay=1e-4,
                                         world_size=world)
            model.be.enable_overlap(args.overlap == "on")
        else:
            if world > 1:
                from mrb_b200.parallel import FlatGradSync
                grad_sync = FlatGradSync(params, world)
            opt = FlatSGD(model.named_parameters(), lr=BENCH_LR[cfg_name], momentum=0.9, weight_decay=1e-4)
    sizes = [(IMG_H, IMG_W)] * per_gpu
    n_batches = 4
    host = [synth_batch(per_gpu, 100 * rank + i, pin=True) for i in range(n_batches)]
    dev = [tuple(t.to(device) for t in b) for b in host]
    defer_prep = bool(train and args.optim == "arena" and os.environ.get("MRB_DEFER_DGRAD_PREP", "1") != "0")
    if defer_prep:
        opt.defer_dgrad_prepare = True

    def eager_step(batch):
        images, boxes, labels = batch
        if not train:
            with torch.no_grad():
                dets = model(images, sizes)
            return sum(d["boxes"].sum() for d in dets)
        if defer_prep:
            model.be.prepare_async()       # data-gradient weights of the updated parameters: side stream, needed by backward only
        losses = model(images, sizes, targets_of(boxes, labels))
        loss = sum(losses.values())
        opt.zero_grad()
        if defer_prep:
            model.be.join_prepare()
        loss.backward()
        if grad_sync is not None:
            grad_sync.sync()               # NCCL all-reduce of one flat fp32 gradient buffer (mean over ranks)
        opt.step()
        if args.optim != "arena":
            model.be.refresh_weights(params)
        return loss

    graph_info = {"enabled": False}
    step = eager_step
    graph = static = static_loss = None
    if use_graph:
        static = tuple(torch.empty_like(t) for t in dev[0])
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    for a, b in zip(static, dev[i % n_batches]):
                        a.copy_(b)
                    eager_step(static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if opt is not None:
                opt.zero_grad()
            ops.STATS["launches"] = 0
            ops.STATS["conv_calls"] = []
            graph = torch.cuda.CUDAGraph()
            cap_stream = torch.cuda.Stream()
            with torch.cuda.graph(graph, stream=cap_stream):
                static_loss = eager_step(static)
            graph_info = {"enabled": True, "launches_per_step": ops.STATS["launches"], "conv_calls": list(ops.STATS["conv_calls"])}
            ops.STATS["conv_calls"] = None

            def step(batch):  # noqa: F811
                for a, b in zip(static, batch):
                    a.copy_(b, non_blocking=True)
                graph.replay()
                return static_loss
        except Exception as e:  # fall back to eager, and say so in the JSON line
            import traceback
            sys.stderr.write("CUDA graph capture failed, running eagerly:\n" + traceback.format_exc() + "\n")
            graph_info = {"enabled": False, "error": repr(e)[:300]}
            torch.cuda.synchronize()
            if args.optim == "arena":
                opt.grad.zero_()
            step = eager_step

    def step_e2e(hbatch):
        if graph_info["enabled"]:
            loss = step(hbatch)
        else:
            loss = step(tuple(t.to(device, non_blocking=True) for t in hbatch))
        return loss.detach().float().cpu()

    for i in range(args.warmup):
        step(dev[i % n_batches])
    ops.STATS["launches"] = 0
    ops.STATS["conv_calls"] = []
    t_dev, _ = timer.run(args.steps, lambda i: step(dev[i % n_batches]))
    if graph_info["enabled"]:
        launches = graph_info["launches_per_step"] * args.steps
        conv_calls = graph_info.pop("conv_calls")
    else:
        launches = ops.STATS["launches"]
        conv_calls = list(ops.STATS["conv_calls"])[:len(ops.STATS["conv_calls"]) // max(args.steps, 1)]
    ops.STATS["conv_calls"] = None
    # ---- end to end from pinned host memory: every step's batch is copied host -> device and every step's loss is
    # read back (blocking) inside the timed region.  With the captured step the H2D copy of batch i+1 is issued on a
    # copy stream while step i computes (what a prefetching loader with pinned memory does; the reference hands pinned
    # batches to `images.to(device)`, engine/trainer.py:64-68), a D2D copy moves it into the graph's static input.
    step_e2e(host[0])
    e2e_mode = "h2d serialised with the step"
    if graph_info["enabled"]:
        e2e_mode = "h2d of batch i+1 prefetched on a copy stream during step i"
        copy_stream = torch.cuda.Stream()
        staging = tuple(torch.empty_like(t) for t in static)
        ev_staged, ev_consumed = torch.cuda.Event(), torch.cuda.Event()

        def prefetch(hbatch):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_consumed)
                for a, b in zip(staging, hbatch):
                    a.copy_(b, non_blocking=True)
                ev_staged.record(copy_stream)

        def e2e_fn(i):
            if i == 0:
                prefetch(host[0])
            cur = torch.cuda.current_stream()
            cur.wait_event(ev_staged)
            for a, b in zip(static, staging):
                a.copy_(b, non_blocking=True)
            ev_consumed.record(cur)
            graph.replay()
            if i + 1 < args.steps:
                prefetch(host[(i + 1) % n_batches])
            return static_loss.detach().float().cpu()
        ev_consumed.record(torch.cuda.current_stream())
    else:
        def e2e_fn(i):
            return step_e2e(host[i % n_batches])
    t_e2e, last = timer.run(args.steps, e2e_fn)
    sustained = None
    if sustained_s > 0 and graph_info["enabled"]:
        n = max(args.steps, int(sustained_s / (t_dev / args.steps)))
        t_s, _ = timer.run(n, lambda i: step(dev[i % n_batches]))
        sustained = {"steps": n, "seconds": round(t_s, 2), "ms_per_step": round(t_s / n * 1e3, 3),
                     "value": round(per_gpu * world * n / t_s, 2), "window_index": len(timer.windows) - 1}
    in_step = None
    if graph_info["enabled"] and train and world == 1 and not args.no_roofline:
        # after the timed regions: ONE replay of the captured step under the CUPTI activity tracer -> wall time during which a
        # tcgen05 conv kernel is running inside the real step (union of their intervals: the weight-gradient kernels overlap
        # the data-gradient chain), total kernels, and the time no kernel of libmrb_b200.so runs
        try:
            in_step = _in_step_profile(graph, conv_calls)
        except Exception as e:  # profiling must never break the bench line
            in_step = {"error": repr(e)[:200]}
    imgs = per_gpu * world * args.steps
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    return {"arm": "harness", "in_step": in_step,
            "model_path": "mrb_b200.model (from-scratch module graph, reference state_dict keys; sync-free fixed-shape host code), "
                          "whole step in one CUDA graph" if graph_info["enabled"] else "mrb_b200.model, eager",
            "value": round(imgs / t_dev, 3), "ms_per_step": round(t_dev / args.steps * 1e3, 2),
            "e2e": {"value": round(imgs / t_e2e, 3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "mode": e2e_mode},
            "gpu_launches": launches, "conv_calls": conv_calls, "cuda_graph": graph_info, "result_last_step": round(float(last), 4),
            "sustained": sustained, "wgrad": model.be.wgrad_impl}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "aten"])
    ap.add_argument("--config", default="mask_r50", choices=sorted(CONFIGS))
    ap.add_argument("--model", default="auto", choices=["auto", "reference", "harness"],
                    help="auto: both arms, the reference graph over layers is the headline when baseline/_ref exists")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-ops", action="store_true")
    ap.add_argument("--sustained", type=float, default=5.0, help="seconds of back-to-back graph replays (harness arm); 0 = off")
    ap.add_argument("--dump-shapes", default=None, help="write the per-shape conv table (JSON) here")
    ap.add_argument("--wgrad", default="tc", choices=["tc", "cudnn"], help="weight-gradient kernel (A/B switch)")
    ap.add_argument("--optim", default="arena", choices=["arena", "flat"])
    ap.add_argument("--overlap", default="on", choices=["on", "off"])
    ap.add_argument("--parallel-heads", default="on", choices=["on", "off"])
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    assert torch.cuda.is_available(), "bench.py needs a CUDA device for --impl b200/aten (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    yaml, per_gpu, train, metric, workload = CONFIGS[args.config]
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()                 # well before the timed region: nvidia-smi needs a few hundred ms to start streaming
    timer = Timer(world)
    arms = []
    if args.impl == "aten":
        a = arm_reference_graph(args, args.config, device, rank, world, timer, aten=True)
        if a is None:
            if rank == 0:
                print(json.dumps({"impl": "aten", "unavailable": "no reference mirror (baseline/_ref) on this box"}))
            _finish(world)
            return
        arms.append(a)
    else:
        harness_ok = args.config in HARNESS_CONFIGS
        if args.model in ("auto", "reference"):
            try:
                a = arm_reference_graph(args, args.config, device, rank, world, timer)
            except Exception:
                if args.model == "reference":
                    raise
                import traceback
                sys.stderr.write("reference-graph arm failed, continuing with the harness arm:\n" + traceback.format_exc() + "\n")
                a = None
            if a is not None:
                arms.append(a)
        if args.model in ("auto", "harness") and harness_ok:
            try:
                arms.append(arm_harness(args, args.config, device, rank, world, timer,
                                        sustained_s=args.sustained if (world == 1 and train) else 0.0))
            except Exception:
                if not arms:
                    raise
                import traceback
                sys.stderr.write("harness arm failed:\n" + traceback.format_exc() + "\n")
    sampler.stop()
    if rank != 0:
        _finish(world)
        return
    head = arms[0]
    clocks = sampler.summary(timer.windows)
    peaks = load_peaks()
    out = {"metric": metric, "value": head["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": dict(workload_config(args.config, world), model_path=head["model_path"], arm=head["arm"]),
           "clocks": clocks, "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "cuda_graph": head["cuda_graph"],
           "result_last_step": head["result_last_step"]}
    if args.impl == "aten":
        out["impl"] = "aten"
    conv_calls = conv_from = None
    arms_out = {}
    for a in arms:
        cc = a.pop("conv_calls", None)
        # the conv roofline is taken over the harness arm's launches when it ran (fixed-shape mask head: the same launch list
        # every step, and the list the in-step profile refers to), else over the first arm's
        if cc and (conv_calls is None or a["arm"] == "harness"):
            conv_calls, conv_from = cc, a["arm"]
        arms_out[a["arm"]] = a
    out["arms"] = arms_out
    for a in arms:
        if a.get("sustained"):
            s = dict(a["sustained"])
            w = timer.windows[s.pop("window_index")]
            s["clocks"] = sampler.summary([w])
            s["note"] = "back-to-back CUDA-graph replays of the harness arm's train step"
            out["sustained"] = s
    out["library_ops"] = {"note": "in-house sm_100a kernels (libmrb_b200.so): conv forward / data gradient / weight gradient / bias "
                          "gradient (tcgen05 + TMA), fused multi-level ROIAlign fwd+bwd, NMS, max/sum pooling, fused SGD update; "
                          "detection glue (RPN decode / post-NMS selection / ROI assign-and-sample / anchor labelling) and the three loss "
                          "stages fwd+bwd (csrc/detect_glue.cu, loss_glue.cu); PyTorch: top-k, RPN anchor sampling, small casts / adds"}
    if not args.no_roofline and conv_calls and args.impl != "aten":
        try:
            rf, rows = conv_roofline(conv_calls, peaks, device)
            flop_img = {"mask_r50": 1631.0}.get(args.config)
            if flop_img:
                # conv-FLOP roofline of the whole step (~1631 GFLOP/image fwd+bwd upper bound)
                rf["step_conv_flop_roofline_frac"] = round(head["value"] / (peaks.get("bf16_tflops", 1590.0) * 1e3 / flop_img) / world, 4)
            rf["conv_calls_from"] = conv_from
            ins = arms_out.get("harness", {}).get("in_step")
            if ins and "conv_tflops_in_step" in ins:
                rf["in_step"] = {"achieved": ins["conv_tflops_in_step"], "frac": round(ins["conv_tflops_in_step"] / rf["peak"], 4),
                                 "conv_kernel_ms": ins["conv_kernel_ms"], "conv_gflop": ins["conv_gflop"],
                                 "what": "the same conv launches INSIDE the harness arm's captured step (CUPTI timeline of one replay): "
                                         "algorithmic conv FLOPs / wall time with a tcgen05 conv kernel running; `achieved`/`frac` above "
                                         "are the isolated, L2-flushed launches"}
            out["roofline"] = rf
            if args.dump_shapes:
                json.dump(rows, open(args.dump_shapes, "w"), indent=1)
        except Exception as e:
            out["roofline"] = {"error": repr(e)[:300]}
    if not args.no_ops and world == 1 and args.impl != "aten":
        try:
            out["ops"] = ops_metrics(device, peaks)
        except Exception as e:
            out["ops"] = {"error": repr(e)[:300]}
    if not args.no_cpu_baseline and world == 1:
        # in a subprocess: the reference arm builds a `maskrcnn_benchmark` made only of the reference mirror, which must
        # not share sys.modules with the product package loaded above
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config,
                                "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            out["cpu_baseline"] = json.loads(line).get("cpu_baseline", {"value": None, "note": line[:200]})
        except Exception as e:  # the baseline must never break the bench line
            out["cpu_baseline"] = {"value": None, "error": repr(e)[:200]}
    print(json.dumps(out))
    _finish(world)


if __name__ == "__main__":
    main()
