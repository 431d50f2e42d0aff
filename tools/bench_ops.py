#!/usr/bin/env python
"""Op-level roofline numbers for the HBM-bound kernels: ROIAlign fwd/bwd (achieved GB/s vs the
ALGORITHMIC bytes: output + unique ROI footprint + rois), NMS (us per call; bytes are tiny), focal loss.
CUDA events on the launching stream, L2 flushed between timed launches, median of 7.  Also times the
oracle / reference CPU kernels on the same inputs (1 thread, as the reference's kernels are serial).
Writes one JSON document to stdout (committed under profiles/)."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import _inputs  # noqa: E402

DEV = "cuda:0"


def timed(fn, flush, reps=7):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sorted(ts)[len(ts) // 2]


def roi_footprint_elems(rois, scale, H, W, C):
    """sum_r C * fh_r * fw_r: rows/cols of the level map a ROI's samples can touch."""
    x1, y1, x2, y2 = (rois[:, i] * scale for i in (1, 2, 3, 4))
    x2 = torch.maximum(x2, x1 + 1)
    y2 = torch.maximum(y2, y1 + 1)
    fh = (torch.floor(y2.clamp(max=H - 1)) + 2 - torch.floor(y1.clamp(min=0))).clamp(min=1, max=H)
    fw = (torch.floor(x2.clamp(max=W - 1)) + 2 - torch.floor(x1.clamp(min=0))).clamp(min=1, max=W)
    return float((fh * fw).sum()) * C


def main():
    from maskrcnn_benchmark import _C
    from mrb_b200 import ops
    import oracle
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    out = {"peak_hbm_gbs": hbm, "peak_source": "MEASURED_PEAKS.json" if "how" in peaks else "fallback", "ops": []}
    torch.set_num_threads(1)

    # ---------------- ROIAlign, single level (BASELINE config 1 and a box-head sized call on P2)
    for name, (feat, rois) in {"config1 1x256x200x336 R=100": _inputs.roi_align_config1(),
                               "P2 2x256x200x336 R=1024": (_inputs.fpn_features(2, 1)[0], _inputs.rois_for_level(1024, 2, 31, max_size=128))}.items():
        n, c, h, w = feat.shape
        r = rois.shape[0]
        for p, s in ((7, 2), (14, 2)):
            alg = 4 * (r * c * p * p + min(roi_footprint_elems(rois, 0.25, h, w, c), n * c * h * w) + 5 * r)
            for layout, x in (("NCHW", feat.to(DEV)), ("NHWC", feat.to(DEV).contiguous(memory_format=torch.channels_last))):
                rd = rois.to(DEV)
                t = timed(lambda: _C.roi_align_forward(x, rd, 0.25, p, p, s), flush)
                rec = {"op": "roi_align_forward", "case": name, "P": p, "S": s, "layout": layout, "us": round(t * 1e6, 1),
                       "algorithmic_MB": round(alg / 1e6, 2), "achieved_gbs": round(alg / t / 1e9, 1), "frac": round(alg / t / 1e9 / hbm, 4)}
                out["ops"].append(rec)
            g = torch.randn(r, c, p, p, device=DEV)
            algb = 4 * (r * c * p * p + 2 * min(roi_footprint_elems(rois, 0.25, h, w, c), n * c * h * w) + n * c * h * w + 5 * r)
            t = timed(lambda: _C.roi_align_backward(g, rd, 0.25, p, p, n, c, h, w, s), flush)
            out["ops"].append({"op": "roi_align_backward", "case": name, "P": p, "S": s, "layout": "NCHW", "us": round(t * 1e6, 1),
                               "algorithmic_MB": round(algb / 1e6, 2), "achieved_gbs": round(algb / t / 1e9, 1),
                               "frac": round(algb / t / 1e9 / hbm, 4)})
        if r <= 128:
            t0 = time.perf_counter(); oracle.roi_align_forward(feat, rois, 0.25, 7, 7, 2); tc = time.perf_counter() - t0
            ref = oracle.ref()
            rec = {"op": "roi_align_forward CPU", "case": name, "P": 7, "S": 2, "oracle_ms": round(tc * 1e3, 1)}
            if ref is not None:
                t0 = time.perf_counter(); ref.roi_align_forward(feat, rois, 0.25, 7, 7, 2); rec["reference_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            out["ops"].append(rec)

    # ---------------- fused multi-level ROIAlign (the Pooler), bf16 NHWC features, train-step sizes
    feats = [f.to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last) for f in _inputs.fpn_features(2, 2)]
    for r, p, nhwc in ((1024, 7, False), (256, 14, True), (2000, 7, False)):
        rois = _inputs.rois_for_level(r, 2, 40 + r).to(DEV)
        scales = (0.25, 0.125, 0.0625, 0.03125)
        lv = torch.floor(4 + torch.log2(torch.sqrt((rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1)) / 224 + 1e-6)).clamp(2, 5).long() - 2
        foot = 0.0
        for l in range(4):
            m = lv == l
            if m.any():
                hh, ww = feats[l].shape[2:]
                foot += min(roi_footprint_elems(rois[m].cpu(), scales[l], hh, ww, 256), 2 * 256 * hh * ww)
        alg = 2 * (r * 256 * p * p + foot) + 20 * r
        t = timed(lambda: ops.roi_align_fpn(feats, rois, scales, p, 2, out_nhwc=nhwc), flush)
        out["ops"].append({"op": "roi_align_fpn_fwd bf16", "R": r, "P": p, "out": "NHWC" if nhwc else "NCHW", "us": round(t * 1e6, 1),
                           "algorithmic_MB": round(alg / 1e6, 2), "achieved_gbs": round(alg / t / 1e9, 1), "frac": round(alg / t / 1e9 / hbm, 4)})

    # ---------------- NMS
    for n in (819, 1000, 2000, 6000, 12000):
        boxes, scores = _inputs.nms_boxes(n, n)
        bd, sd = boxes.to(DEV), scores.to(DEV)
        t = timed(lambda: _C.nms(bd, sd, 0.7), flush)          # includes the 4-byte D2H that sizes the result
        k = int(_C.nms(bd, sd, 0.7).numel())
        alg = 20 * n + 16 * n * math.ceil(n / 64) + 8 * k
        t0 = time.perf_counter(); oracle.nms(boxes, scores, 0.7); tc = time.perf_counter() - t0
        rec = {"op": "nms", "N": n, "kept": k, "us_per_call": round(t * 1e6, 1), "algorithmic_MB": round(alg / 1e6, 3),
               "achieved_gbs": round(alg / t / 1e9, 2), "oracle_cpu_ms": round(tc * 1e3, 2)}
        ref = oracle.ref()
        if ref is not None:
            t0 = time.perf_counter(); ref.nms(boxes, scores, 0.7); rec["reference_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        out["ops"].append(rec)
    # batched: the 10 (image, level) problems of one RPN train step
    sizes = [2000, 2000, 2000, 2000, 2000, 2000, 2000, 2000, 819, 819]
    bs = [_inputs.nms_boxes(s, 500 + i) for i, s in enumerate(sizes)]
    bd = torch.cat([b for b, _ in bs]).to(DEV)
    sd = torch.cat([s for _, s in bs]).to(DEV)
    t = timed(lambda: ops.nms_batched(bd, sd, sizes, 0.7), flush)
    out["ops"].append({"op": "nms_batched (10 RPN problems)", "us": round(t * 1e6, 1), "us_per_problem": round(t * 1e5, 1)})

    # ---------------- focal loss (RetinaNet 800x1344: 201600 anchors x 80)
    logits, targets = _inputs.focal_inputs(201600, 80, 0)
    ld, td = logits.to(DEV), targets.to(DEV)
    t = timed(lambda: _C.sigmoid_focalloss_forward(ld, td, 80, 2.0, 0.25), flush)
    alg = 4 * 201600 * 80 * 2 + 4 * 201600
    out["ops"].append({"op": "sigmoid_focalloss_forward", "A": 201600, "us": round(t * 1e6, 1), "achieved_gbs": round(alg / t / 1e9, 1),
                       "frac": round(alg / t / 1e9 / hbm, 4)})
    dl = torch.rand_like(ld)
    t = timed(lambda: _C.sigmoid_focalloss_backward(ld, td, dl, 80, 2.0, 0.25), flush)
    alg = 4 * 201600 * 80 * 3 + 4 * 201600
    out["ops"].append({"op": "sigmoid_focalloss_backward", "A": 201600, "us": round(t * 1e6, 1), "achieved_gbs": round(alg / t / 1e9, 1),
                       "frac": round(alg / t / 1e9 / hbm, 4)})
    # ---------------- deformable conv v2 on the three R-50 DCN layer shapes (configs/dcn/*_R_50_FPN_1x.yaml)
    from maskrcnn_benchmark import layers
    for (c, h, w) in ((128, 100, 168), (256, 50, 84), (512, 25, 42)):
        g = torch.Generator().manual_seed(c)
        x = torch.randn(2, c, h, w, generator=g).to(DEV).requires_grad_(True)
        wt = (torch.randn(c, c, 3, 3, generator=g) / (3 * c ** 0.5)).to(DEV).requires_grad_(True)
        off = (torch.randn(2, 18, h, w, generator=g) * 2).to(DEV).requires_grad_(True)
        msk = torch.rand(2, 9, h, w, generator=g).to(DEV).requires_grad_(True)
        flops = 2.0 * 2 * h * w * c * c * 9
        tf = timed(lambda: layers.modulated_deform_conv(x, off, msk, wt, None, 1, 1, 1, 1, 1), flush, reps=3)
        y = layers.modulated_deform_conv(x, off, msk, wt, None, 1, 1, 1, 1, 1)
        go = torch.randn_like(y)
        tb = timed(lambda: torch.autograd.grad(y, (x, off, msk, wt), go, retain_graph=True), flush, reps=3)
        out["ops"].append({"op": "modulated_deform_conv 3x3 (fp32, SIMT GEMM)", "shape": [2, c, h, w], "fwd_us": round(tf * 1e6, 1),
                           "fwd_tflops": round(flops / tf / 1e12, 2), "bwd_us": round(tb * 1e6, 1),
                           "bwd_tflops": round(2 * flops / tb / 1e12, 2)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
