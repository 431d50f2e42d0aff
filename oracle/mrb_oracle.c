/*
 * mrb_oracle.c -- CPU restatement of the maskrcnn-benchmark csrc hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
 * it, and only as the checker.  The product (libmrb_b200.so) never links or
 * calls anything in this directory and has no CPU fallback.
 *
 * Every function restates, in plain scalar C, the algorithm of the reference
 * file:line it cites (paths relative to /root/reference/maskrcnn_benchmark).
 * Arithmetic is fp32, evaluated in the reference's operation order, and this
 * file must be compiled with -ffp-contract=off so that gcc never fuses a
 * multiply-add the reference's x86-64 build (no FMA) performs as two ops.
 *
 * Pinning (see tests/test_oracle.py):
 *   - orc_nms          vs reference tests/test_nms.py:11-58,60-217 golden vectors
 *                      and vs oracle/_ref (the reference's own nms_cpu.cpp compiled here)
 *   - orc_roi_align_fwd vs oracle/_ref (reference ROIAlign_cpu.cpp) bit-exact
 *   - roi_align_bwd / roi_pool / focal / deform_*: the reference has no CPU
 *     implementation and no numeric test ("parity unpinned" by the reference's
 *     own tests); they are pinned against torchvision CPU ops (roi_pool,
 *     roi_align autograd, deform_conv2d) and the reference's python
 *     sigmoid_focal_loss_cpu via committed fixtures in tests/golden/.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ NMS -- */
/* csrc/cpu/nms_cpu.cpp:5-65.  dets [n,4] xyxy, scores [n].  `order` is the
 * descending-score permutation (the caller supplies it: the reference uses
 * scores.sort(0, descending=True) -- ties resolved by the stable sort, i.e.
 * ascending index; pass order=NULL to have it computed that way here).
 * Writes kept ORIGINAL indices ascending (at::nonzero(suppressed==0),
 * nms_cpu.cpp:64) into keep[]; returns the count. */
typedef struct { float s; int64_t i; } orc_si;
static int orc_cmp_desc(const void* a, const void* b) {
  const orc_si* x = (const orc_si*)a; const orc_si* y = (const orc_si*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i < y->i) ? -1 : (x->i > y->i);
}

ORC_API int64_t orc_nms(const float* dets, const float* scores, int64_t n,
                        float threshold, const int64_t* order_in, int64_t* keep) {
  if (n == 0) return 0;
  float* areas = (float*)malloc(sizeof(float) * n);
  uint8_t* suppressed = (uint8_t*)calloc(n, 1);
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * n);
  if (order_in) {
    memcpy(order, order_in, sizeof(int64_t) * n);
  } else {
    orc_si* si = (orc_si*)malloc(sizeof(orc_si) * n);
    for (int64_t i = 0; i < n; i++) { si[i].s = scores[i]; si[i].i = i; }
    qsort(si, n, sizeof(orc_si), orc_cmp_desc);
    for (int64_t i = 0; i < n; i++) order[i] = si[i].i;
    free(si);
  }
  /* nms_cpu.cpp:22  areas = (x2 - x1 + 1) * (y2 - y1 + 1) */
  for (int64_t i = 0; i < n; i++) {
    float w = dets[i * 4 + 2] - dets[i * 4 + 0]; w = w + 1.0f;
    float h = dets[i * 4 + 3] - dets[i * 4 + 1]; h = h + 1.0f;
    areas[i] = w * h;
  }
  /* nms_cpu.cpp:37-63 */
  for (int64_t _i = 0; _i < n; _i++) {
    int64_t i = order[_i];
    if (suppressed[i] == 1) continue;
    float ix1 = dets[i * 4 + 0], iy1 = dets[i * 4 + 1];
    float ix2 = dets[i * 4 + 2], iy2 = dets[i * 4 + 3];
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; _j++) {
      int64_t j = order[_j];
      if (suppressed[j] == 1) continue;
      float xx1 = fmaxf(ix1, dets[j * 4 + 0]);
      float yy1 = fmaxf(iy1, dets[j * 4 + 1]);
      float xx2 = fminf(ix2, dets[j * 4 + 2]);
      float yy2 = fminf(iy2, dets[j * 4 + 3]);
      float w = xx2 - xx1; w = w + 1.0f; w = fmaxf(0.0f, w);
      float h = yy2 - yy1; h = h + 1.0f; h = fmaxf(0.0f, h);
      float inter = w * h;
      float den = iarea + areas[j]; den = den - inter;
      float ovr = inter / den;
      if (ovr >= threshold) suppressed[j] = 1;   /* nms_cpu.cpp:60  '>=' */
    }
  }
  int64_t k = 0;
  for (int64_t i = 0; i < n; i++) if (!suppressed[i]) keep[k++] = i;
  free(areas); free(suppressed); free(order);
  return k;
}

/* ------------------------------------------------------------ ROIAlign -- */
/* Shared sample-point rule: csrc/cpu/ROIAlign_cpu.cpp:45-98 ==
 * csrc/cuda/ROIAlign_cuda.cu:15-62 (bilinear_interpolate) and :125-174
 * (bilinear_interpolate_gradient). Returns 0 if the sample is outside. */
static inline int orc_ra_sample(int height, int width, float y, float x,
                                int* y_low, int* x_low, int* y_high, int* x_high,
                                float* w1, float* w2, float* w3, float* w4) {
  if (y < -1.0 || y > height || x < -1.0 || x > width) return 0;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= height - 1) { yh = yl = height - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= width - 1) { xh = xl = width - 1; x = (float)xl; } else { xh = xl + 1; }
  float ly = y - yl, lx = x - xl;
  float hy = (float)(1. - ly), hx = (float)(1. - lx);
  *w1 = hy * hx; *w2 = hy * lx; *w3 = ly * hx; *w4 = ly * lx;
  *y_low = yl; *x_low = xl; *y_high = yh; *x_high = xh;
  return 1;
}

/* ROI geometry: ROIAlign_cpu.cpp:145-170 / ROIAlign_cuda.cu:77-104 */
typedef struct { int b; float sw, sh, bin_h, bin_w; int gh, gw; float count; } orc_roi_geom;
static inline orc_roi_geom orc_ra_geom(const float* roi, float scale, int ph, int pw, int sampling_ratio) {
  orc_roi_geom g;
  g.b = (int)roi[0];
  g.sw = roi[1] * scale; g.sh = roi[2] * scale;
  float ew = roi[3] * scale, eh = roi[4] * scale;
  float rw = fmaxf(ew - g.sw, 1.0f), rh = fmaxf(eh - g.sh, 1.0f);
  g.bin_h = rh / (float)ph; g.bin_w = rw / (float)pw;
  g.gh = (sampling_ratio > 0) ? sampling_ratio : (int)ceilf(rh / ph);
  g.gw = (sampling_ratio > 0) ? sampling_ratio : (int)ceilf(rw / pw);
  g.count = (float)(g.gh * g.gw);
  return g;
}

/* csrc/cpu/ROIAlign_cpu.cpp:113-219 (same result as ROIAlign_cuda.cu:64-122).
 * input [N,C,H,W] NCHW fp32, rois [R,5], output [R,C,PH,PW]. */
ORC_API void orc_roi_align_fwd(const float* input, const float* rois, int R, int C, int H, int W,
                               int PH, int PW, float scale, int sampling_ratio, float* out) {
  for (int n = 0; n < R; n++) {
    orc_roi_geom g = orc_ra_geom(rois + n * 5, scale, PH, PW, sampling_ratio);
    for (int c = 0; c < C; c++) {
      const float* src = input + ((int64_t)g.b * C + c) * H * W;
      for (int ph = 0; ph < PH; ph++) for (int pw = 0; pw < PW; pw++) {
        float acc = 0.f;
        for (int iy = 0; iy < g.gh; iy++) {
          const float yy = g.sh + ph * g.bin_h + (float)(iy + .5f) * g.bin_h / (float)g.gh;
          for (int ix = 0; ix < g.gw; ix++) {
            const float xx = g.sw + pw * g.bin_w + (float)(ix + .5f) * g.bin_w / (float)g.gw;
            int yl, xl, yh, xh; float w1, w2, w3, w4;
            if (!orc_ra_sample(H, W, yy, xx, &yl, &xl, &yh, &xh, &w1, &w2, &w3, &w4)) continue;
            /* ROIAlign_cpu.cpp:199-202: left-to-right sum of the four products, then += */
            float v = w1 * src[yl * W + xl];
            v = v + w2 * src[yl * W + xh];
            v = v + w3 * src[yh * W + xl];
            v = v + w4 * src[yh * W + xh];
            acc += v;
          }
        }
        acc /= g.count;
        out[(((int64_t)n * C + c) * PH + ph) * PW + pw] = acc;
      }
    }
  }
}

/* csrc/cuda/ROIAlign_cuda.cu:177-254 (no CPU implementation in the reference:
 * csrc/ROIAlign.h:44).  grad [R,C,PH,PW] -> grad_input [N,C,H,W] (zeroed here,
 * ROIAlign_cuda.cu:316).  Summation order = (n, c, ph, pw, iy, ix) serial; the
 * CUDA reference's atomicAdd order is unspecified. */
ORC_API void orc_roi_align_bwd(const float* grad, const float* rois, int R, int N, int C, int H, int W,
                               int PH, int PW, float scale, int sampling_ratio, float* gin) {
  memset(gin, 0, sizeof(float) * (size_t)N * C * H * W);
  for (int n = 0; n < R; n++) {
    orc_roi_geom g = orc_ra_geom(rois + n * 5, scale, PH, PW, sampling_ratio);
    for (int c = 0; c < C; c++) {
      float* dst = gin + ((int64_t)g.b * C + c) * H * W;
      for (int ph = 0; ph < PH; ph++) for (int pw = 0; pw < PW; pw++) {
        const float top = grad[(((int64_t)n * C + c) * PH + ph) * PW + pw];
        for (int iy = 0; iy < g.gh; iy++) {
          const float yy = g.sh + ph * g.bin_h + (float)(iy + .5f) * g.bin_h / (float)g.gh;
          for (int ix = 0; ix < g.gw; ix++) {
            const float xx = g.sw + pw * g.bin_w + (float)(ix + .5f) * g.bin_w / (float)g.gw;
            int yl, xl, yh, xh; float w1, w2, w3, w4;
            if (!orc_ra_sample(H, W, yy, xx, &yl, &xl, &yh, &xh, &w1, &w2, &w3, &w4)) continue;
            /* ROIAlign_cuda.cu:236-239: g = top * w / count */
            dst[yl * W + xl] += top * w1 / g.count;
            dst[yl * W + xh] += top * w2 / g.count;
            dst[yh * W + xl] += top * w3 / g.count;
            dst[yh * W + xh] += top * w4 / g.count;
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------- ROIPool -- */
/* csrc/cuda/ROIPool_cuda.cu:16-77.  argmax is int32 index into the H*W plane
 * (-1 for an empty bin). */
ORC_API void orc_roi_pool_fwd(const float* input, const float* rois, int R, int C, int H, int W,
                              int PH, int PW, float scale, float* out, int32_t* argmax) {
  for (int n = 0; n < R; n++) {
    const float* roi = rois + n * 5;
    int b = (int)roi[0];
    int rsw = (int)roundf(roi[1] * scale), rsh = (int)roundf(roi[2] * scale);
    int rew = (int)roundf(roi[3] * scale), reh = (int)roundf(roi[4] * scale);
    int rw = rew - rsw + 1; if (rw < 1) rw = 1;
    int rh = reh - rsh + 1; if (rh < 1) rh = 1;
    float bin_h = (float)rh / (float)PH, bin_w = (float)rw / (float)PW;
    for (int c = 0; c < C; c++) {
      const float* src = input + ((int64_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ph++) for (int pw = 0; pw < PW; pw++) {
        int hstart = (int)floorf((float)ph * bin_h), wstart = (int)floorf((float)pw * bin_w);
        int hend = (int)ceilf((float)(ph + 1) * bin_h), wend = (int)ceilf((float)(pw + 1) * bin_w);
        hstart = hstart + rsh; if (hstart < 0) hstart = 0; if (hstart > H) hstart = H;
        hend = hend + rsh; if (hend < 0) hend = 0; if (hend > H) hend = H;
        wstart = wstart + rsw; if (wstart < 0) wstart = 0; if (wstart > W) wstart = W;
        wend = wend + rsw; if (wend < 0) wend = 0; if (wend > W) wend = W;
        int is_empty = (hend <= hstart) || (wend <= wstart);
        float maxval = is_empty ? 0.f : -FLT_MAX;
        int maxidx = -1;
        for (int h = hstart; h < hend; ++h) for (int w = wstart; w < wend; ++w) {
          int bi = h * W + w;
          if (src[bi] > maxval) { maxval = src[bi]; maxidx = bi; }
        }
        int64_t o = (((int64_t)n * C + c) * PH + ph) * PW + pw;
        out[o] = maxval; argmax[o] = maxidx;
      }
    }
  }
}

/* csrc/cuda/ROIPool_cuda.cu:79-108 */
ORC_API void orc_roi_pool_bwd(const float* grad, const float* rois, const int32_t* argmax, int R, int N,
                              int C, int H, int W, int PH, int PW, float* gin) {
  memset(gin, 0, sizeof(float) * (size_t)N * C * H * W);
  for (int n = 0; n < R; n++) {
    int b = (int)rois[n * 5];
    for (int c = 0; c < C; c++) {
      float* dst = gin + ((int64_t)b * C + c) * H * W;
      for (int p = 0; p < PH * PW; p++) {
        int64_t o = ((int64_t)n * C + c) * PH * PW + p;
        if (argmax[o] != -1) dst[argmax[o]] += grad[o];
      }
    }
  }
}

/* ---------------------------------------------------- SigmoidFocalLoss -- */
/* csrc/cuda/SigmoidFocalLoss_cuda.cu:20-58.  The reference evaluates the
 * T=float instantiation with double literals (1., 1.0 - alpha ...), i.e. the
 * intermediate expressions are promoted to double wherever a double literal
 * appears, while expf/powf/logf are fp32 calls.  Restated with the same
 * promotions. */
ORC_API void orc_sigmoid_focal_fwd(const float* logits, const int32_t* targets, int64_t A, int num_classes,
                                   float gamma, float alpha, float* losses) {
  for (int64_t i = 0; i < A * num_classes; i++) {
    int64_t n = i / num_classes; int d = (int)(i % num_classes);
    int t = targets[n];
    float c1 = (float)(t == (d + 1));
    float c2 = (float)((t >= 0) & (t != (d + 1)));
    float zn = (float)(1.0 - alpha);
    float zp = alpha;
    float x = logits[i];
    float p = (float)(1. / (1. + expf(-x)));
    float term1 = (float)(powf((float)(1. - p), gamma) * logf(fmaxf(p, FLT_MIN)));
    int ge = (x >= 0);
    float term2 = (float)(powf(p, gamma) *
        (-1. * x * ge - logf((float)(1. + expf((float)(x - 2. * x * ge))))));
    float l = 0.0f;
    l += -c1 * term1 * zp;
    l += -c2 * term2 * zn;
    losses[i] = l;
  }
}

/* csrc/cuda/SigmoidFocalLoss_cuda.cu:61-101 */
ORC_API void orc_sigmoid_focal_bwd(const float* logits, const int32_t* targets, const float* d_losses,
                                   int64_t A, int num_classes, float gamma, float alpha, float* d_logits) {
  for (int64_t i = 0; i < A * num_classes; i++) {
    int64_t n = i / num_classes; int d = (int)(i % num_classes);
    int t = targets[n];
    float c1 = (float)(t == (d + 1));
    float c2 = (float)((t >= 0) & (t != (d + 1)));
    float zn = (float)(1.0 - alpha);
    float zp = alpha;
    float x = logits[i];
    float p = (float)(1. / (1. + expf(-x)));
    float term1 = (float)(powf((float)(1. - p), gamma) *
                          (1. - p - (p * gamma * logf(fmaxf(p, FLT_MIN)))));
    int ge = (x >= 0);
    float term2 = (float)(powf(p, gamma) *
        ((-1. * x * ge - logf((float)(1. + expf((float)(x - 2. * x * ge))))) * (1. - p) * gamma - p));
    float g = 0.0f;
    g += -c1 * term1 * zp;
    g += -c2 * term2 * zn;
    d_logits[i] = g * d_losses[i];
  }
}
