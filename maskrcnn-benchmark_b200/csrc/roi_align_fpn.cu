// roi_align_fpn.cu -- the whole FPN `Pooler.forward` in one launch (sm_100a).
//
// Replaces, for NHWC feature maps, the reference's multi-level pooling loop (modeling/poolers.py:91-121):
// LevelMapper (poolers.py:11-42) + per-level `nonzero` (a host sync each) + per-level ROIAlign launch
// (csrc/cuda/ROIAlign_cuda.cu:64-122) + index scatter into a zero-filled result.  Here every ROI picks
// its level inside the kernel and writes its row of the result directly.
//   * input : up to 5 levels, logical [N,C,H_l,W_l] stored NHWC, bf16 or fp32
//   * output: [R,C,P,P] in NCHW order (what the box head's fc6 expects, flattened) or [R,P,P,C] (NHWC, what
//     the mask head's conv engine consumes), bf16 or fp32
//   * one CTA per (ROI, 128-channel slab); warp per bin; lane owns 4 consecutive channels -> every
//     bilinear tap is one fully coalesced 256 B (bf16) / 512 B (fp32) request.
// Sampling arithmetic is the reference's (see roi_align.cu) in fp32 with the same operation order.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace mrb {

constexpr int kMaxLevels = 5;
constexpr int kFpnThreads = 256;
constexpr int kFpnSlab = 128;
constexpr int kMaxP = 32;                     // pooled sizes up to 32 use the per-CTA axis tables

struct FpnArgs {
  const void* feat[kMaxLevels];
  float* grad[kMaxLevels];
  int H[kMaxLevels], W[kMaxLevels];
  float scale[kMaxLevels];
  int num_levels, C, P, sampling_ratio, k_min, k_max, lvl0;
  float s0, eps;
  int out_nhwc;
};

struct RoiGeomF {
  int b, level;
  float sw, sh, bin_h, bin_w;
  int gh, gw;
  float count;
};

__device__ __forceinline__ RoiGeomF fpn_geom(const float* __restrict__ roi, const FpnArgs& a) {
  RoiGeomF g;
  g.b = (int)roi[0];
  // LevelMapper (poolers.py:31-42): area with the +1 convention, floor(lvl0 + log2(sqrt(area)/s0 + eps))
  const float area = (roi[3] - roi[1] + 1.f) * (roi[4] - roi[2] + 1.f);
  float lv = floorf((float)a.lvl0 + log2f(sqrtf(area) / a.s0 + a.eps));
  lv = fminf(fmaxf(lv, (float)a.k_min), (float)a.k_max);
  g.level = min(max((int)lv - a.k_min, 0), a.num_levels - 1);
  const float scale = a.scale[g.level];
  g.sw = __fmul_rn(roi[1], scale);
  g.sh = __fmul_rn(roi[2], scale);
  const float ew = __fmul_rn(roi[3], scale), eh = __fmul_rn(roi[4], scale);
  const float rw = fmaxf(__fsub_rn(ew, g.sw), 1.f), rh = fmaxf(__fsub_rn(eh, g.sh), 1.f);
  g.bin_h = __fdiv_rn(rh, (float)a.P);
  g.bin_w = __fdiv_rn(rw, (float)a.P);
  g.gh = (a.sampling_ratio > 0) ? a.sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)a.P));
  g.gw = (a.sampling_ratio > 0) ? a.sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)a.P));
  g.count = (float)(g.gh * g.gw);
  return g;
}

__device__ __forceinline__ float fpn_coord(float start, int p, float bin, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid));
}

struct Tap {
  int yl, xl, yh, xh;
  float w1, w2, w3, w4;
  bool valid;
};

__device__ __forceinline__ Tap fpn_tap(int H, int W, float y, float x) {
  Tap s;
  s.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (!s.valid) return s;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
  const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
  const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
  s.w1 = __fmul_rn(hy, hx); s.w2 = __fmul_rn(hy, lx); s.w3 = __fmul_rn(ly, hx); s.w4 = __fmul_rn(ly, lx);
  s.yl = yl; s.xl = xl; s.yh = yh; s.xh = xh;
  return s;
}

// One axis of a bin with two samples: (row, weight) pairs of both samples with coinciding rows merged (weight 0 =
// unused entry).  Same clamping / validity rules as fpn_tap.
struct Axis2 {
  int r[4];
  float w[4];
};
__device__ __forceinline__ void axis2(int L, float c0, float c1, Axis2& a) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float c = i ? c1 : c0;
    const bool valid = !(c < -1.0f || c > (float)L);
    if (c <= 0) c = 0;
    int lo = (int)c, hi;
    if (lo >= L - 1) { hi = lo = L - 1; c = (float)lo; } else { hi = lo + 1; }
    float l = __fsub_rn(c, (float)lo), h = __fsub_rn(1.f, l);
    if (!valid) { l = 0.f; h = 0.f; lo = 0; hi = 0; }
    a.r[2 * i] = lo; a.w[2 * i] = h;
    a.r[2 * i + 1] = hi; a.w[2 * i + 1] = l;
  }
  if (a.r[2] == a.r[0]) { a.w[0] += a.w[2]; a.w[2] = 0.f; }
  else if (a.r[2] == a.r[1]) { a.w[1] += a.w[2]; a.w[2] = 0.f; }
  if (a.r[3] == a.r[1]) { a.w[1] += a.w[3]; a.w[3] = 0.f; }
  else if (a.r[3] == a.r[0]) { a.w[0] += a.w[3]; a.w[3] = 0.f; }
}

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(kFpnThreads)
roi_align_fpn_fwd_kernel(FpnArgs a, const float* __restrict__ rois, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = blockIdx.x, c0 = blockIdx.y * kFpnSlab;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RoiGeomF g = fpn_geom(rois + (size_t)n * 5, a);
  const int H = a.H[g.level], W = a.W[g.level], C = a.C, PP = a.P * a.P, PPS = PP | 1;
  const int cl = c0 + lane * 4;
  const bool c_ok = cl < C;
  const T* __restrict__ src = reinterpret_cast<const T*>(a.feat[g.level]) + (size_t)g.b * H * W * C + (c_ok ? cl : 0);
  float* tile = reinterpret_cast<float*>(smem_raw);  // [slab][PPS], NCHW-order output only
  // Separable sampling geometry once per CTA: the (rows, weights) of bin row ph and of bin column pw -- 2 P axis evaluations
  // instead of one pair per bin in every lane (ncu on R = 1024, P = 7: the kernel was issue bound, SM throughput 65 % with 7 %
  // of the L2 bandwidth in use).
  __shared__ Axis2 s_ay[kMaxP], s_ax[kMaxP];
  const bool sep = sizeof(T) == 2 && g.gh == 2 && g.gw == 2 && a.P <= kMaxP;
  if (sep) {
    if ((int)threadIdx.x < a.P)
      axis2(H, fpn_coord(g.sh, threadIdx.x, g.bin_h, 0, 2), fpn_coord(g.sh, threadIdx.x, g.bin_h, 1, 2), s_ay[threadIdx.x]);
    else if ((int)threadIdx.x < 2 * a.P)
      axis2(W, fpn_coord(g.sw, threadIdx.x - a.P, g.bin_w, 0, 2), fpn_coord(g.sw, threadIdx.x - a.P, g.bin_w, 1, 2), s_ax[threadIdx.x - a.P]);
    __syncthreads();
  }
  for (int bin = warp; bin < PP; bin += kFpnThreads / 32) {
    const int ph = bin / a.P, pw = bin - ph * a.P;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (sep) {
      // bf16 features, sampling_ratio 2: the separable merged-weight form of the backward kernel -- (distinct rows) x
      // (distinct columns) loads per bin, typically 4-9 instead of 16.  The fp32 path below keeps the reference's
      // per-sample summation order (bit-exact); bf16 results are rounded to 8 bits anyway.
      const Axis2 ay = s_ay[ph], ax = s_ax[pw];
      // All (row, column) taps of the bin are requested before the first one is consumed: the gather is latency bound
      // (a load -> fma chain per tap kept ONE 256-byte request in flight per warp), so memory-level parallelism is what
      // buys bandwidth.  Zero-weight taps are predicated off (no traffic).
      uint2 raw[16];
      float wg[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wgt = ay.w[i] * ax.w[j];
          wg[i * 4 + j] = wgt;
          raw[i * 4 + j] = make_uint2(0u, 0u);
          if (c_ok && wgt != 0.f)
            raw[i * 4 + j] = __ldg(reinterpret_cast<const uint2*>(src + ((size_t)ay.r[i] * W + ax.r[j]) * C));
        }
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[t]);
        const float2 lo = __bfloat1622float2(h2[0]), hi = __bfloat1622float2(h2[1]);
        acc[0] = fmaf(wg[t], lo.x, acc[0]); acc[1] = fmaf(wg[t], lo.y, acc[1]);
        acc[2] = fmaf(wg[t], hi.x, acc[2]); acc[3] = fmaf(wg[t], hi.y, acc[3]);
      }
    } else
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = fpn_coord(g.sh, ph, g.bin_h, iy, g.gh);
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = fpn_coord(g.sw, pw, g.bin_w, ix, g.gw);
        const Tap t = fpn_tap(H, W, y, x);
        if (!t.valid || !c_ok) continue;
        float v1[4], v2[4], v3[4], v4[4];
        load4<T>(src + ((size_t)t.yl * W + t.xl) * C, v1);
        load4<T>(src + ((size_t)t.yl * W + t.xh) * C, v2);
        load4<T>(src + ((size_t)t.yh * W + t.xl) * C, v3);
        load4<T>(src + ((size_t)t.yh * W + t.xh) * C, v4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          acc[k] = __fadd_rn(acc[k], __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w1, v1[k]), __fmul_rn(t.w2, v2[k])),
                                                         __fmul_rn(t.w3, v3[k])), __fmul_rn(t.w4, v4[k])));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = __fdiv_rn(acc[k], g.count);
    if (a.out_nhwc) {
      if (c_ok) store4<T>(out + ((size_t)n * PP + bin) * C + cl, acc);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) tile[(lane * 4 + k) * PPS + bin] = acc[k];
    }
  }
  if (!a.out_nhwc) {
    __syncthreads();
    const int cn = min(kFpnSlab, C - c0);
    T* __restrict__ dst = out + ((size_t)n * C + c0) * PP;
    for (int o = threadIdx.x; o < cn * PP; o += kFpnThreads) {
      const int c = o / PP, bin = o - c * PP;
      dst[o] = from_f<T>(tile[c * PPS + bin]);
    }
  }
}

// ---- forward, second form (bf16 features, sampling_ratio 2, C % 8 == 0): one CTA per ROI, a lane owns 8 consecutive channels
// (16-byte loads: a warp request covers 256 channels = 512 B), the taps of a bin go out one row (4 columns) at a time and rows
// of zero weight are skipped, which keeps the kernel at 4 CTAs per SM (the first form holds all 16 taps of 8 B + 16 weights
// per lane: 80 registers, 3 CTAs per SM, twice the load instructions per byte).  Same taps and weights as the first form's
// separable path; a zero-weight tap contributes an exact +0, so skipping it does not change the sum.
constexpr int kV2Lane = 8;
__global__ void __launch_bounds__(kFpnThreads, 4)
roi_align_fpn_fwd_v2_kernel(FpnArgs a, const float* __restrict__ rois, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RoiGeomF g = fpn_geom(rois + (size_t)n * 5, a);
  const int H = a.H[g.level], W = a.W[g.level], C = a.C, PP = a.P * a.P, PPS = PP | 1;
  // gridDim.y CTAs share the bins of a ROI (few ROIs, many bins: the mask head's 256 x 14 x 14)
  const int bins_per = (PP + (int)gridDim.y - 1) / (int)gridDim.y;
  const int bin_lo = (int)blockIdx.y * bins_per, bin_hi = min(bin_lo + bins_per, PP);
  float* tile = reinterpret_cast<float*>(smem_raw);  // [256][PPS], NCHW-order output only
  __shared__ Axis2 s_ay[kMaxP], s_ax[kMaxP];
  if ((int)threadIdx.x < a.P)
    axis2(H, fpn_coord(g.sh, threadIdx.x, g.bin_h, 0, 2), fpn_coord(g.sh, threadIdx.x, g.bin_h, 1, 2), s_ay[threadIdx.x]);
  else if ((int)threadIdx.x < 2 * a.P)
    axis2(W, fpn_coord(g.sw, threadIdx.x - a.P, g.bin_w, 0, 2), fpn_coord(g.sw, threadIdx.x - a.P, g.bin_w, 1, 2), s_ax[threadIdx.x - a.P]);
  __syncthreads();
  const __nv_bfloat16* __restrict__ base = reinterpret_cast<const __nv_bfloat16*>(a.feat[g.level]) + (size_t)g.b * H * W * C;
  for (int c0 = 0; c0 < C; c0 += 32 * kV2Lane) {
    const int cl = c0 + lane * kV2Lane;
    const bool c_ok = cl < C;
    const __nv_bfloat16* __restrict__ src = base + (c_ok ? cl : 0);
    for (int bin = bin_lo + warp; bin < bin_hi; bin += kFpnThreads / 32) {
      const int ph = bin / a.P, pw = bin - ph * a.P;
      const Axis2 ay = s_ay[ph], ax = s_ax[pw];
      float acc[kV2Lane];
#pragma unroll
      for (int k = 0; k < kV2Lane; ++k) acc[k] = 0.f;
      const size_t co0 = (size_t)ax.r[0] * C, co1 = (size_t)ax.r[1] * C, co2 = (size_t)ax.r[2] * C, co3 = (size_t)ax.r[3] * C;
#pragma unroll
      for (int i = 0; i < 4; ++i) {          // one row of taps (4 columns) in flight per batch
        if (ay.w[i] == 0.f) continue;        // warp-uniform
        const __nv_bfloat16* rp = src + (size_t)ay.r[i] * W * C;
        uint4 raw[4];
        float wg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wgt = ay.w[i] * ax.w[j];
          wg[j] = wgt;
          raw[j] = make_uint4(0u, 0u, 0u, 0u);
          if (c_ok && wgt != 0.f)
            raw[j] = __ldg(reinterpret_cast<const uint4*>(rp + (j == 0 ? co0 : (j == 1 ? co1 : (j == 2 ? co2 : co3)))));
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[t]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(h2[q]);
            acc[2 * q] = fmaf(wg[t], f.x, acc[2 * q]);
            acc[2 * q + 1] = fmaf(wg[t], f.y, acc[2 * q + 1]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kV2Lane; ++k) acc[k] = __fdiv_rn(acc[k], g.count);
      if (a.out_nhwc) {
        if (c_ok) {
          uint4 o;
          __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int q = 0; q < 4; ++q) oh[q] = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
          *reinterpret_cast<uint4*>(out + ((size_t)n * PP + bin) * C + cl) = o;
        }
      } else {
#pragma unroll
        for (int k = 0; k < kV2Lane; ++k) tile[(lane * kV2Lane + k) * PPS + bin] = acc[k];
      }
    }
    if (!a.out_nhwc) {
      __syncthreads();
      const int cn = min(32 * kV2Lane, C - c0);
      __nv_bfloat16* __restrict__ dst = out + ((size_t)n * C + c0) * PP;
      for (int c = warp; c < cn; c += kFpnThreads / 32)
        for (int bin = bin_lo + lane; bin < bin_hi; bin += 32) dst[(size_t)c * PP + bin] = __float2bfloat16_rn(tile[c * PPS + bin]);
      __syncthreads();
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kFpnThreads)
roi_align_fpn_bwd_kernel(FpnArgs a, const float* __restrict__ rois, const T* __restrict__ gout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = blockIdx.x, c0 = blockIdx.y * kFpnSlab;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RoiGeomF g = fpn_geom(rois + (size_t)n * 5, a);
  const int H = a.H[g.level], W = a.W[g.level], C = a.C, PP = a.P * a.P, PPS = PP | 1;
  float* tile = reinterpret_cast<float*>(smem_raw);
  const int cn = min(kFpnSlab, C - c0);
  if (!a.out_nhwc) {
    const T* __restrict__ top = gout + ((size_t)n * C + c0) * PP;
    for (int o = threadIdx.x; o < cn * PP; o += kFpnThreads) {
      const int c = o / PP, bin = o - c * PP;
      tile[c * PPS + bin] = to_f<T>(top[o]);
    }
    __syncthreads();
  }
  __shared__ Axis2 s_ay[kMaxP], s_ax[kMaxP];
  const bool sep = g.gh == 2 && g.gw == 2 && a.P <= kMaxP;
  if (sep) {
    if ((int)threadIdx.x < a.P)
      axis2(H, fpn_coord(g.sh, threadIdx.x, g.bin_h, 0, 2), fpn_coord(g.sh, threadIdx.x, g.bin_h, 1, 2), s_ay[threadIdx.x]);
    else if ((int)threadIdx.x < 2 * a.P)
      axis2(W, fpn_coord(g.sw, threadIdx.x - a.P, g.bin_w, 0, 2), fpn_coord(g.sw, threadIdx.x - a.P, g.bin_w, 1, 2), s_ax[threadIdx.x - a.P]);
    __syncthreads();
  }
  const int cl = c0 + lane * 4;
  if (cl >= C) return;
  float* __restrict__ dst = a.grad[g.level] + (size_t)g.b * H * W * C + cl;
  for (int bin = warp; bin < PP; bin += kFpnThreads / 32) {
    const int ph = bin / a.P, pw = bin - ph * a.P;
    float t4[4];
    if (a.out_nhwc) {
      load4<T>(gout + ((size_t)n * PP + bin) * C + cl, t4);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) t4[k] = tile[(lane * 4 + k) * PPS + bin];
    }
    if (sep) {
      // sampling_ratio 2 (every reference config).  Bilinear weights and the validity test are separable, so the
      // 2 x 2 samples x 4 corners of a bin collapse to (distinct rows) x (distinct columns) of summed weights:
      // typically 2-3 x 2-3 red.adds per bin instead of 16 (samples are half a bin apart, bins ~1-2 pixels wide).
      const Axis2 ay = s_ay[ph], ax = s_ax[pw];
      const float inv = __fdiv_rn(1.f, g.count);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (ay.w[i] == 0.f) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wgt = ay.w[i] * ax.w[j] * inv;
          if (wgt == 0.f) continue;
          atomicAdd(reinterpret_cast<float4*>(dst + ((size_t)ay.r[i] * W + ax.r[j]) * C),
                    make_float4(t4[0] * wgt, t4[1] * wgt, t4[2] * wgt, t4[3] * wgt));
        }
      }
      continue;
    }
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = fpn_coord(g.sh, ph, g.bin_h, iy, g.gh);
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = fpn_coord(g.sw, pw, g.bin_w, ix, g.gw);
        const Tap t = fpn_tap(H, W, y, x);
        if (!t.valid) continue;
        float4 g1, g2, g3, g4;
        float* f1 = &g1.x; float* f2 = &g2.x; float* f3 = &g3.x; float* f4 = &g4.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          f1[k] = __fdiv_rn(__fmul_rn(t4[k], t.w1), g.count);
          f2[k] = __fdiv_rn(__fmul_rn(t4[k], t.w2), g.count);
          f3[k] = __fdiv_rn(__fmul_rn(t4[k], t.w3), g.count);
          f4[k] = __fdiv_rn(__fmul_rn(t4[k], t.w4), g.count);
        }
        atomicAdd(reinterpret_cast<float4*>(dst + ((size_t)t.yl * W + t.xl) * C), g1);
        atomicAdd(reinterpret_cast<float4*>(dst + ((size_t)t.yl * W + t.xh) * C), g2);
        atomicAdd(reinterpret_cast<float4*>(dst + ((size_t)t.yh * W + t.xl) * C), g3);
        atomicAdd(reinterpret_cast<float4*>(dst + ((size_t)t.yh * W + t.xh) * C), g4);
      }
    }
  }
}

static int fpn_fill(FpnArgs& a, const void* const* feats, float* const* grads, const int* heights, const int* widths,
                    const float* scales, int num_levels, int channels, int pooled, int sampling_ratio, int k_min, int k_max,
                    float canonical_scale, int canonical_level, int out_nhwc) {
  if (num_levels <= 0 || num_levels > kMaxLevels || channels <= 0 || channels % 4 || pooled <= 0) return MRB_ERR_BAD_ARG;
  if (k_max - k_min + 1 != num_levels) return MRB_ERR_BAD_ARG;
  for (int l = 0; l < num_levels; ++l) {
    a.feat[l] = feats ? feats[l] : nullptr;
    a.grad[l] = grads ? grads[l] : nullptr;
    a.H[l] = heights[l]; a.W[l] = widths[l]; a.scale[l] = scales[l];
    if (heights[l] <= 0 || widths[l] <= 0) return MRB_ERR_BAD_ARG;
    if (feats && (!feats[l] || ((uintptr_t)feats[l] & 15))) return MRB_ERR_BAD_ARG;
    if (grads && (!grads[l] || ((uintptr_t)grads[l] & 15))) return MRB_ERR_BAD_ARG;
  }
  a.num_levels = num_levels; a.C = channels; a.P = pooled; a.sampling_ratio = sampling_ratio;
  a.k_min = k_min; a.k_max = k_max; a.lvl0 = canonical_level; a.s0 = canonical_scale; a.eps = 1e-6f;
  a.out_nhwc = out_nhwc;
  return MRB_OK;
}

}  // namespace mrb
using namespace mrb;

MRB_API int mrb_roi_align_fpn_fwd(const void* const* feats_host, const int* heights_host, const int* widths_host,
                                  const float* scales_host, int num_levels, const float* rois, void* output, int num_rois,
                                  int batch, int channels, int pooled, int sampling_ratio, int k_min, int k_max,
                                  float canonical_scale, int canonical_level, int dtype, int out_nhwc, mrb_stream_t stream) {
  if (num_rois < 0 || !feats_host || !heights_host || !widths_host || !scales_host) return MRB_ERR_BAD_ARG;
  FpnArgs a;
  int rc = fpn_fill(a, feats_host, nullptr, heights_host, widths_host, scales_host, num_levels, channels, pooled, sampling_ratio,
                    k_min, k_max, canonical_scale, canonical_level, out_nhwc);
  if (rc) return rc;
  if (num_rois == 0) return MRB_OK;
  if (!rois || !output) return MRB_ERR_BAD_ARG;
  const size_t smem = out_nhwc ? 0 : (size_t)kFpnSlab * ((pooled * pooled) | 1) * sizeof(float);
  if (smem > 200 * 1024) return MRB_ERR_UNSUPPORTED;
  dim3 grid(num_rois, ceil_div(channels, kFpnSlab));
  static const bool no_v2 = [] { const char* e = getenv("MRB_ROIALIGN_FPN_V2"); return e && e[0] == '0'; }();
  const size_t smem2 = out_nhwc ? 0 : (size_t)(32 * kV2Lane) * ((pooled * pooled) | 1) * sizeof(float);
  if (dtype == MRB_BF16 && !no_v2 && sampling_ratio == 2 && channels % kV2Lane == 0 && pooled <= kMaxP && smem2 <= 72 * 1024) {
    // (adaptive sampling grids, other channel counts and large NCHW tiles take the first form)
    if (smem2 > 48 * 1024) MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fpn_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    // at least ~4 CTAs per SM in total: split a ROI's bins over up to 8 CTAs when there are few ROIs
    int split = (4 * kNumSMs + num_rois - 1) / num_rois;
    split = split < 1 ? 1 : (split > 8 ? 8 : split);
    if (split > pooled * pooled) split = pooled * pooled;
    roi_align_fpn_fwd_v2_kernel<<<dim3(num_rois, split), kFpnThreads, smem2, (cudaStream_t)stream>>>(a, rois, (__nv_bfloat16*)output);
  } else if (dtype == MRB_BF16) {
    if (smem > 48 * 1024) MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fpn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_fpn_fwd_kernel<__nv_bfloat16><<<grid, kFpnThreads, smem, (cudaStream_t)stream>>>(a, rois, (__nv_bfloat16*)output);
  } else if (dtype == MRB_F32) {
    if (smem > 48 * 1024) MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fpn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_fpn_fwd_kernel<float><<<grid, kFpnThreads, smem, (cudaStream_t)stream>>>(a, rois, (float*)output);
  } else {
    return MRB_ERR_BAD_ARG;
  }
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_roi_align_fpn_bwd(const void* grad_output, float* const* grad_feats_host, const int* heights_host,
                                  const int* widths_host, const float* scales_host, int num_levels, const float* rois,
                                  int num_rois, int batch, int channels, int pooled, int sampling_ratio, int k_min, int k_max,
                                  float canonical_scale, int canonical_level, int dtype, int out_nhwc, mrb_stream_t stream) {
  if (num_rois < 0 || !grad_feats_host || !heights_host || !widths_host || !scales_host) return MRB_ERR_BAD_ARG;
  FpnArgs a;
  int rc = fpn_fill(a, nullptr, grad_feats_host, heights_host, widths_host, scales_host, num_levels, channels, pooled,
                    sampling_ratio, k_min, k_max, canonical_scale, canonical_level, out_nhwc);
  if (rc) return rc;
  if (num_rois == 0) return MRB_OK;
  if (!rois || !grad_output) return MRB_ERR_BAD_ARG;
  const size_t smem = out_nhwc ? 0 : (size_t)kFpnSlab * ((pooled * pooled) | 1) * sizeof(float);
  if (smem > 200 * 1024) return MRB_ERR_UNSUPPORTED;
  dim3 grid(num_rois, ceil_div(channels, kFpnSlab));
  if (dtype == MRB_BF16) {
    if (smem > 48 * 1024) MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fpn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_fpn_bwd_kernel<__nv_bfloat16><<<grid, kFpnThreads, smem, (cudaStream_t)stream>>>(a, rois, (const __nv_bfloat16*)grad_output);
  } else if (dtype == MRB_F32) {
    if (smem > 48 * 1024) MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fpn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_fpn_bwd_kernel<float><<<grid, kFpnThreads, smem, (cudaStream_t)stream>>>(a, rois, (const float*)grad_output);
  } else {
    return MRB_ERR_BAD_ARG;
  }
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
