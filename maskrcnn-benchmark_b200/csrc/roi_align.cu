// roi_align.cu -- ROIAlign forward/backward for sm_100a.
//
// Replaces RoIAlignForward / RoIAlignBackwardFeature (reference
// csrc/cuda/ROIAlign_cuda.cu:64-122,177-254) and ROIAlign_forward_cpu
// (csrc/cpu/ROIAlign_cpu.cpp:113-257).  Semantics are the reference's: ROI coordinates
// are not rounded, malformed ROIs are forced to 1x1, sampling grid = sampling_ratio or
// ceil(roi/pooled), samples outside [-1, H] x [-1, W] contribute zero.
//
// Design (HBM-bound op):
//   * one CTA owns one ROI x one channel slab, so the sample table (positions + bilinear
//     weights, shared by all channels -- the Caffe2 "pre_calc" idea of ROIAlign_cpu.cpp:17-111)
//     is computed once per CTA into shared memory, laid out [sample][bin] so that a warp
//     reading consecutive bins is bank-conflict free;
//   * NCHW: lanes run over the (ph,pw) bins of one channel plane: the 16 taps of a warp
//     stay inside one ROI footprint of one plane (L1-resident), stores are fully coalesced;
//   * NHWC (torch.channels_last): lanes run over channels with 128-bit loads: every tap is
//     one fully coalesced 512 B request; the [C_slab][P*P] result tile is transposed through
//     shared memory so the NCHW-ordered output is written with coalesced 128 B lines;
//   * forward arithmetic uses explicit round-to-nearest mul/add/div (no FMA contraction) in
//     the reference's operation order => bit-identical to the reference CPU kernel.
//   * backward accumulates with red.global.add (vector red.v4.f32 for NHWC).
#include "common.cuh"

namespace mrb {

struct RoiGeom {
  int b;
  float sw, sh, bin_h, bin_w;
  int gh, gw;
  float count;
};

// ROIAlign_cuda.cu:77-104 / ROIAlign_cpu.cpp:145-170
__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int PH, int PW,
                                            int sampling_ratio) {
  RoiGeom g;
  g.b = (int)roi[0];
  g.sw = __fmul_rn(roi[1], scale);
  g.sh = __fmul_rn(roi[2], scale);
  const float ew = __fmul_rn(roi[3], scale), eh = __fmul_rn(roi[4], scale);
  const float rw = fmaxf(__fsub_rn(ew, g.sw), 1.f), rh = fmaxf(__fsub_rn(eh, g.sh), 1.f);
  g.bin_h = __fdiv_rn(rh, (float)PH);
  g.bin_w = __fdiv_rn(rw, (float)PW);
  g.gh = (sampling_ratio > 0) ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.gw = (sampling_ratio > 0) ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  g.count = (float)(g.gh * g.gw);
  return g;
}

// start + p*bin + (i+.5)*bin/grid, evaluated left to right without contraction
__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                   __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid));
}

struct Sample {
  int yl, xl, yh, xh;
  float w1, w2, w3, w4;
  bool valid;
};

// bilinear_interpolate / bilinear_interpolate_gradient, ROIAlign_cuda.cu:15-62,125-174
__device__ __forceinline__ Sample make_sample(int H, int W, float y, float x) {
  Sample s;
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
    s.valid = false;
    s.yl = s.xl = s.yh = s.xh = 0;
    s.w1 = s.w2 = s.w3 = s.w4 = 0.f;
    return s;
  }
  s.valid = true;
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

// make_sample split by axis: what depends on one coordinate only
struct AxisSample {
  int lo, hi;
  float l, h;     // l = c - lo, h = 1 - l   (after clamping)
  int valid;
};
constexpr int kRaMaxP = 32;

__device__ __forceinline__ AxisSample make_axis_sample(int L, float c) {
  AxisSample a;
  a.valid = !(c < -1.0f || c > (float)L);
  if (c <= 0) c = 0;
  int lo = (int)c, hi;
  if (lo >= L - 1) { hi = lo = L - 1; c = (float)lo; } else { hi = lo + 1; }
  a.lo = lo; a.hi = hi;
  a.l = __fsub_rn(c, (float)lo);
  a.h = __fsub_rn(1.f, a.l);
  return a;
}

__device__ __forceinline__ Sample combine_axis_samples(const AxisSample& y, const AxisSample& x) {
  Sample s;
  s.valid = y.valid && x.valid;
  if (!s.valid) {
    s.yl = s.xl = s.yh = s.xh = 0;
    s.w1 = s.w2 = s.w3 = s.w4 = 0.f;
    return s;
  }
  s.yl = y.lo; s.yh = y.hi; s.xl = x.lo; s.xh = x.hi;
  s.w1 = __fmul_rn(y.h, x.h); s.w2 = __fmul_rn(y.h, x.l); s.w3 = __fmul_rn(y.l, x.h); s.w4 = __fmul_rn(y.l, x.l);
  return s;
}

__device__ __forceinline__ float tap_sum(float w1, float v1, float w2, float v2, float w3, float v3,
                                         float w4, float v4) {
  // ROIAlign_cpu.cpp:199-202: ((w1*v1 + w2*v2) + w3*v3) + w4*v4
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)),
                   __fmul_rn(w4, v4));
}

// --------------------------------------------------------------------------------------------
// NCHW forward.  grid = (num_rois, channel_slabs), block = 256.
// dynamic smem: sample table, tab_cap entries x (4 int offsets + 4 float weights).
// --------------------------------------------------------------------------------------------
constexpr int kRaThreads = 256;

__global__ void __launch_bounds__(kRaThreads)
roi_align_fwd_nchw_kernel(const float* __restrict__ input, const float* __restrict__ rois,
                          float* __restrict__ output, int C, int H, int W, int PH, int PW, float scale,
                          int sampling_ratio, int slab, int tab_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * slab;
  const int cn = min(slab, C - c0);
  const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, PH, PW, sampling_ratio);
  const int PP = PH * PW;
  const int S2 = g.gh * g.gw;
  const int ns = S2 * PP;
  const float* __restrict__ src0 = input + ((size_t)g.b * C + c0) * H * W;
  float* __restrict__ dst0 = output + ((size_t)n * C + c0) * PP;
  const size_t plane = (size_t)H * W;

  if (ns <= tab_cap) {
    int* t_off = reinterpret_cast<int*>(smem_raw);             // [4][ns]
    float* t_w = reinterpret_cast<float*>(smem_raw) + 4 * ns;  // [4][ns]
    for (int i = threadIdx.x; i < ns; i += kRaThreads) {
      const int s = i / PP, bin = i - s * PP;
      const int ph = bin / PW, pw = bin - ph * PW;
      const int iy = s / g.gw, ix = s - iy * g.gw;
      const float y = sample_coord(g.sh, ph, g.bin_h, iy, g.gh);
      const float x = sample_coord(g.sw, pw, g.bin_w, ix, g.gw);
      const Sample sm = make_sample(H, W, y, x);
      t_off[0 * ns + i] = sm.yl * W + sm.xl;
      t_off[1 * ns + i] = sm.yl * W + sm.xh;
      t_off[2 * ns + i] = sm.yh * W + sm.xl;
      t_off[3 * ns + i] = sm.yh * W + sm.xh;
      t_w[0 * ns + i] = sm.w1; t_w[1 * ns + i] = sm.w2; t_w[2 * ns + i] = sm.w3; t_w[3 * ns + i] = sm.w4;
    }
    __syncthreads();
    const int total = cn * PP;
    for (int o = threadIdx.x; o < total; o += kRaThreads) {
      const int c = o / PP, bin = o - c * PP;
      const float* __restrict__ src = src0 + (size_t)c * plane;
      float acc = 0.f;
      if (S2 == 4) {
        // the 16 taps first (memory-level parallelism), then the reference's summation order
        float v[16], wt[16];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int i = s * PP + bin;
#pragma unroll
          for (int k = 0; k < 4; ++k) { v[4 * s + k] = __ldg(src + t_off[k * ns + i]); wt[4 * s + k] = t_w[k * ns + i]; }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
          acc = __fadd_rn(acc, tap_sum(wt[4 * s], v[4 * s], wt[4 * s + 1], v[4 * s + 1], wt[4 * s + 2], v[4 * s + 2], wt[4 * s + 3], v[4 * s + 3]));
        dst0[o] = __fdiv_rn(acc, g.count);
        continue;
      }
      for (int s = 0; s < S2; ++s) {
        const int i = s * PP + bin;
        const float v1 = __ldg(src + t_off[i]), v2 = __ldg(src + t_off[ns + i]);
        const float v3 = __ldg(src + t_off[2 * ns + i]), v4 = __ldg(src + t_off[3 * ns + i]);
        acc = __fadd_rn(acc, tap_sum(t_w[i], v1, t_w[ns + i], v2, t_w[2 * ns + i], v3, t_w[3 * ns + i], v4));
      }
      dst0[o] = __fdiv_rn(acc, g.count);
    }
  } else {
    // adaptive grids too large for the table: evaluate samples on the fly
    const int total = cn * PP;
    for (int o = threadIdx.x; o < total; o += kRaThreads) {
      const int c = o / PP, bin = o - c * PP;
      const int ph = bin / PW, pw = bin - ph * PW;
      const float* __restrict__ src = src0 + (size_t)c * plane;
      float acc = 0.f;
      for (int iy = 0; iy < g.gh; ++iy) {
        const float y = sample_coord(g.sh, ph, g.bin_h, iy, g.gh);
        for (int ix = 0; ix < g.gw; ++ix) {
          const float x = sample_coord(g.sw, pw, g.bin_w, ix, g.gw);
          const Sample sm = make_sample(H, W, y, x);
          if (!sm.valid) continue;
          acc = __fadd_rn(acc, tap_sum(sm.w1, __ldg(src + sm.yl * W + sm.xl), sm.w2, __ldg(src + sm.yl * W + sm.xh),
                                       sm.w3, __ldg(src + sm.yh * W + sm.xl), sm.w4, __ldg(src + sm.yh * W + sm.xh)));
        }
      }
      dst0[o] = __fdiv_rn(acc, g.count);
    }
  }
}

// --------------------------------------------------------------------------------------------
// NHWC forward.  grid = (num_rois, C/slab), block = 256 (8 warps); slab = 32*VEC channels.
// Warp w owns bins w, w+8, ...; lane l owns channels [c0 + l*VEC, +VEC).
// Result tile [slab][PP(+pad)] staged in shared memory, then written NCHW-ordered.
// --------------------------------------------------------------------------------------------
template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void vec_load(const float* p, float (&v)[VEC]) {
  using T = typename VecT<VEC>::type;
  const T t = __ldg(reinterpret_cast<const T*>(p));
  const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = f[i];
}

template <int VEC>
__global__ void __launch_bounds__(kRaThreads)
roi_align_fwd_nhwc_kernel(const float* __restrict__ input, const float* __restrict__ rois,
                          float* __restrict__ output, int C, int H, int W, int PH, int PW, float scale,
                          int sampling_ratio) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int SLAB = 32 * VEC;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * SLAB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, PH, PW, sampling_ratio);
  const int PP = PH * PW;
  const int PPS = PP | 1;  // odd row stride -> conflict-free transposed access
  float* tile = reinterpret_cast<float*>(smem_raw);  // [SLAB][PPS]
  const int cl = c0 + lane * VEC;
  const bool c_ok = cl < C;  // C % VEC == 0 is guaranteed by the launcher
  const float* __restrict__ src = input + (size_t)g.b * H * W * C + (c_ok ? cl : 0);

  // One-axis halves of the samples, once per CTA (the validity test, clamping and bilinear weights of bilinear_interpolate
  // are separable in y and x): 2 PH + 2 PW evaluations instead of 4 per bin in every lane of every warp.  Combining them
  // reproduces make_sample's arithmetic exactly (same products, same rounding).
  __shared__ AxisSample s_y[2 * kRaMaxP], s_x[2 * kRaMaxP];
  const bool fast = g.gh == 2 && g.gw == 2 && PH <= kRaMaxP && PW <= kRaMaxP;
  if (fast) {
    const int t = threadIdx.x;
    if (t < 2 * PH) s_y[t] = make_axis_sample(H, sample_coord(g.sh, t >> 1, g.bin_h, t & 1, 2));
    else if (t < 2 * PH + 2 * PW) s_x[t - 2 * PH] = make_axis_sample(W, sample_coord(g.sw, (t - 2 * PH) >> 1, g.bin_w, (t - 2 * PH) & 1, 2));
    __syncthreads();
  }
  for (int bin = warp; bin < PP; bin += kRaThreads / 32) {
    const int ph = bin / PW, pw = bin - ph * PW;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    if (fast) {
      // sampling_ratio 2 (every reference config): request the 16 taps of the bin first, then reduce them in the
      // reference's order (sample (0,0), (0,1), (1,0), (1,1); w1 v1 + w2 v2 + w3 v3 + w4 v4 within a sample) -- the same
      // arithmetic, bit for bit, with 16 requests in flight per warp instead of 4.
      Sample sm[4];
      float v[16][VEC];
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        sm[s2] = combine_axis_samples(s_y[ph * 2 + (s2 >> 1)], s_x[pw * 2 + (s2 & 1)]);
        const bool ok = sm[s2].valid && c_ok;
        const size_t o1 = ok ? ((size_t)sm[s2].yl * W + sm[s2].xl) * C : 0, o2 = ok ? ((size_t)sm[s2].yl * W + sm[s2].xh) * C : 0;
        const size_t o3 = ok ? ((size_t)sm[s2].yh * W + sm[s2].xl) * C : 0, o4 = ok ? ((size_t)sm[s2].yh * W + sm[s2].xh) * C : 0;
        vec_load<VEC>(src + o1, v[4 * s2]);
        vec_load<VEC>(src + o2, v[4 * s2 + 1]);
        vec_load<VEC>(src + o3, v[4 * s2 + 2]);
        vec_load<VEC>(src + o4, v[4 * s2 + 3]);
      }
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        if (!sm[s2].valid || !c_ok) continue;
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          acc[k] = __fadd_rn(acc[k], tap_sum(sm[s2].w1, v[4 * s2][k], sm[s2].w2, v[4 * s2 + 1][k], sm[s2].w3, v[4 * s2 + 2][k],
                                             sm[s2].w4, v[4 * s2 + 3][k]));
      }
    } else
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = sample_coord(g.sh, ph, g.bin_h, iy, g.gh);
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = sample_coord(g.sw, pw, g.bin_w, ix, g.gw);
        const Sample sm = make_sample(H, W, y, x);
        if (!sm.valid || !c_ok) continue;
        float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
        vec_load<VEC>(src + ((size_t)sm.yl * W + sm.xl) * C, v1);
        vec_load<VEC>(src + ((size_t)sm.yl * W + sm.xh) * C, v2);
        vec_load<VEC>(src + ((size_t)sm.yh * W + sm.xl) * C, v3);
        vec_load<VEC>(src + ((size_t)sm.yh * W + sm.xh) * C, v4);
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          acc[k] = __fadd_rn(acc[k], tap_sum(sm.w1, v1[k], sm.w2, v2[k], sm.w3, v3[k], sm.w4, v4[k]));
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) tile[(lane * VEC + k) * PPS + bin] = __fdiv_rn(acc[k], g.count);
  }
  __syncthreads();
  const int cn = min(SLAB, C - c0);
  float* __restrict__ dst = output + ((size_t)n * C + c0) * PP;
  for (int o = threadIdx.x; o < cn * PP; o += kRaThreads) {
    const int c = o / PP, bin = o - c * PP;
    dst[o] = tile[c * PPS + bin];
  }
}

// --------------------------------------------------------------------------------------------
// Backward.  g_k = top * w_k / count (ROIAlign_cuda.cu:236-239), accumulated with red.add.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRaThreads)
roi_align_bwd_nchw_kernel(const float* __restrict__ grad, const float* __restrict__ rois,
                          float* __restrict__ gin, int C, int H, int W, int PH, int PW, float scale,
                          int sampling_ratio, int slab) {
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * slab;
  const int cn = min(slab, C - c0);
  const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, PH, PW, sampling_ratio);
  const int PP = PH * PW;
  const size_t plane = (size_t)H * W;
  float* __restrict__ dst0 = gin + ((size_t)g.b * C + c0) * plane;
  const float* __restrict__ top0 = grad + ((size_t)n * C + c0) * PP;
  const int total = cn * PP;
  for (int o = threadIdx.x; o < total; o += kRaThreads) {
    const int c = o / PP, bin = o - c * PP;
    const int ph = bin / PW, pw = bin - ph * PW;
    float* __restrict__ dst = dst0 + (size_t)c * plane;
    const float top = top0[o];
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = sample_coord(g.sh, ph, g.bin_h, iy, g.gh);
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = sample_coord(g.sw, pw, g.bin_w, ix, g.gw);
        const Sample sm = make_sample(H, W, y, x);
        if (!sm.valid) continue;
        atomicAdd(dst + sm.yl * W + sm.xl, __fdiv_rn(__fmul_rn(top, sm.w1), g.count));
        atomicAdd(dst + sm.yl * W + sm.xh, __fdiv_rn(__fmul_rn(top, sm.w2), g.count));
        atomicAdd(dst + sm.yh * W + sm.xl, __fdiv_rn(__fmul_rn(top, sm.w3), g.count));
        atomicAdd(dst + sm.yh * W + sm.xh, __fdiv_rn(__fmul_rn(top, sm.w4), g.count));
      }
    }
  }
}

template <int VEC>
__device__ __forceinline__ void vec_red_add(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  } else if constexpr (VEC == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else {
    atomicAdd(p, v[0]);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kRaThreads)
roi_align_bwd_nhwc_kernel(const float* __restrict__ grad, const float* __restrict__ rois,
                          float* __restrict__ gin, int C, int H, int W, int PH, int PW, float scale,
                          int sampling_ratio) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int SLAB = 32 * VEC;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * SLAB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, PH, PW, sampling_ratio);
  const int PP = PH * PW;
  const int PPS = PP | 1;
  float* tile = reinterpret_cast<float*>(smem_raw);  // [SLAB][PPS] grad_output slab
  const int cn = min(SLAB, C - c0);
  const float* __restrict__ top = grad + ((size_t)n * C + c0) * PP;
  for (int o = threadIdx.x; o < cn * PP; o += kRaThreads) {
    const int c = o / PP, bin = o - c * PP;
    tile[c * PPS + bin] = top[o];
  }
  __syncthreads();
  const int cl = c0 + lane * VEC;
  if (cl >= C) return;
  float* __restrict__ dst = gin + (size_t)g.b * H * W * C + cl;
  for (int bin = warp; bin < PP; bin += kRaThreads / 32) {
    const int ph = bin / PW, pw = bin - ph * PW;
    float t[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) t[k] = tile[(lane * VEC + k) * PPS + bin];
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = sample_coord(g.sh, ph, g.bin_h, iy, g.gh);
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = sample_coord(g.sw, pw, g.bin_w, ix, g.gw);
        const Sample sm = make_sample(H, W, y, x);
        if (!sm.valid) continue;
        float g1[VEC], g2[VEC], g3[VEC], g4[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          g1[k] = __fdiv_rn(__fmul_rn(t[k], sm.w1), g.count);
          g2[k] = __fdiv_rn(__fmul_rn(t[k], sm.w2), g.count);
          g3[k] = __fdiv_rn(__fmul_rn(t[k], sm.w3), g.count);
          g4[k] = __fdiv_rn(__fmul_rn(t[k], sm.w4), g.count);
        }
        vec_red_add<VEC>(dst + ((size_t)sm.yl * W + sm.xl) * C, g1);
        vec_red_add<VEC>(dst + ((size_t)sm.yl * W + sm.xh) * C, g2);
        vec_red_add<VEC>(dst + ((size_t)sm.yh * W + sm.xl) * C, g3);
        vec_red_add<VEC>(dst + ((size_t)sm.yh * W + sm.xh) * C, g4);
      }
    }
  }
}

static int pick_slab(int num_rois, int C) {
  // enough CTAs for >= 4 waves at 8 CTAs/SM, but keep at least 16 channels per CTA so the
  // sample table is amortised
  int slab = C;
  while (slab > 16 && (int64_t)num_rois * ceil_div(C, slab) < (int64_t)kNumSMs * 8) slab = (slab + 1) / 2;
  return slab;
}

}  // namespace mrb

using namespace mrb;

static int check_ra_args(const void* a, const void* b, const void* c, int num_rois, int batch, int C, int H,
                         int W, int PH, int PW, int layout) {
  if (num_rois < 0 || batch < 0 || C < 0 || H < 0 || W < 0 || PH <= 0 || PW <= 0) return MRB_ERR_BAD_ARG;
  if (layout != MRB_LAYOUT_NCHW && layout != MRB_LAYOUT_NHWC) return MRB_ERR_BAD_ARG;
  if ((int64_t)H * W >= (1ll << 31)) return MRB_ERR_UNSUPPORTED;
  if (num_rois > 0 && C > 0 && (!a || !b || !c)) return MRB_ERR_BAD_ARG;
  return MRB_OK;
}

MRB_API int mrb_roi_align_fwd(const float* input, const float* rois, float* output, int num_rois, int batch,
                              int channels, int height, int width, int pooled_h, int pooled_w,
                              float spatial_scale, int sampling_ratio, int layout, mrb_stream_t stream_) {
  int rc = check_ra_args(input, rois, output, num_rois, batch, channels, height, width, pooled_h, pooled_w, layout);
  if (rc) return rc;
  if (num_rois == 0 || channels == 0) return MRB_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int PP = pooled_h * pooled_w;
  if (layout == MRB_LAYOUT_NHWC && channels % 2 == 0) {
    const bool v4 = (channels % 4 == 0) && ((size_t)128 * (PP | 1) * 4 <= 100 * 1024);
    const int slab = v4 ? 128 : 64;
    const size_t smem = (size_t)slab * (PP | 1) * sizeof(float);
    if (smem > 200 * 1024) return MRB_ERR_UNSUPPORTED;
    dim3 grid(num_rois, ceil_div(channels, slab));
    if (v4) {
      MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fwd_nhwc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      roi_align_fwd_nhwc_kernel<4><<<grid, kRaThreads, smem, stream>>>(input, rois, output, channels, height, width,
                                                                      pooled_h, pooled_w, spatial_scale, sampling_ratio);
    } else {
      MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fwd_nhwc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      roi_align_fwd_nhwc_kernel<2><<<grid, kRaThreads, smem, stream>>>(input, rois, output, channels, height, width,
                                                                      pooled_h, pooled_w, spatial_scale, sampling_ratio);
    }
    MRB_LAUNCH_CHECK();
    return MRB_OK;
  }
  if (layout == MRB_LAYOUT_NHWC) {
    const size_t smem = (size_t)32 * (PP | 1) * sizeof(float);
    if (smem > 200 * 1024) return MRB_ERR_UNSUPPORTED;
    MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fwd_nhwc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(num_rois, ceil_div(channels, 32));
    roi_align_fwd_nhwc_kernel<1><<<grid, kRaThreads, smem, stream>>>(input, rois, output, channels, height, width,
                                                                    pooled_h, pooled_w, spatial_scale, sampling_ratio);
    MRB_LAUNCH_CHECK();
    return MRB_OK;
  }
  // NCHW
  const int slab = pick_slab(num_rois, channels);
  int tab_cap = (sampling_ratio > 0) ? sampling_ratio * sampling_ratio * PP : 4 * PP;
  if (tab_cap > 3072) tab_cap = 3072;  // 96 KB
  const size_t smem = (size_t)tab_cap * 32;
  MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_fwd_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(num_rois, ceil_div(channels, slab));
  roi_align_fwd_nchw_kernel<<<grid, kRaThreads, smem, stream>>>(input, rois, output, channels, height, width, pooled_h,
                                                               pooled_w, spatial_scale, sampling_ratio, slab, tab_cap);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_roi_align_bwd(const float* grad_output, const float* rois, float* grad_input, int num_rois,
                              int batch, int channels, int height, int width, int pooled_h, int pooled_w,
                              float spatial_scale, int sampling_ratio, int layout, mrb_stream_t stream_) {
  int rc = check_ra_args(grad_output, rois, grad_input, num_rois, batch, channels, height, width, pooled_h, pooled_w, layout);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t total = (size_t)batch * channels * height * width;
  if (total == 0) return MRB_OK;
  if (!grad_input) return MRB_ERR_BAD_ARG;
  MRB_CUDA_TRY(cudaMemsetAsync(grad_input, 0, total * sizeof(float), stream));
  if (num_rois == 0) return MRB_OK;
  const int PP = pooled_h * pooled_w;
  if (layout == MRB_LAYOUT_NHWC) {
    const int vec = (channels % 4 == 0 && (size_t)128 * (PP | 1) * 4 <= 100 * 1024) ? 4 : (channels % 2 == 0 ? 2 : 1);
    const int slab = 32 * vec;
    const size_t smem = (size_t)slab * (PP | 1) * sizeof(float);
    if (smem > 200 * 1024) return MRB_ERR_UNSUPPORTED;
    dim3 grid(num_rois, ceil_div(channels, slab));
#define MRB_LAUNCH_BWD(V)                                                                                             \
  MRB_CUDA_TRY(cudaFuncSetAttribute(roi_align_bwd_nhwc_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  roi_align_bwd_nhwc_kernel<V><<<grid, kRaThreads, smem, stream>>>(grad_output, rois, grad_input, channels, height, width, \
                                                                  pooled_h, pooled_w, spatial_scale, sampling_ratio)
    if (vec == 4) { MRB_LAUNCH_BWD(4); } else if (vec == 2) { MRB_LAUNCH_BWD(2); } else { MRB_LAUNCH_BWD(1); }
#undef MRB_LAUNCH_BWD
    MRB_LAUNCH_CHECK();
    return MRB_OK;
  }
  const int slab = pick_slab(num_rois, channels);
  dim3 grid(num_rois, ceil_div(channels, slab));
  roi_align_bwd_nchw_kernel<<<grid, kRaThreads, 0, stream>>>(grad_output, rois, grad_input, channels, height, width,
                                                            pooled_h, pooled_w, spatial_scale, sampling_ratio, slab);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
