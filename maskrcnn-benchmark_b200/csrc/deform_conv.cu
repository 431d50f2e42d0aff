// deform_conv.cu -- deformable convolution v1 / v2 (modulated), forward and backward, sm_100a.
//
// Replaces deformable_im2col / col2im / col2im_coord and their modulated twins (reference
// csrc/cuda/deform_conv_kernel_cuda.cu:197-874) and the host orchestration in
// csrc/cuda/deform_conv_cuda.cu:158-691 (which calls cuBLAS through at::addmm_).
//
// Round-1 structure (correctness first, one code path for v1 and v2 -- v1 is v2 with mask == 1):
//   forward : per image   sample -> columns[Cin*kh*kw, Ho*Wo] (workspace) ; out = W . columns (+bias)
//   backward: per image   colgrad = W^T . gout ; coord kernel -> d_offset, d_mask ;
//                         col2im (red.add) -> d_input ; sample -> columns ; dW += gout . columns^T
// The GEMMs are an in-house fp32 SIMT kernel with arbitrary strides (so no transposed copies such as
// deform_conv_cuda.cu:451-459 are needed) and split-K for the weight gradient; fp32 keeps the op
// within 1e-4 of the reference.  Everything runs on the caller's stream, one image at a time, with
// one columns buffer in the caller's workspace (the reference allocates columns + output_buffer +
// gradOutputBuffer per call).
#include "common.cuh"

namespace mrb {

struct DcnGeom {
  int cin, H, W, cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, Ho, Wo;
};

// deformable_im2col_bilinear / dmcn_im2col_bilinear (deform_conv_kernel_cuda.cu:91-121,473-503)
__device__ __forceinline__ float dcn_bilinear(const float* __restrict__ im, int H, int W, float h, float w) {
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (h_low >= 0 && w_low >= 0) v1 = __ldg(im + h_low * W + w_low);
  if (h_low >= 0 && w_high <= W - 1) v2 = __ldg(im + h_low * W + w_high);
  if (h_high <= H - 1 && w_low >= 0) v3 = __ldg(im + h_high * W + w_low);
  if (h_high <= H - 1 && w_high <= W - 1) v4 = __ldg(im + h_high * W + w_high);
  return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

// --- sample: columns[(c*kh*kw + tap), pix] = bilinear(input[c], p0 + tap + offset) * mask ----------
// deformable_im2col_gpu_kernel / modulated_ (deform_conv_kernel_cuda.cu:197-250,577-640), batch 1.
__global__ void __launch_bounds__(256)
dcn_sample_kernel(int n, DcnGeom g, const float* __restrict__ im, const float* __restrict__ offset,
                  const float* __restrict__ mask, float* __restrict__ col) {
  const int P = g.Ho * g.Wo, taps = g.kh * g.kw, cpg = g.cin / g.dg;
  for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < n; index += gridDim.x * blockDim.x) {
    const int pix = index % P, c = index / P;
    const int w_col = pix % g.Wo, h_col = pix / g.Wo;
    const int d = c / cpg;
    const int h_in = h_col * g.sh - g.ph, w_in = w_col * g.sw - g.pw;
    const float* __restrict__ imc = im + (size_t)c * g.H * g.W;
    const float* __restrict__ off = offset + (size_t)d * 2 * taps * P + pix;
    const float* __restrict__ msk = mask ? mask + (size_t)d * taps * P + pix : nullptr;
    float* __restrict__ out = col + (size_t)c * taps * P + pix;
    for (int t = 0; t < taps; ++t) {
      const int i = t / g.kw, j = t - i * g.kw;
      const float h_im = h_in + i * g.dh + off[(size_t)(2 * t) * P];
      const float w_im = w_in + j * g.dw + off[(size_t)(2 * t + 1) * P];
      float val = 0.f;
      if (h_im > -1 && w_im > -1 && h_im < g.H && w_im < g.W) val = dcn_bilinear(imc, g.H, g.W, h_im, w_im);
      if (msk) val *= msk[(size_t)t * P];
      out[(size_t)t * P] = val;
    }
  }
}

// --- col2im: d_input += scatter(colgrad * mask) ---------------------------------------------------
// deformable_col2im_gpu_kernel / modulated_ (deform_conv_kernel_cuda.cu:286-342,642-700).  The
// reference scans a 5x5 window for the <=4 pixels with |delta| < 1; the same four taps are addressed
// directly here (floor/floor+1), with the reference's get_gradient_weight formula.
__global__ void __launch_bounds__(256)
dcn_col2im_kernel(int n, DcnGeom g, const float* __restrict__ colgrad, const float* __restrict__ offset,
                  const float* __restrict__ mask, float* __restrict__ grad_im) {
  const int P = g.Ho * g.Wo, taps = g.kh * g.kw, cpg = g.cin / g.dg;
  for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < n; index += gridDim.x * blockDim.x) {
    const int pix = index % P;
    const int t = (index / P) % taps;
    const int c = index / P / taps;
    const int i = t / g.kw, j = t - i * g.kw;
    const int d = c / cpg;
    const int w_out = pix % g.Wo, h_out = pix / g.Wo;
    const float* __restrict__ off = offset + (size_t)d * 2 * taps * P + pix;
    const float h = h_out * g.sh - g.ph + i * g.dh + off[(size_t)(2 * t) * P];
    const float w = w_out * g.sw - g.pw + j * g.dw + off[(size_t)(2 * t + 1) * P];
    float top = colgrad[index];
    if (mask) top *= mask[((size_t)d * taps + t) * P + pix];
    // get_gradient_weight returns 0 outside (-1, H) x (-1, W)
    if (h <= -1 || h >= g.H || w <= -1 || w >= g.W) continue;
    const int h_low = (int)floorf(h), w_low = (int)floorf(w);
    float* __restrict__ dst = grad_im + (size_t)c * g.H * g.W;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = h_low + dy, x = w_low + dx;
        if (y < 0 || y >= g.H || x < 0 || x >= g.W) continue;
        // the reference additionally requires |h - y| < 1 and |w - x| < 1 (true for these taps unless
        // h (w) is an exact integer and dy (dx) == 1, where the weight below is 0 anyway)
        const float wy = dy ? (h + 1 - y) : (y + 1 - h);
        const float wx = dx ? (w + 1 - x) : (x + 1 - w);
        const float wgt = wy * wx;
        if (wgt != 0.f) atomicAdd(dst + y * g.W + x, wgt * top);
      }
  }
}

// --- coord: d_offset (and d_mask) --------------------------------------------------------------
// deformable_col2im_coord_gpu_kernel / modulated_ (deform_conv_kernel_cuda.cu:380-443,702-774).
// One thread per (dg, tap, pixel) produces BOTH offset components and the mask gradient, looping over
// the channels of the deformable group (the reference runs one thread per offset channel).
__global__ void __launch_bounds__(256)
dcn_coord_kernel(int n, DcnGeom g, const float* __restrict__ colgrad, const float* __restrict__ im,
                 const float* __restrict__ offset, const float* __restrict__ mask, float* __restrict__ grad_offset,
                 float* __restrict__ grad_mask) {
  const int P = g.Ho * g.Wo, taps = g.kh * g.kw, cpg = g.cin / g.dg;
  for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < n; index += gridDim.x * blockDim.x) {
    const int pix = index % P;
    const int t = (index / P) % taps;
    const int d = index / P / taps;
    const int i = t / g.kw, j = t - i * g.kw;
    const int w_out = pix % g.Wo, h_out = pix / g.Wo;
    const size_t obase = (size_t)d * 2 * taps * P + pix;
    float h = h_out * g.sh - g.ph + i * g.dh + offset[obase + (size_t)(2 * t) * P];
    float w = w_out * g.sw - g.pw + j * g.dw + offset[obase + (size_t)(2 * t + 1) * P];
    const float m = mask ? mask[((size_t)d * taps + t) * P + pix] : 1.f;
    float gh = 0.f, gw = 0.f, gm = 0.f;
    const bool inside = !(h <= -1 || w <= -1 || h >= g.H || w >= g.W);
    if (inside) {
      const int h_low = (int)floorf(h), w_low = (int)floorf(w);
      const int h_high = h_low + 1, w_high = w_low + 1;
      const bool ok1 = h_low >= 0 && w_low >= 0, ok2 = h_low >= 0 && w_high <= g.W - 1;
      const bool ok3 = h_high <= g.H - 1 && w_low >= 0, ok4 = h_high <= g.H - 1 && w_high <= g.W - 1;
      const float lh = h - h_low, lw = w - w_low;
      for (int cc = 0; cc < cpg; ++cc) {
        const int c = d * cpg + cc;
        const float* __restrict__ imc = im + (size_t)c * g.H * g.W;
        const float cg = colgrad[((size_t)c * taps + t) * P + pix];
        const float v1 = ok1 ? __ldg(imc + h_low * g.W + w_low) : 0.f;
        const float v2 = ok2 ? __ldg(imc + h_low * g.W + w_high) : 0.f;
        const float v3 = ok3 ? __ldg(imc + h_high * g.W + w_low) : 0.f;
        const float v4 = ok4 ? __ldg(imc + h_high * g.W + w_high) : 0.f;
        // get_coordinate_weight, bp_dir 0 (d/dh) and 1 (d/dw)
        const float wh = -(1 - lw) * v1 - lw * v2 + (1 - lw) * v3 + lw * v4;
        const float ww = -(1 - lh) * v1 + (1 - lh) * v2 - lh * v3 + lh * v4;
        gh += wh * cg * m;
        gw += ww * cg * m;
        gm += cg * ((1 - lh) * (1 - lw) * v1 + (1 - lh) * lw * v2 + lh * (1 - lw) * v3 + lh * lw * v4);
      }
    }
    grad_offset[obase + (size_t)(2 * t) * P] = gh;
    grad_offset[obase + (size_t)(2 * t + 1) * P] = gw;
    if (grad_mask) grad_mask[((size_t)d * taps + t) * P + pix] = gm;
  }
}

// --- generic strided fp32 GEMM: C(m,n) (=|+=) alpha * sum_k A(m,k) B(k,n) (+ bias[m]) -------------
constexpr int GM = 64, GN = 64, GK = 16, GT = 256;

struct GemmArgs {
  const float* A; long long a_rs, a_cs;
  const float* B; long long b_rs, b_cs;
  float* C; long long c_rs, c_cs;
  const float* bias;
  int M, N, K, k_per_split;
  float alpha;
  int accumulate;  // 1: atomicAdd into C, 0: overwrite
};

__global__ void __launch_bounds__(GT)
dcn_gemm_kernel(GemmArgs p) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int k_begin = blockIdx.z * p.k_per_split;
  const int k_end = min(p.K, k_begin + p.k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 micro-tile
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
    for (int e = threadIdx.x; e < GM * GK; e += GT) {
      // pick the fast-varying index along the unit-stride axis of A
      int mm, kk;
      if (p.a_cs == 1) { kk = e % GK; mm = e / GK; } else { mm = e % GM; kk = e / GM; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < k_end) ? __ldg(p.A + m * p.a_rs + k * p.a_cs) : 0.f;
    }
    for (int e = threadIdx.x; e < GN * GK; e += GT) {
      int nn, kk;
      if (p.b_cs == 1) { nn = e % GN; kk = e / GN; } else { kk = e % GK; nn = e / GK; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < p.N && k < k_end) ? __ldg(p.B + k * p.b_rs + n * p.b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const float bv = (p.bias && blockIdx.z == 0) ? p.bias[m] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float* c = p.C + m * p.c_rs + n * p.c_cs;
      const float v = p.alpha * acc[i][j] + bv;
      if (p.accumulate) atomicAdd(c, v); else *c = v;
    }
  }
}

static int gemm(cudaStream_t s, const float* A, long long a_rs, long long a_cs, const float* B, long long b_rs,
                long long b_cs, float* C, long long c_rs, long long c_cs, const float* bias, int M, int N, int K,
                float alpha, bool accumulate, int splits) {
  GemmArgs p{A, a_rs, a_cs, B, b_rs, b_cs, C, c_rs, c_cs, bias, M, N, K, 0, alpha, accumulate ? 1 : 0};
  if (!accumulate) splits = 1;
  splits = max(1, min(splits, ceil_div(K, GK)));
  p.k_per_split = ceil_div(ceil_div(K, splits), GK) * GK;
  splits = ceil_div(K, p.k_per_split);
  dim3 grid(ceil_div(N, GN), ceil_div(M, GM), splits);
  dcn_gemm_kernel<<<grid, GT, 0, s>>>(p);
  return (int)cudaGetLastError();
}

__global__ void __launch_bounds__(256)
dcn_bias_grad_kernel(const float* __restrict__ gout, float* __restrict__ gbias, int P) {
  // one CTA per output channel of one image: gbias[m] += sum_pix gout[m, pix]
  const int m = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < P; i += 256) s += gout[(size_t)m * P + i];
  __shared__ float red[8];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(gbias + m, t);
  }
}

static int dcn_geom(const mrb_dcn_params* p, DcnGeom& g) {
  if (!p) return MRB_ERR_BAD_ARG;
  if (p->batch < 0 || p->cin <= 0 || p->cout <= 0 || p->kh <= 0 || p->kw <= 0 || p->stride_h <= 0 || p->stride_w <= 0 ||
      p->dil_h <= 0 || p->dil_w <= 0 || p->groups <= 0 || p->deformable_groups <= 0 || p->pad_h < 0 || p->pad_w < 0)
    return MRB_ERR_BAD_ARG;
  if (p->cin % p->groups || p->cout % p->groups || p->cin % p->deformable_groups) return MRB_ERR_BAD_ARG;
  g.cin = p->cin; g.H = p->height; g.W = p->width; g.cout = p->cout; g.kh = p->kh; g.kw = p->kw;
  g.sh = p->stride_h; g.sw = p->stride_w; g.ph = p->pad_h; g.pw = p->pad_w; g.dh = p->dil_h; g.dw = p->dil_w;
  g.groups = p->groups; g.dg = p->deformable_groups;
  g.Ho = (g.H + 2 * g.ph - (g.dh * (g.kh - 1) + 1)) / g.sh + 1;
  g.Wo = (g.W + 2 * g.pw - (g.dw * (g.kw - 1) + 1)) / g.sw + 1;
  if (g.Ho <= 0 || g.Wo <= 0) return MRB_ERR_BAD_ARG;
  if ((int64_t)g.cin * g.kh * g.kw * g.Ho * g.Wo >= (1ll << 31)) return MRB_ERR_UNSUPPORTED;
  return MRB_OK;
}

}  // namespace mrb
using namespace mrb;

MRB_API size_t mrb_deform_conv_workspace_bytes(const mrb_dcn_params* p) {
  DcnGeom g;
  if (dcn_geom(p, g)) return 0;
  return (size_t)g.cin * g.kh * g.kw * g.Ho * g.Wo * sizeof(float);  // one image's columns
}

MRB_API int mrb_deform_conv_fwd(const mrb_dcn_params* p, const float* input, const float* offset, const float* mask,
                                const float* weight, const float* bias, float* output, void* workspace,
                                size_t workspace_bytes, mrb_stream_t stream_) {
  DcnGeom g;
  int rc = dcn_geom(p, g);
  if (rc) return rc;
  if (p->batch == 0) return MRB_OK;
  if (!input || !offset || !weight || !output || !workspace) return MRB_ERR_BAD_ARG;
  if (workspace_bytes < mrb_deform_conv_workspace_bytes(p)) return MRB_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream_;
  float* col = (float*)workspace;
  const int P = g.Ho * g.Wo, taps = g.kh * g.kw;
  const int Kg = g.cin / g.groups * taps, Mg = g.cout / g.groups;
  for (int b = 0; b < p->batch; ++b) {
    const float* im = input + (size_t)b * g.cin * g.H * g.W;
    const float* off = offset + (size_t)b * g.dg * 2 * taps * P;
    const float* msk = mask ? mask + (size_t)b * g.dg * taps * P : nullptr;
    const int n = g.cin * P;
    dcn_sample_kernel<<<grid_for(n, 256, 8, 8), 256, 0, s>>>(n, g, im, off, msk, col);
    MRB_LAUNCH_CHECK();
    for (int gi = 0; gi < g.groups; ++gi) {
      rc = gemm(s, weight + (size_t)gi * Mg * Kg, Kg, 1, col + (size_t)gi * Kg * P, P, 1,
                output + ((size_t)b * g.cout + (size_t)gi * Mg) * P, P, 1, bias ? bias + gi * Mg : nullptr, Mg, P, Kg,
                1.f, false, 1);
      if (rc) return rc;
    }
  }
  return MRB_OK;
}

MRB_API int mrb_deform_conv_bwd(const mrb_dcn_params* p, const float* input, const float* offset, const float* mask,
                                const float* weight, const float* grad_output, float* grad_input, float* grad_offset,
                                float* grad_mask, float* grad_weight, float* grad_bias, float scale, void* workspace,
                                size_t workspace_bytes, mrb_stream_t stream_) {
  DcnGeom g;
  int rc = dcn_geom(p, g);
  if (rc) return rc;
  if (p->batch == 0) return MRB_OK;
  if (!input || !offset || !grad_output || !workspace) return MRB_ERR_BAD_ARG;
  const bool need_in = grad_input || grad_offset || grad_mask;
  if (need_in && (!weight || !grad_input || !grad_offset)) return MRB_ERR_BAD_ARG;
  if (grad_mask && !mask) return MRB_ERR_BAD_ARG;
  if (workspace_bytes < mrb_deform_conv_workspace_bytes(p)) return MRB_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream_;
  float* col = (float*)workspace;
  const int P = g.Ho * g.Wo, taps = g.kh * g.kw;
  const int Kg = g.cin / g.groups * taps, Mg = g.cout / g.groups;
  for (int b = 0; b < p->batch; ++b) {
    const float* im = input + (size_t)b * g.cin * g.H * g.W;
    const float* off = offset + (size_t)b * g.dg * 2 * taps * P;
    const float* msk = mask ? mask + (size_t)b * g.dg * taps * P : nullptr;
    const float* gout = grad_output + (size_t)b * g.cout * P;
    if (need_in) {
      for (int gi = 0; gi < g.groups; ++gi) {
        // colgrad_g[Kg, P] = W_g^T[Kg, Mg] . gout_g[Mg, P]
        rc = gemm(s, weight + (size_t)gi * Mg * Kg, 1, Kg, gout + (size_t)gi * Mg * P, P, 1, col + (size_t)gi * Kg * P, P, 1,
                  nullptr, Kg, P, Mg, 1.f, false, 1);
        if (rc) return rc;
      }
      const int nc = g.dg * taps * P;
      dcn_coord_kernel<<<grid_for(nc, 256, 8, 8), 256, 0, s>>>(nc, g, col, im, off, msk,
                                                             grad_offset + (size_t)b * g.dg * 2 * taps * P,
                                                             grad_mask ? grad_mask + (size_t)b * g.dg * taps * P : nullptr);
      MRB_LAUNCH_CHECK();
      const int ni = g.cin * taps * P;
      dcn_col2im_kernel<<<grid_for(ni, 256, 8, 8), 256, 0, s>>>(ni, g, col, off, msk, grad_input + (size_t)b * g.cin * g.H * g.W);
      MRB_LAUNCH_CHECK();
    }
    if (grad_weight) {
      const int n = g.cin * P;
      dcn_sample_kernel<<<grid_for(n, 256, 8, 8), 256, 0, s>>>(n, g, im, off, msk, col);
      MRB_LAUNCH_CHECK();
      for (int gi = 0; gi < g.groups; ++gi) {
        // dW_g[Mg, Kg] += scale * gout_g[Mg, P] . col_g^T[P, Kg]   (split over P)
        const int tiles = ceil_div(Mg, GM) * ceil_div(Kg, GN);
        const int splits = max(1, (2 * kNumSMs) / max(tiles, 1));
        rc = gemm(s, gout + (size_t)gi * Mg * P, P, 1, col + (size_t)gi * Kg * P, 1, P, grad_weight + (size_t)gi * Mg * Kg, Kg, 1,
                  nullptr, Mg, Kg, P, scale, true, splits);
        if (rc) return rc;
      }
    }
    if (grad_bias) {
      dcn_bias_grad_kernel<<<g.cout, 256, 0, s>>>(gout, grad_bias, P);
      MRB_LAUNCH_CHECK();
    }
  }
  return MRB_OK;
}
