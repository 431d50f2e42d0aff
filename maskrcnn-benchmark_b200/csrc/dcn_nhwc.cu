// dcn_nhwc.cu -- deformable convolution v1 / v2 for the tensor-core path: bilinear sampler and its backward on
// NHWC bf16 activations, around the tcgen05 GEMMs of conv_tc.cu.
//
// Replaces, for the model's own DCN layers (reference modeling/backbone/resnet.py:286-300 -> layers/misc.py:114-203
// DFConv2d -> layers/dcn/deform_conv_func.py), the fp32 NCHW kernels deformable_im2col / col2im / col2im_coord and
// their modulated twins (csrc/cuda/deform_conv_kernel_cuda.cu:197-874) together with the cuBLAS GEMMs of
// csrc/cuda/deform_conv_cuda.cu:158-691:
//
//   forward : cols[pix][tap*C + c] = bilinear(x[n, :, :, c], p(pix, tap) + offset) * mask      (this file, bf16)
//             y = cols . W^T  (+ frozen-BN scale/shift, ReLU)          -> mrb_conv2d_fwd, 1x1 over K = 9C
//   backward: gcols = g . W                                            -> mrb_conv2d_dgrad_prepared
//             dW    = g^T . cols                                       -> mrb_conv2d_wgrad
//             dx (red.add, fp32 NHWC), d_offset, d_mask(logit)  <- gcols, x, offsets             (this file)
//
// `cols` is [N*Ho*Wo][taps*C] bf16 -- the K axis ordered (tap, channel) is exactly the memory order of a KRSC
// (torch.channels_last) 3x3 filter, so the model's weight tensor is the GEMM's B operand without a copy.
// The offset / mask tensor `om` is the fp32 NHWC output of the offset conv ([N,Ho,Wo,oc_pitch]; channels
// 2t, 2t+1 = (dh, dw) of tap t, and for v2 channels 2*taps + t = mask LOGIT of tap t: DFConv2d applies the
// sigmoid, layers/misc.py:186-187, fused here in both directions).  deformable_groups == 1, groups == 1.
// One warp handles one (pixel, tap): the 4 corner rows are contiguous C-vectors in NHWC, so every global access is a
// full-width coalesced 16-byte-per-lane access; offsets and bilinear weights are warp-uniform scalars.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mrb {

struct DcnNhwc {
  int N, H, W, C, Ho, Wo, kh, kw, stride, pad, dil, oc_pitch, modulated;
};

struct Corner {
  bool inside;        // sample position within (-1, H) x (-1, W)
  bool ok[4];         // corner inside the image
  long long off[4];   // element offset of the corner's channel vector (valid if ok)
  float wgt[4];       // bilinear weights hh*hw, hh*lw, lh*hw, lh*lw
  float lh, lw;
};

__device__ __forceinline__ void dcn_corner(const DcnNhwc& g, int n, int ho, int wo, int tap, const float* __restrict__ om_pix,
                                           Corner& c, float& mask, float& mask_logit_sig) {
  const int r = tap / g.kw, q = tap - r * g.kw;
  const float h = (float)(ho * g.stride - g.pad + r * g.dil) + __ldg(om_pix + 2 * tap);
  const float w = (float)(wo * g.stride - g.pad + q * g.dil) + __ldg(om_pix + 2 * tap + 1);
  mask = 1.f; mask_logit_sig = 1.f;
  if (g.modulated) {
    const float l = __ldg(om_pix + 2 * g.kh * g.kw + tap);
    mask = 1.f / (1.f + __expf(-l));
    mask_logit_sig = mask;
  }
  c.inside = (h > -1.f) && (w > -1.f) && (h < (float)g.H) && (w < (float)g.W);
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - (float)h_low, lw = w - (float)w_low, hh = 1.f - lh, hw = 1.f - lw;
  c.lh = lh; c.lw = lw;
  c.ok[0] = c.inside && h_low >= 0 && w_low >= 0;
  c.ok[1] = c.inside && h_low >= 0 && w_high <= g.W - 1;
  c.ok[2] = c.inside && h_high <= g.H - 1 && w_low >= 0;
  c.ok[3] = c.inside && h_high <= g.H - 1 && w_high <= g.W - 1;
  const long long base = (long long)n * g.H * g.W;
  c.off[0] = (base + (long long)h_low * g.W + w_low) * g.C;
  c.off[1] = (base + (long long)h_low * g.W + w_high) * g.C;
  c.off[2] = (base + (long long)h_high * g.W + w_low) * g.C;
  c.off[3] = (base + (long long)h_high * g.W + w_high) * g.C;
  c.wgt[0] = hh * hw; c.wgt[1] = hh * lw; c.wgt[2] = lh * hw; c.wgt[3] = lh * lw;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(p[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

// ---------------------------------------------------------------------------------------- sampler
__global__ void __launch_bounds__(256)
dcn_sample_nhwc_kernel(const DcnNhwc g, const __nv_bfloat16* __restrict__ x, const float* __restrict__ om,
                       __nv_bfloat16* __restrict__ cols) {
  const int taps = g.kh * g.kw, lane = threadIdx.x & 31;
  const long long items = (long long)g.N * g.Ho * g.Wo * taps;
  const int c8n = g.C >> 3;
  for (long long it = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); it < items; it += (long long)gridDim.x * 8) {
    const int tap = (int)(it % taps);
    const long long pix = it / taps;
    const int wo = (int)(pix % g.Wo), ho = (int)((pix / g.Wo) % g.Ho), n = (int)(pix / ((long long)g.Wo * g.Ho));
    Corner c; float m, ms;
    dcn_corner(g, n, ho, wo, tap, om + pix * g.oc_pitch, c, m, ms);
    __nv_bfloat16* __restrict__ dst = cols + (pix * taps + tap) * g.C;
    for (int c8 = lane; c8 < c8n; c8 += 32) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c.ok[k]) {
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(x + c.off[k]) + c8), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(c.wgt[k], f[j], acc[j]);
        }
      }
      uint4 pk;
      __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) p2[j] = __floats2bfloat162_rn(acc[2 * j] * m, acc[2 * j + 1] * m);
      reinterpret_cast<uint4*>(dst)[c8] = pk;
    }
  }
}

// ---------------------------------------------------------------------------------------- backward
// Per (pixel, tap): d_offset (2 values), d_mask-logit (1 value, v2), and the scatter of gcols * mask * bilinear weight
// into grad_x (fp32 NHWC, red.add.v4.f32).  Formulas: deform_conv_kernel_cuda.cu get_gradient_weight (col2im),
// get_coordinate_weight (coord), dmcn_* twins; d(val)/dh = -hw*v1 - lw*v2 + hw*v3 + lw*v4, d/dw likewise.
__global__ void __launch_bounds__(256)
dcn_backward_nhwc_kernel(const DcnNhwc g, const __nv_bfloat16* __restrict__ x, const float* __restrict__ om,
                         const __nv_bfloat16* __restrict__ gcols, float* __restrict__ grad_x, float* __restrict__ grad_om) {
  const int taps = g.kh * g.kw, lane = threadIdx.x & 31;
  const long long items = (long long)g.N * g.Ho * g.Wo * taps;
  const int c8n = g.C >> 3;
  for (long long it = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); it < items; it += (long long)gridDim.x * 8) {
    const int tap = (int)(it % taps);
    const long long pix = it / taps;
    const int wo = (int)(pix % g.Wo), ho = (int)((pix / g.Wo) % g.Ho), n = (int)(pix / ((long long)g.Wo * g.Ho));
    Corner c; float m, ms;
    dcn_corner(g, n, ho, wo, tap, om + pix * g.oc_pitch, c, m, ms);
    float gh = 0.f, gw = 0.f, gm = 0.f;
    if (c.inside) {
      const __nv_bfloat16* __restrict__ src = gcols + (pix * taps + tap) * g.C;
      const float hh = 1.f - c.lh, hw = 1.f - c.lw;
      for (int c8 = lane; c8 < c8n; c8 += 32) {
        float gc[8], v[4][8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(src) + c8), gc);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c.ok[k]) unpack8(__ldg(reinterpret_cast<const uint4*>(x + c.off[k]) + c8), v[k]);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float wh = -hw * v[0][j] - c.lw * v[1][j] + hw * v[2][j] + c.lw * v[3][j];
          const float ww = -hh * v[0][j] + hh * v[1][j] - c.lh * v[2][j] + c.lh * v[3][j];
          gh = fmaf(wh, gc[j], gh);
          gw = fmaf(ww, gc[j], gw);
          gm = fmaf(gc[j], c.wgt[0] * v[0][j] + c.wgt[1] * v[1][j] + c.wgt[2] * v[2][j] + c.wgt[3] * v[3][j], gm);
        }
        if (grad_x) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float wk = c.wgt[k] * m;
            if (c.ok[k] && wk != 0.f) {
              float* d = grad_x + c.off[k] + (long long)c8 * 8;
              atomicAdd(reinterpret_cast<float4*>(d), make_float4(wk * gc[0], wk * gc[1], wk * gc[2], wk * gc[3]));
              atomicAdd(reinterpret_cast<float4*>(d + 4), make_float4(wk * gc[4], wk * gc[5], wk * gc[6], wk * gc[7]));
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gh += __shfl_xor_sync(0xffffffffu, gh, o);
        gw += __shfl_xor_sync(0xffffffffu, gw, o);
        gm += __shfl_xor_sync(0xffffffffu, gm, o);
      }
    }
    if (lane == 0) {
      float* go = grad_om + pix * g.oc_pitch;
      go[2 * tap] = gh * m;
      go[2 * tap + 1] = gw * m;
      if (g.modulated) go[2 * taps + tap] = gm * ms * (1.f - ms);      // through the sigmoid
    }
  }
}

static int dcn_nhwc_check(const DcnNhwc& g) {
  if (g.N < 0 || g.H <= 0 || g.W <= 0 || g.C <= 0 || g.Ho <= 0 || g.Wo <= 0 || g.kh <= 0 || g.kw <= 0 || g.stride <= 0) return MRB_ERR_BAD_ARG;
  if (g.C % 8) return MRB_ERR_UNSUPPORTED;
  if (g.oc_pitch < g.kh * g.kw * (g.modulated ? 3 : 2)) return MRB_ERR_BAD_ARG;
  return MRB_OK;
}

}  // namespace mrb
using namespace mrb;

MRB_API int mrb_dcn_sample_nhwc(const void* input_bf16, const float* offset_mask, void* columns_bf16, int batch, int height,
                                int width, int channels, int out_h, int out_w, int kh, int kw, int stride, int pad, int dilation,
                                int oc_pitch, int modulated, mrb_stream_t stream) {
  DcnNhwc g{batch, height, width, channels, out_h, out_w, kh, kw, stride, pad, dilation, oc_pitch, modulated};
  int rc = dcn_nhwc_check(g);
  if (rc) return rc;
  if (batch == 0) return MRB_OK;
  if (!input_bf16 || !offset_mask || !columns_bf16 || ((uintptr_t)input_bf16 & 15) || ((uintptr_t)columns_bf16 & 15)) return MRB_ERR_BAD_ARG;
  const long long items = (long long)batch * out_h * out_w * kh * kw;
  long long ctas = (items + 7) / 8;
  const long long cap = (long long)kNumSMs * 8 * 4;
  if (ctas > cap) ctas = cap;
  dcn_sample_nhwc_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(g, (const __nv_bfloat16*)input_bf16, offset_mask,
                                                                          (__nv_bfloat16*)columns_bf16);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_dcn_backward_nhwc(const void* input_bf16, const float* offset_mask, const void* grad_columns_bf16,
                                  float* grad_input_f32, float* grad_offset_mask, int batch, int height, int width, int channels,
                                  int out_h, int out_w, int kh, int kw, int stride, int pad, int dilation, int oc_pitch,
                                  int modulated, mrb_stream_t stream) {
  DcnNhwc g{batch, height, width, channels, out_h, out_w, kh, kw, stride, pad, dilation, oc_pitch, modulated};
  int rc = dcn_nhwc_check(g);
  if (rc) return rc;
  if (batch == 0) return MRB_OK;
  if (!input_bf16 || !offset_mask || !grad_columns_bf16 || !grad_offset_mask) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)input_bf16 & 15) || ((uintptr_t)grad_columns_bf16 & 15) || ((uintptr_t)grad_input_f32 & 15)) return MRB_ERR_BAD_ARG;
  const long long items = (long long)batch * out_h * out_w * kh * kw;
  long long ctas = (items + 7) / 8;
  const long long cap = (long long)kNumSMs * 8 * 4;
  if (ctas > cap) ctas = cap;
  dcn_backward_nhwc_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(g, (const __nv_bfloat16*)input_bf16, offset_mask,
                                                                            (const __nv_bfloat16*)grad_columns_bf16, grad_input_f32,
                                                                            grad_offset_mask);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
