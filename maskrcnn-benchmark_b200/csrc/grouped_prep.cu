// grouped_prep.cu -- operand preparation for the grouped (MRB_CONV_GROUPED64) convolutions of ResNeXt bodies
// (reference modeling/backbone/resnet.py:302-311, `groups=num_groups`): the block-diagonal expansion of a grouped filter
// for the forward and data-gradient launches, and the collapse of the expanded weight gradient back to the grouped layout.
// Three elementwise kernels replacing ~20 small PyTorch launches (arange / scatter / gather / permute / cast) per layer
// and step; a 104-layer X-101-32x8d step has 100 grouped 3x3 layers.
//
//   w      : [C][taps][Cg] bf16 (KRSC of the grouped filter [C, Cg, kh, kw]), Cg = C / groups, Cg | 64
//   w_exp  : [C][taps][64] bf16: row co holds its filter at columns [l0, l0 + Cg), l0 = (co / Cg * Cg) % 64, zeros elsewhere
//   wd_exp : [C][taps][64] bf16 (rows = INPUT channels): wd_exp[ci][t'][k] = w[co][taps-1-t'][ci - g*Cg] * scale[co] with
//            co = 64*(ci/64) + k when co and ci share the group g, else 0      (flipped taps: data gradient)
//   gw128  : [C][taps][128] fp32, what conv_wgrad_tc_kernel produces in grouped mode (row co against the 128 input
//            channels of its Cout tile); grad[co][cg][t] (+)= gw128[co][t][64*((co % 128) / 64) + l0 + cg]
#include <cuda_bf16.h>

#include "common.cuh"

namespace mrb {

__global__ void __launch_bounds__(256)
grouped_expand_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ scale, __nv_bfloat16* __restrict__ w_exp,
                      __nv_bfloat16* __restrict__ wd_exp, int C, int taps, int cg) {
  const long long total = (long long)C * taps * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 63);
    const int t = (int)((i >> 6) % taps);
    const int r = (int)((i >> 6) / taps);            // row: output channel for w_exp, input channel for wd_exp
    if (w_exp) {
      const int l0 = (r / cg * cg) & 63;
      const int q = j - l0;
      w_exp[i] = (q >= 0 && q < cg) ? w[((long long)r * taps + t) * cg + q] : __float2bfloat16_rn(0.f);
    }
    if (wd_exp) {
      const int co = (r & ~63) + j;                  // output channel of column j in input channel r's super-group
      float v = 0.f;
      if (co / cg == r / cg) {
        v = __bfloat162float(w[((long long)co * taps + (taps - 1 - t)) * cg + (r - r / cg * cg)]);
        if (scale) v *= __ldg(scale + co);
      }
      wd_exp[i] = __float2bfloat16_rn(v);
    }
  }
}

// grad strides: element strides of the destination [C, Cg, kh, kw] tensor (co, cg, tap): KRSC or NCHW-contiguous
__global__ void __launch_bounds__(256)
grouped_collapse_wgrad_kernel(const float* __restrict__ gw128, float* __restrict__ grad, int C, int taps, int cg, long long s_co,
                              long long s_cg, long long s_tap, int accumulate) {
  const long long total = (long long)C * taps * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cg);
    const int t = (int)((i / cg) % taps);
    const int co = (int)(i / cg / taps);
    const int l0 = (co / cg * cg) & 63;
    const float v = gw128[((long long)co * taps + t) * 128 + (((co & 127) >> 6) << 6) + l0 + q];
    float* d = grad + co * s_co + q * s_cg + t * s_tap;
    if (accumulate) *d += v; else *d = v;
  }
}

}  // namespace mrb
using namespace mrb;

MRB_API int mrb_grouped_expand_weights(const void* weight_bf16, const float* scale, void* w_exp_bf16, void* wd_exp_bf16, int channels,
                                       int taps, int groups, mrb_stream_t stream) {
  if (!weight_bf16 || channels <= 0 || taps <= 0 || groups <= 0 || channels % groups || channels % 64) return MRB_ERR_BAD_ARG;
  const int cg = channels / groups;
  if (64 % cg) return MRB_ERR_UNSUPPORTED;
  if (!w_exp_bf16 && !wd_exp_bf16) return MRB_OK;
  const long long total = (long long)channels * taps * 64;
  grouped_expand_kernel<<<grid_for(total, 256, 8, 2), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)weight_bf16, scale,
                                                                                     (__nv_bfloat16*)w_exp_bf16, (__nv_bfloat16*)wd_exp_bf16,
                                                                                     channels, taps, cg);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_grouped_collapse_wgrad(const float* grad_expanded128, float* grad_weight, int channels, int taps, int groups,
                                       long long stride_co, long long stride_cg, long long stride_tap, int accumulate,
                                       mrb_stream_t stream) {
  if (!grad_expanded128 || !grad_weight || channels <= 0 || taps <= 0 || groups <= 0 || channels % groups || channels % 128) return MRB_ERR_BAD_ARG;
  const int cg = channels / groups;
  if (64 % cg) return MRB_ERR_UNSUPPORTED;
  const long long total = (long long)channels * taps * cg;
  grouped_collapse_wgrad_kernel<<<grid_for(total, 256, 8, 2), 256, 0, (cudaStream_t)stream>>>(grad_expanded128, grad_weight, channels, taps,
                                                                                             cg, stride_co, stride_cg, stride_tap, accumulate);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
