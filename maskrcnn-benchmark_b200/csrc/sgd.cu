// sgd.cu -- the parameter update of the train step as ONE streaming pass (HBM bound, 26 B / parameter).
//
// Reference semantics: torch.optim.SGD(momentum, weight_decay) as configured by solver/build.py:7-20
//   d = g * grad_scale + wd * p;   m = momentum * m + d;   p = p - lr * m
// fused with the two passes that otherwise surround it in a bf16 tensor-core step: the bf16 operand copy of
// the updated weights (what the convolution engine reads next step) and the zeroing of the gradient
// accumulator (the weight-gradient kernels red.add straight into it).
#include "common.cuh"
#include <cuda_bf16.h>

namespace mrb {

__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, __nv_bfloat16* __restrict__ w16,
                    long long n, float lr, float momentum, float wd, float gscale, int zero_grad) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    mv.x = momentum * mv.x + (gv.x * gscale + wd * pv.x);
    mv.y = momentum * mv.y + (gv.y * gscale + wd * pv.y);
    mv.z = momentum * mv.z + (gv.z * gscale + wd * pv.z);
    mv.w = momentum * mv.w + (gv.w * gscale + wd * pv.w);
    pv.x -= lr * mv.x; pv.y -= lr * mv.y; pv.z -= lr * mv.z; pv.w -= lr * mv.w;
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(w16)[i] = o;
    }
  }
  // ragged tail (n % 4 elements)
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float pv = p[t], mv = momentum * m[t] + (g[t] * gscale + wd * pv);
    pv -= lr * mv;
    p[t] = pv; m[t] = mv;
    if (zero_grad) g[t] = 0.f;
    if (w16) w16[t] = __float2bfloat16_rn(pv);
  }
}

}  // namespace mrb

MRB_API int mrb_sgd_momentum_step(float* param, float* grad, float* momentum_buf, void* param_bf16, long long n, float lr,
                                  float momentum, float weight_decay, float grad_scale, int zero_grad, mrb_stream_t stream_) {
  if (n < 0) return MRB_ERR_BAD_ARG;
  if (n == 0) return MRB_OK;
  if (!param || !grad || !momentum_buf) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum_buf) & 15) return MRB_ERR_UNSUPPORTED;
  if (param_bf16 && ((uintptr_t)param_bf16 & 7)) return MRB_ERR_UNSUPPORTED;
  const int grid = mrb::grid_for((n + 3) / 4, 256, 8, 2);
  mrb::sgd_momentum_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(param, grad, momentum_buf, (__nv_bfloat16*)param_bf16, n, lr,
                                                                     momentum, weight_decay, grad_scale, zero_grad);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
