// pool.cu -- the two pooling passes of the ResNet+FPN path as streaming NHWC kernels (HBM bound).
//
//  * max pool (stem: 3x3 / stride 2 / pad 1, reference modeling/backbone/resnet.py:307; FPN P6: 1x1 / stride 2,
//    modeling/backbone/fpn.py:77-79).  Forward only and without an index tensor: the stem is frozen in every R-50
//    config and P6 is a pure subsample, so nothing ever needs the argmax.
//  * 2x2 sum pool: the backward of the nearest-2x upsample of FPN's top-down path (fpn.py:59-64): the gradient of the
//    coarse map is the sum over each 2x2 block of the fine map's gradient.
// One thread = 8 channels (one 16-byte word) of one output pixel; consecutive threads walk the channel axis, so every
// global access is a full coalesced line.  Algorithmic bytes: (input + output) * 2 B per element.
#include "common.cuh"
#include <cuda_bf16.h>

namespace mrb {

__global__ void __launch_bounds__(256)
max_pool_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int N, int H, int W, int C8, int k, int s, int p,
                     int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long t = i / C8;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const int h0 = ho * s - p, w0 = wo * s - p;
    const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
    __nv_bfloat162 m[4] = {ninf, ninf, ninf, ninf};
    for (int dy = 0; dy < k; ++dy) {
      const int h = h0 + dy;
      if ((unsigned)h >= (unsigned)H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int w = w0 + dx;
        if ((unsigned)w >= (unsigned)W) continue;
        const uint4 v = __ldg(x + ((long long)(n * H + h) * W + w) * C8 + c);
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2_nan(m[j], v2[j]);     // NaN propagates, as in ATen
      }
    }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = m[j];
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256)
sum_pool2x2_nhwc_kernel(const uint4* __restrict__ g, uint4* __restrict__ out, int N, int H, int W, int C8, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long t = i / C8;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int h = 2 * ho + dy;
      if (h >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int w = 2 * wo + dx;
        if (w >= W) continue;
        const uint4 v = __ldcs(g + ((long long)(n * H + h) * W + w) * C8 + c);
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(v2[j]);
          acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
        }
      }
    }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    out[i] = o;
  }
}

}  // namespace mrb

MRB_API int mrb_max_pool_nhwc(const void* input_bf16, void* output_bf16, int batch, int height, int width, int channels,
                              int kernel, int stride, int pad, mrb_stream_t stream) {
  if (batch < 0 || height <= 0 || width <= 0 || channels <= 0 || kernel <= 0 || stride <= 0 || pad < 0 || 2 * pad > kernel)
    return MRB_ERR_BAD_ARG;
  if (channels % 8) return MRB_ERR_UNSUPPORTED;
  const int Ho = (height + 2 * pad - kernel) / stride + 1, Wo = (width + 2 * pad - kernel) / stride + 1;
  if (Ho <= 0 || Wo <= 0) return MRB_ERR_BAD_ARG;
  if (batch == 0) return MRB_OK;
  if (!input_bf16 || !output_bf16 || (((uintptr_t)input_bf16 | (uintptr_t)output_bf16) & 15)) return MRB_ERR_BAD_ARG;
  const long long total = (long long)batch * Ho * Wo * (channels / 8);
  mrb::max_pool_nhwc_kernel<<<mrb::grid_for(total, 256, 8, 4), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)input_bf16, (uint4*)output_bf16, batch, height, width, channels / 8, kernel, stride, pad, Ho, Wo);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_sum_pool2x2_nhwc(const void* grad_bf16, void* out_bf16, int batch, int height, int width, int channels,
                                 mrb_stream_t stream) {
  if (batch < 0 || height <= 0 || width <= 0 || channels <= 0) return MRB_ERR_BAD_ARG;
  if (channels % 8) return MRB_ERR_UNSUPPORTED;
  if (batch == 0) return MRB_OK;
  if (!grad_bf16 || !out_bf16 || (((uintptr_t)grad_bf16 | (uintptr_t)out_bf16) & 15)) return MRB_ERR_BAD_ARG;
  const int Ho = (height + 1) / 2, Wo = (width + 1) / 2;
  const long long total = (long long)batch * Ho * Wo * (channels / 8);
  mrb::sum_pool2x2_nhwc_kernel<<<mrb::grid_for(total, 256, 8, 4), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)grad_bf16, (uint4*)out_bf16, batch, height, width, channels / 8, Ho, Wo);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
