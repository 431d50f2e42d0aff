// roi_pool.cu -- ROIPool forward/backward for sm_100a (NCHW fp32).
// Replaces RoIPoolFForward / RoIPoolFBackward (reference csrc/cuda/ROIPool_cuda.cu:16-108).
// Integer bins from rounded ROI coordinates, max + int32 argmax, empty bin -> (0, -1).
// One CTA per (ROI, channel slab): the bin rectangle table is computed once per CTA; lanes run
// over (ph, pw) bins of one plane so window reads of a warp stay inside one ROI footprint.
// The op has no caller in modeling/ (Pooler hard-codes ROIAlign, modeling/poolers.py:48,66); it is
// here for `_C` surface completeness.
#include <cfloat>
#include "common.cuh"

namespace mrb {

constexpr int kRpThreads = 256;

__global__ void __launch_bounds__(kRpThreads)
roi_pool_fwd_kernel(const float* __restrict__ input, const float* __restrict__ rois, float* __restrict__ output,
                    int32_t* __restrict__ argmax, int C, int H, int W, int PH, int PW, float scale, int slab) {
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * slab;
  const int cn = min(slab, C - c0);
  const float* roi = rois + (size_t)n * 5;
  const int b = (int)roi[0];
  // ROIPool_cuda.cu:30-41
  const int rsw = (int)roundf(__fmul_rn(roi[1], scale)), rsh = (int)roundf(__fmul_rn(roi[2], scale));
  const int rew = (int)roundf(__fmul_rn(roi[3], scale)), reh = (int)roundf(__fmul_rn(roi[4], scale));
  const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);
  const float bin_h = __fdiv_rn((float)rh, (float)PH), bin_w = __fdiv_rn((float)rw, (float)PW);
  const int PP = PH * PW;
  const size_t plane = (size_t)H * W;
  const float* __restrict__ src0 = input + ((size_t)b * C + c0) * plane;
  const size_t o0 = ((size_t)n * C + c0) * PP;
  for (int o = threadIdx.x; o < cn * PP; o += kRpThreads) {
    const int c = o / PP, bin = o - c * PP;
    const int ph = bin / PW, pw = bin - ph * PW;
    int hstart = (int)floorf(__fmul_rn((float)ph, bin_h)), wstart = (int)floorf(__fmul_rn((float)pw, bin_w));
    int hend = (int)ceilf(__fmul_rn((float)(ph + 1), bin_h)), wend = (int)ceilf(__fmul_rn((float)(pw + 1), bin_w));
    hstart = min(max(hstart + rsh, 0), H); hend = min(max(hend + rsh, 0), H);
    wstart = min(max(wstart + rsw, 0), W); wend = min(max(wend + rsw, 0), W);
    const bool is_empty = (hend <= hstart) || (wend <= wstart);
    float maxval = is_empty ? 0.f : -FLT_MAX;
    int maxidx = -1;
    const float* __restrict__ src = src0 + (size_t)c * plane;
    for (int h = hstart; h < hend; ++h)
      for (int w = wstart; w < wend; ++w) {
        const float v = __ldg(src + h * W + w);
        if (v > maxval) { maxval = v; maxidx = h * W + w; }
      }
    output[o0 + o] = maxval;
    argmax[o0 + o] = maxidx;
  }
}

__global__ void __launch_bounds__(kRpThreads)
roi_pool_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ rois,
                    const int32_t* __restrict__ argmax, float* __restrict__ gin, int C, int H, int W, int PP,
                    int slab) {
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * slab;
  const int cn = min(slab, C - c0);
  const int b = (int)rois[(size_t)n * 5];
  const size_t plane = (size_t)H * W;
  float* __restrict__ dst0 = gin + ((size_t)b * C + c0) * plane;
  const size_t o0 = ((size_t)n * C + c0) * PP;
  for (int o = threadIdx.x; o < cn * PP; o += kRpThreads) {
    const int c = o / PP;
    const int am = argmax[o0 + o];
    if (am != -1) atomicAdd(dst0 + (size_t)c * plane + am, grad[o0 + o]);
  }
}

static int rp_slab(int num_rois, int C) {
  int slab = C;
  while (slab > 16 && (int64_t)num_rois * ceil_div(C, slab) < (int64_t)kNumSMs * 8) slab = (slab + 1) / 2;
  return slab;
}
}  // namespace mrb
using namespace mrb;

MRB_API int mrb_roi_pool_fwd(const float* input, const float* rois, float* output, int32_t* argmax, int num_rois,
                             int batch, int channels, int height, int width, int pooled_h, int pooled_w,
                             float spatial_scale, mrb_stream_t stream) {
  if (num_rois < 0 || channels < 0 || pooled_h <= 0 || pooled_w <= 0 || height < 0 || width < 0) return MRB_ERR_BAD_ARG;
  if (num_rois == 0 || channels == 0) return MRB_OK;
  if (!input || !rois || !output || !argmax) return MRB_ERR_BAD_ARG;
  const int slab = rp_slab(num_rois, channels);
  dim3 grid(num_rois, ceil_div(channels, slab));
  roi_pool_fwd_kernel<<<grid, kRpThreads, 0, (cudaStream_t)stream>>>(input, rois, output, argmax, channels, height,
                                                                    width, pooled_h, pooled_w, spatial_scale, slab);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_roi_pool_bwd(const float* grad_output, const float* rois, const int32_t* argmax, float* grad_input,
                             int num_rois, int batch, int channels, int height, int width, int pooled_h,
                             int pooled_w, mrb_stream_t stream) {
  if (num_rois < 0 || channels < 0 || pooled_h <= 0 || pooled_w <= 0 || batch < 0) return MRB_ERR_BAD_ARG;
  const size_t total = (size_t)batch * channels * height * width;
  if (total == 0) return MRB_OK;
  if (!grad_input) return MRB_ERR_BAD_ARG;
  MRB_CUDA_TRY(cudaMemsetAsync(grad_input, 0, total * sizeof(float), (cudaStream_t)stream));
  if (num_rois == 0) return MRB_OK;
  if (!grad_output || !rois || !argmax) return MRB_ERR_BAD_ARG;
  const int slab = rp_slab(num_rois, channels);
  dim3 grid(num_rois, ceil_div(channels, slab));
  roi_pool_bwd_kernel<<<grid, kRpThreads, 0, (cudaStream_t)stream>>>(grad_output, rois, argmax, grad_input, channels,
                                                                    height, width, pooled_h * pooled_w, slab);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
