// focal_loss.cu -- SigmoidFocalLoss forward/backward for sm_100a.
// Replaces SigmoidFocalLossForward / SigmoidFocalLossBackward (reference
// csrc/cuda/SigmoidFocalLoss_cuda.cu:20-101).  Pure streaming op: 2 (fwd) / 3 (bwd) fp32 words
// per logit, so the kernel is written to be HBM-bound rather than SFU-bound: one exp, one
// log1p and one reciprocal per element (the reference spends 2 expf + 2 logf + 2 powf),
// 128-bit loads/stores, one target load per 4 logits.  With e = exp(-|x|):
//     p        = sigmoid(x)
//     log p    = min(x,0) - log1p(e)      (clamped at log(FLT_MIN) like logf(max(p,FLT_MIN)))
//     log(1-p) = -max(x,0) - log1p(e)     (== the reference's stable "-x*(x>=0) - log(1+exp(x-2x(x>=0)))")
//     (1-p)^g  = exp(g*log(1-p)),  p^g = exp(g*log p)   (g == 2 -> a multiply)
#include <cfloat>
#include "common.cuh"

namespace mrb {

constexpr int kFlThreads = 256;
constexpr float kLogFltMin = -87.33654475f;  // logf(FLT_MIN)

struct FocalTerms { float p, log_p, log_1mp, pow_1mp, pow_p; };

__device__ __forceinline__ FocalTerms focal_terms(float x, float gamma, bool gamma_is_2) {
  FocalTerms t;
  const float e = __expf(-fabsf(x));
  const float inv = __fdividef(1.f, 1.f + e);
  t.p = (x >= 0.f) ? inv : e * inv;
  const float l1p = log1pf(e);
  t.log_p = fmaxf(fminf(x, 0.f) - l1p, kLogFltMin);
  t.log_1mp = -fmaxf(x, 0.f) - l1p;
  const float omp = (x >= 0.f) ? e * inv : inv;  // 1 - p without cancellation
  if (gamma_is_2) {
    t.pow_1mp = omp * omp;
    t.pow_p = t.p * t.p;
  } else {
    t.pow_1mp = __expf(gamma * t.log_1mp);
    t.pow_p = __expf(gamma * (fminf(x, 0.f) - l1p));
  }
  return t;
}

__device__ __forceinline__ float focal_fwd_one(float x, int t, int d, float gamma, float alpha, bool g2) {
  const float c1 = (t == d + 1) ? 1.f : 0.f;
  const float c2 = (t >= 0 && t != d + 1) ? 1.f : 0.f;
  const FocalTerms f = focal_terms(x, gamma, g2);
  const float term1 = f.pow_1mp * f.log_p;
  const float term2 = f.pow_p * f.log_1mp;
  return -c1 * term1 * alpha - c2 * term2 * (1.f - alpha);
}

__device__ __forceinline__ float focal_bwd_one(float x, int t, int d, float gamma, float alpha, bool g2, float dl) {
  const float c1 = (t == d + 1) ? 1.f : 0.f;
  const float c2 = (t >= 0 && t != d + 1) ? 1.f : 0.f;
  const FocalTerms f = focal_terms(x, gamma, g2);
  const float omp = 1.f - f.p;
  // SigmoidFocalLoss_cuda.cu:86-93
  const float term1 = f.pow_1mp * (omp - f.p * gamma * f.log_p);
  const float term2 = f.pow_p * (f.log_1mp * omp * gamma - f.p);
  return (-c1 * term1 * alpha - c2 * term2 * (1.f - alpha)) * dl;
}

template <bool BWD>
__global__ void __launch_bounds__(kFlThreads)
focal_vec4_kernel(const float4* __restrict__ logits, const int32_t* __restrict__ targets,
                  const float4* __restrict__ d_losses, float4* __restrict__ out, int64_t n_vec, int num_classes,
                  float gamma, float alpha) {
  const bool g2 = (gamma == 2.f);
  const int vec_per_row = num_classes >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / vec_per_row;
    const int d0 = (int)(i - row * vec_per_row) << 2;
    const int t = __ldg(targets + row);
    const float4 x = __ldcs(logits + i);
    float4 r;
    if (BWD) {
      const float4 dl = __ldcs(d_losses + i);
      r.x = focal_bwd_one(x.x, t, d0 + 0, gamma, alpha, g2, dl.x);
      r.y = focal_bwd_one(x.y, t, d0 + 1, gamma, alpha, g2, dl.y);
      r.z = focal_bwd_one(x.z, t, d0 + 2, gamma, alpha, g2, dl.z);
      r.w = focal_bwd_one(x.w, t, d0 + 3, gamma, alpha, g2, dl.w);
    } else {
      r.x = focal_fwd_one(x.x, t, d0 + 0, gamma, alpha, g2);
      r.y = focal_fwd_one(x.y, t, d0 + 1, gamma, alpha, g2);
      r.z = focal_fwd_one(x.z, t, d0 + 2, gamma, alpha, g2);
      r.w = focal_fwd_one(x.w, t, d0 + 3, gamma, alpha, g2);
    }
    __stcs(out + i, r);
  }
}

template <bool BWD>
__global__ void __launch_bounds__(kFlThreads)
focal_scalar_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                    const float* __restrict__ d_losses, float* __restrict__ out, int64_t n, int num_classes,
                    float gamma, float alpha) {
  const bool g2 = (gamma == 2.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / num_classes;
    const int d = (int)(i - row * num_classes);
    const int t = __ldg(targets + row);
    out[i] = BWD ? focal_bwd_one(logits[i], t, d, gamma, alpha, g2, d_losses[i])
                 : focal_fwd_one(logits[i], t, d, gamma, alpha, g2);
  }
}

template <bool BWD>
static int focal_launch(const float* logits, const int32_t* targets, const float* d_losses, float* out,
                        int64_t A, int num_classes, float gamma, float alpha, cudaStream_t stream) {
  if (A < 0 || num_classes <= 0) return MRB_ERR_BAD_ARG;
  const int64_t n = A * num_classes;
  if (n == 0) return MRB_OK;
  if (!logits || !targets || !out || (BWD && !d_losses)) return MRB_ERR_BAD_ARG;
  const bool aligned = (((uintptr_t)logits | (uintptr_t)out | (uintptr_t)(BWD ? d_losses : logits)) & 15) == 0;
  if (num_classes % 4 == 0 && aligned) {
    const int64_t nv = n >> 2;
    const int grid = grid_for(nv, kFlThreads, 8, 2);
    focal_vec4_kernel<BWD><<<grid, kFlThreads, 0, stream>>>((const float4*)logits, targets, (const float4*)d_losses,
                                                           (float4*)out, nv, num_classes, gamma, alpha);
  } else {
    const int grid = grid_for(n, kFlThreads, 8, 4);
    focal_scalar_kernel<BWD><<<grid, kFlThreads, 0, stream>>>(logits, targets, d_losses, out, n, num_classes, gamma, alpha);
  }
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
}  // namespace mrb
using namespace mrb;

MRB_API int mrb_sigmoid_focal_fwd(const float* logits, const int32_t* targets, float* losses, int64_t num_anchors,
                                  int num_classes, float gamma, float alpha, mrb_stream_t stream) {
  return focal_launch<false>(logits, targets, nullptr, losses, num_anchors, num_classes, gamma, alpha, (cudaStream_t)stream);
}

MRB_API int mrb_sigmoid_focal_bwd(const float* logits, const int32_t* targets, const float* d_losses, float* d_logits,
                                  int64_t num_anchors, int num_classes, float gamma, float alpha, mrb_stream_t stream) {
  return focal_launch<true>(logits, targets, d_losses, d_logits, num_anchors, num_classes, gamma, alpha, (cudaStream_t)stream);
}
