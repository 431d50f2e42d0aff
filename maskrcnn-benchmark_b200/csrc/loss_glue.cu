// loss_glue.cu -- the three loss stages of the Mask R-CNN train step, forward and backward, one or two launches each.
//
// In the reference each is a chain of indexing / elementwise / reduction ops on BoxList fields (20-40 launches forward and
// as many again in autograd's backward, plus dense index_put / slice gradients):
//   * RPNLossComputation.__call__ (modeling/rpn/loss.py:92-131): sampled objectness BCE-with-logits + smooth-L1
//     (beta = 1/9, size_average=False) over the sampled positives, both divided by the number of sampled anchors.
//     Here the kernel reads the sampled anchors' logits / deltas straight out of the RPN head's per-level NHWC outputs
//     [N, H, W, A + 4A] and the backward scatters into the (zeroed) dense gradients of those outputs.
//   * FastRCNNLossComputation.__call__ (modeling/roi_heads/box_head/loss.py:120-167): cross-entropy over the sampled
//     proposals + smooth-L1 (beta = 1) on the regression outputs of the labelled class of the positives, divided by the
//     number of sampled proposals.  Input: the predictor GEMM's [R, C + 4C] fp32 output; the backward writes its dense
//     gradient in one pass.
//   * MaskRCNNLossComputation.__call__ (modeling/roi_heads/mask_head/loss.py:100-133): BCE-with-logits between the mask
//     logits of the labelled class (mask_logits[positive_inds, labels_pos]) and the rasterised targets.  Fixed-shape form:
//     per-ROI mean over M x M, weighted mean over the ROIs (weight 1 on positives, 0 on padding).  Input: the logit conv's
//     bf16 NHWC output [R, M, M, Cpad]; the backward writes the dense bf16 gradient (zero except the labelled channel).
// Reductions run in a fixed order (deterministic).  No CPU path.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mrb {

constexpr int kLossThreads = 1024;
constexpr int kMaxRpnLevels = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// sum over the CTA in a fixed order (warp partials combined by thread 0); result valid in thread 0.  `ws`: 32 floats.
__device__ __forceinline__ float block_sum_fixed(float v, float* ws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) ws[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += ws[w];
  return s;
}

__device__ __forceinline__ float bce_logits(float x, float t) {   // max(x, 0) - x t + log(1 + exp(-|x|))
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float smooth_l1(float d, float beta) {   // layers/smooth_l1_loss.py:5-16
  const float n = fabsf(d);
  return n < beta ? 0.5f * n * n / beta : n - 0.5f * beta;
}
__device__ __forceinline__ float smooth_l1_grad(float d, float beta) {
  const float n = fabsf(d);
  return n < beta ? d / beta : (d > 0.f ? 1.f : -1.f);
}

// ------------------------------------------------------------------------------------------ RPN loss
struct RpnLevels {
  int L, N, A, ld;                    // levels, images, anchors per location, floats per location (>= 5A: A logits, 4A deltas, padding)
  int off[kMaxRpnLevels + 1];         // first anchor index of every level in the concatenated anchor list
  int hw[kMaxRpnLevels];              // locations per image
  float* out[kMaxRpnLevels];          // [N, hw, 5A] fp32: A logits then 4A deltas per location (forward: read, backward: gradient)
};

__device__ __forceinline__ size_t rpn_locate(const RpnLevels& lv, int image, int anchor, int* level) {
  int l = 0;
  while (l + 1 < lv.L && anchor >= lv.off[l + 1]) ++l;
  const int local = anchor - lv.off[l];
  const int pix = local / lv.A, a = local - pix * lv.A;
  *level = l;
  return ((size_t)image * lv.hw[l] + pix) * (size_t)lv.ld + a;     // logit; deltas at + (A - a) + 4 a
}

// sel_idx / sel_label / sel_weight: [N, S] sampled anchors (weight 0 = padding); pos_idx / pos_ok / reg_targets: [N, P]
// result: {objectness loss, box loss, number of sampled anchors}
__global__ void __launch_bounds__(kLossThreads)
rpn_loss_fwd_kernel(RpnLevels lv, const int64_t* __restrict__ sel_idx, const float* __restrict__ sel_label,
                    const float* __restrict__ sel_weight, int S, const int64_t* __restrict__ pos_idx,
                    const unsigned char* __restrict__ pos_ok, const float4* __restrict__ reg_t, int P, float beta,
                    float* __restrict__ result) {
  __shared__ float ws[32];
  float obj = 0.f, num = 0.f, box = 0.f;
  for (int t = threadIdx.x; t < lv.N * S; t += blockDim.x) {
    const float w = sel_weight[t];
    if (w > 0.f) {
      int l;
      const size_t o = rpn_locate(lv, t / S, (int)sel_idx[t], &l);
      obj += bce_logits(lv.out[l][o], sel_label[t]);
      num += w;
    }
  }
  for (int t = threadIdx.x; t < lv.N * P; t += blockDim.x) {
    if (pos_ok[t]) {
      int l;
      const int anchor = (int)pos_idx[t];
      const size_t o = rpn_locate(lv, t / P, anchor, &l);
      const int a = (anchor - lv.off[l]) % lv.A;
      const float* d = lv.out[l] + o + (lv.A - a) + 4 * a;
      const float4 tg = reg_t[t];
      box += smooth_l1(d[0] - tg.x, beta) + smooth_l1(d[1] - tg.y, beta) + smooth_l1(d[2] - tg.z, beta) +
             smooth_l1(d[3] - tg.w, beta);
    }
  }
  const float so = block_sum_fixed(obj, ws);
  const float sn = block_sum_fixed(num, ws);
  const float sb = block_sum_fixed(box, ws);
  if (threadIdx.x == 0) {
    const float n = fmaxf(sn, 1.f);
    result[0] = so / n;
    result[1] = sb / n;
    result[2] = n;
  }
}

// gradients into the ZEROED dense per-level tensors `gr`; g_obj / g_box: upstream gradients of the two losses (device scalars)
__global__ void __launch_bounds__(256)
rpn_loss_bwd_kernel(RpnLevels fw, RpnLevels gr, const int64_t* __restrict__ sel_idx, const float* __restrict__ sel_label,
                    const float* __restrict__ sel_weight, int S, const int64_t* __restrict__ pos_idx,
                    const unsigned char* __restrict__ pos_ok, const float4* __restrict__ reg_t, int P, float beta,
                    const float* __restrict__ result, const float* __restrict__ g_obj, const float* __restrict__ g_box) {
  const float inv_n = 1.f / result[2];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < fw.N * S) {
    if (sel_weight[t] > 0.f) {
      int l;
      const size_t o = rpn_locate(fw, t / S, (int)sel_idx[t], &l);
      gr.out[l][o] = g_obj[0] * inv_n * (sigmoidf_(fw.out[l][o]) - sel_label[t]);
    }
  } else if (t < fw.N * (S + P)) {
    const int u = t - fw.N * S;
    if (pos_ok[u]) {
      int l;
      const int anchor = (int)pos_idx[u];
      const size_t o = rpn_locate(fw, u / P, anchor, &l);
      const int a = (anchor - fw.off[l]) % fw.A;
      const size_t d = o + (fw.A - a) + 4 * a;
      const float4 tg = reg_t[u];
      const float* x = fw.out[l] + d;
      float* g = gr.out[l] + d;
      const float s = g_box[0] * inv_n;
      g[0] = s * smooth_l1_grad(x[0] - tg.x, beta);
      g[1] = s * smooth_l1_grad(x[1] - tg.y, beta);
      g[2] = s * smooth_l1_grad(x[2] - tg.z, beta);
      g[3] = s * smooth_l1_grad(x[3] - tg.w, beta);
    }
  }
}

// ------------------------------------------------------------------------------------------ box head loss
// o [R, ld]: class logits in columns [0, C), regression outputs in [C + 4 c, C + 4 c + 4); labels [R] (-1 = not sampled)
// result: {classification loss, box loss, number of sampled rows}
__global__ void __launch_bounds__(256)
box_loss_rows_kernel(const float* __restrict__ o, int ld, int C, const int64_t* __restrict__ labels,
                     const float4* __restrict__ reg_t, int R, float* __restrict__ rows) {   // rows [R][3]: cls, box, sampled
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int64_t lab = labels[r];
  float cls = 0.f, box = 0.f, cnt = 0.f;
  if (lab >= 0) {
    const float* row = o + (size_t)r * ld;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    if (lane == 0) {
      cls = (m + logf(s)) - row[lab];
      cnt = 1.f;
      if (lab > 0) {
        const float* p = row + C + 4 * lab;
        const float4 tg = reg_t[r];
        box = smooth_l1(p[0] - tg.x, 1.f) + smooth_l1(p[1] - tg.y, 1.f) + smooth_l1(p[2] - tg.z, 1.f) +
              smooth_l1(p[3] - tg.w, 1.f);
      }
    }
  }
  if (lane == 0) {
    rows[(size_t)r * 3 + 0] = cls;
    rows[(size_t)r * 3 + 1] = box;
    rows[(size_t)r * 3 + 2] = cnt;
  }
}

__global__ void __launch_bounds__(kLossThreads)
box_loss_sum_kernel(const float* __restrict__ rows, int R, float* __restrict__ result) {
  __shared__ float ws[32];
  float c = 0.f, b = 0.f, n = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    c += rows[(size_t)r * 3 + 0];
    b += rows[(size_t)r * 3 + 1];
    n += rows[(size_t)r * 3 + 2];
  }
  const float sc = block_sum_fixed(c, ws);
  const float sb = block_sum_fixed(b, ws);
  const float sn = block_sum_fixed(n, ws);
  if (threadIdx.x == 0) {
    result[0] = sc / sn;                 // F.cross_entropy(ignore_index=-1): mean over the sampled rows
    result[1] = sb / fmaxf(sn, 1.f);     // loss.py:160-165: / labels.numel() (here: the sampled rows)
    result[2] = sn;
  }
}

__global__ void __launch_bounds__(256)
box_loss_bwd_kernel(const float* __restrict__ o, int ld, int C, const int64_t* __restrict__ labels,
                    const float4* __restrict__ reg_t, int R, const float* __restrict__ result, const float* __restrict__ g_cls,
                    const float* __restrict__ g_box, float* __restrict__ d_o) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int64_t lab = labels[r];
  const float* row = o + (size_t)r * ld;
  float* g = d_o + (size_t)r * ld;
  const float n = result[2];
  if (lab < 0) {
    for (int c = lane; c < ld; c += 32) g[c] = 0.f;
    return;
  }
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
  s = warp_sum(s);
  const float gc = g_cls[0] / n, gb = g_box[0] / fmaxf(n, 1.f);
  for (int c = lane; c < C; c += 32) g[c] = gc * (expf(row[c] - m) / s - (c == lab ? 1.f : 0.f));
  for (int c = C + lane; c < ld; c += 32) {
    float v = 0.f;
    const int k = c - C;
    if (lab > 0 && (k >> 2) == lab) {
      const float4 tg = reg_t[r];
      const float t = (k & 3) == 0 ? tg.x : ((k & 3) == 1 ? tg.y : ((k & 3) == 2 ? tg.z : tg.w));
      v = gb * smooth_l1_grad(row[c] - t, 1.f);
    }
    g[c] = v;
  }
}

// ------------------------------------------------------------------------------------------ mask loss
// y [R, M*M, C] bf16; labels [R]; targets [R, M*M] fp32; weights [R].  rows[r] = mean BCE of ROI r.
__global__ void __launch_bounds__(256)
mask_loss_rows_kernel(const __nv_bfloat16* __restrict__ y, int C, int MM, const int64_t* __restrict__ labels,
                      const float* __restrict__ targets, float* __restrict__ rows) {
  __shared__ float ws[32];
  const int r = blockIdx.x;
  const int lab = (int)labels[r];
  const __nv_bfloat16* p = y + (size_t)r * MM * C + lab;
  float s = 0.f;
  for (int i = threadIdx.x; i < MM; i += blockDim.x) s += bce_logits(__bfloat162float(p[(size_t)i * C]), targets[(size_t)r * MM + i]);
  s = block_sum_fixed(s, ws);
  if (threadIdx.x == 0) rows[r] = s / (float)MM;
}

// result: {loss, sum of weights clamped at 1}
__global__ void __launch_bounds__(kLossThreads)
weighted_mean_kernel(const float* __restrict__ rows, const float* __restrict__ weights, int R, float* __restrict__ result) {
  __shared__ float ws[32];
  float s = 0.f, w = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float wr = weights[r];
    if (wr > 0.f) {
      s += rows[r];
      w += wr;
    }
  }
  const float ss = block_sum_fixed(s, ws);
  const float sw = block_sum_fixed(w, ws);
  if (threadIdx.x == 0) {
    const float n = fmaxf(sw, 1.f);
    result[0] = ss / n;
    result[1] = n;
  }
}

// dense gradient [R, M*M, C] bf16: zero except the labelled channel.  One thread per (ROI, pixel), 16-byte stores.
__global__ void __launch_bounds__(256)
mask_loss_bwd_kernel(const __nv_bfloat16* __restrict__ y, int C, int MM, const int64_t* __restrict__ labels,
                     const float* __restrict__ targets, const float* __restrict__ weights, int R,
                     const float* __restrict__ result, const float* __restrict__ g, __nv_bfloat16* __restrict__ d_y) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * MM) return;
  const int r = (int)(t / MM);
  const int lab = (int)labels[r];
  const float w = weights[r];
  float v = 0.f;
  if (w > 0.f) v = g[0] / (result[1] * (float)MM) * (sigmoidf_(__bfloat162float(y[(size_t)t * C + lab])) - targets[t]);
  const __nv_bfloat16 bv = __float2bfloat16(v);
  uint4* out = reinterpret_cast<uint4*>(d_y + (size_t)t * C);
  const int lab_vec = lab >> 3, lab_in = lab & 7;
  for (int q = 0; q < (C >> 3); ++q) {
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (q == lab_vec) {
      unsigned short bits = __bfloat16_as_ushort(bv);
      unsigned* zw = reinterpret_cast<unsigned*>(&z);
      zw[lab_in >> 1] = (lab_in & 1) ? ((unsigned)bits << 16) : (unsigned)bits;
    }
    out[q] = z;
  }
}

static int fill_levels(RpnLevels* lv, void* const* outs_host, const int* hw_host, int num_levels, int num_images,
                       int anchors_per_location, int pixel_stride) {
  if (num_levels <= 0 || num_levels > kMaxRpnLevels || num_images <= 0 || anchors_per_location <= 0 || !outs_host || !hw_host ||
      pixel_stride < 5 * anchors_per_location)
    return MRB_ERR_BAD_ARG;
  lv->ld = pixel_stride;
  lv->L = num_levels;
  lv->N = num_images;
  lv->A = anchors_per_location;
  int off = 0;
  for (int l = 0; l < kMaxRpnLevels; ++l) {
    lv->off[l] = off;
    lv->hw[l] = l < num_levels ? hw_host[l] : 0;
    lv->out[l] = l < num_levels ? (float*)outs_host[l] : nullptr;
    if (l < num_levels) {
      if (!outs_host[l] || hw_host[l] <= 0) return MRB_ERR_BAD_ARG;
      off += hw_host[l] * anchors_per_location;
    }
  }
  lv->off[kMaxRpnLevels] = off;
  return MRB_OK;
}

}  // namespace mrb
using namespace mrb;

MRB_API int mrb_rpn_loss_fwd(void* const* head_outputs_host, const int* locations_host, int num_levels, int num_images,
                             int anchors_per_location, int pixel_stride, const int64_t* sel_idx, const float* sel_label, const float* sel_weight,
                             int num_sel, const int64_t* pos_idx, const uint8_t* pos_ok, const float* reg_targets, int num_pos,
                             float beta, float* result, mrb_stream_t stream) {
  RpnLevels lv;
  const int rc = fill_levels(&lv, head_outputs_host, locations_host, num_levels, num_images, anchors_per_location, pixel_stride);
  if (rc != MRB_OK) return rc;
  if (!sel_idx || !sel_label || !sel_weight || !pos_idx || !pos_ok || !reg_targets || !result || num_sel <= 0 || num_pos <= 0)
    return MRB_ERR_BAD_ARG;
  if ((uintptr_t)reg_targets & 15) return MRB_ERR_BAD_ARG;
  rpn_loss_fwd_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(lv, sel_idx, sel_label, sel_weight, num_sel, pos_idx, pos_ok,
                                                                     (const float4*)reg_targets, num_pos, beta, result);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_rpn_loss_bwd(void* const* head_outputs_host, void* const* grad_outputs_host, const int* locations_host,
                             int num_levels, int num_images, int anchors_per_location, int pixel_stride, const int64_t* sel_idx,
                             const float* sel_label, const float* sel_weight, int num_sel, const int64_t* pos_idx,
                             const uint8_t* pos_ok, const float* reg_targets, int num_pos, float beta, const float* result,
                             const float* grad_objectness, const float* grad_box, mrb_stream_t stream) {
  RpnLevels fw, gr;
  int rc = fill_levels(&fw, head_outputs_host, locations_host, num_levels, num_images, anchors_per_location, pixel_stride);
  if (rc != MRB_OK) return rc;
  rc = fill_levels(&gr, grad_outputs_host, locations_host, num_levels, num_images, anchors_per_location, pixel_stride);
  if (rc != MRB_OK) return rc;
  if (!sel_idx || !sel_label || !sel_weight || !pos_idx || !pos_ok || !reg_targets || !result || !grad_objectness || !grad_box ||
      num_sel <= 0 || num_pos <= 0)
    return MRB_ERR_BAD_ARG;
  const int total = num_images * (num_sel + num_pos);
  rpn_loss_bwd_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(fw, gr, sel_idx, sel_label, sel_weight, num_sel, pos_idx,
                                                                               pos_ok, (const float4*)reg_targets, num_pos, beta,
                                                                               result, grad_objectness, grad_box);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_box_loss_fwd(const float* outputs, int ld, int num_classes, const int64_t* labels, const float* reg_targets,
                             int num_rois, float* row_scratch, float* result, mrb_stream_t stream) {
  if (!outputs || !labels || !reg_targets || !row_scratch || !result || num_rois <= 0 || num_classes <= 0 || ld < 5 * num_classes)
    return MRB_ERR_BAD_ARG;
  if ((uintptr_t)reg_targets & 15) return MRB_ERR_BAD_ARG;
  box_loss_rows_kernel<<<ceil_div(num_rois, 8), 256, 0, (cudaStream_t)stream>>>(outputs, ld, num_classes, labels,
                                                                                 (const float4*)reg_targets, num_rois, row_scratch);
  MRB_LAUNCH_CHECK();
  box_loss_sum_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(row_scratch, num_rois, result);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_box_loss_bwd(const float* outputs, int ld, int num_classes, const int64_t* labels, const float* reg_targets,
                             int num_rois, const float* result, const float* grad_cls, const float* grad_box, float* grad_outputs,
                             mrb_stream_t stream) {
  if (!outputs || !labels || !reg_targets || !result || !grad_cls || !grad_box || !grad_outputs || num_rois <= 0 ||
      num_classes <= 0 || ld < 5 * num_classes)
    return MRB_ERR_BAD_ARG;
  box_loss_bwd_kernel<<<ceil_div(num_rois, 8), 256, 0, (cudaStream_t)stream>>>(outputs, ld, num_classes, labels,
                                                                                (const float4*)reg_targets, num_rois, result,
                                                                                grad_cls, grad_box, grad_outputs);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_mask_loss_fwd(const void* logits_bf16, int channels, int mask_pixels, const int64_t* labels, const float* targets,
                              const float* weights, int num_rois, float* row_losses, float* result, mrb_stream_t stream) {
  if (!logits_bf16 || !labels || !targets || !weights || !row_losses || !result || num_rois <= 0 || channels <= 0 ||
      mask_pixels <= 0)
    return MRB_ERR_BAD_ARG;
  mask_loss_rows_kernel<<<num_rois, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits_bf16, channels, mask_pixels, labels,
                                                                    targets, row_losses);
  MRB_LAUNCH_CHECK();
  weighted_mean_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(row_losses, weights, num_rois, result);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_mask_loss_bwd(const void* logits_bf16, int channels, int mask_pixels, const int64_t* labels, const float* targets,
                              const float* weights, int num_rois, const float* result, const float* grad_loss,
                              void* grad_logits_bf16, mrb_stream_t stream) {
  if (!logits_bf16 || !labels || !targets || !weights || !result || !grad_loss || !grad_logits_bf16 || num_rois <= 0 ||
      channels <= 0 || (channels & 7) || mask_pixels <= 0)
    return MRB_ERR_BAD_ARG;
  if ((uintptr_t)grad_logits_bf16 & 15) return MRB_ERR_BAD_ARG;
  const long long total = (long long)num_rois * mask_pixels;
  mask_loss_bwd_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)logits_bf16, channels, mask_pixels, labels, targets, weights, num_rois, result, grad_loss,
      (__nv_bfloat16*)grad_logits_bf16);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
