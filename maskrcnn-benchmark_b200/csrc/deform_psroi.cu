// deform_psroi.cu -- deformable position-sensitive ROI pooling for sm_100a (NCHW fp32).
// Replaces DeformablePSROIPoolForwardKernel / DeformablePSROIPoolBackwardAccKernel (reference
// csrc/cuda/deform_pool_kernel_cuda.cu:53-264).  No model in the reference instantiates this op;
// it is provided so that the `_C` surface is complete.  One thread per (roi, ctop, ph, pw) bin,
// the geometry shared by forward and backward lives in PsBin.
#include "common.cuh"

namespace mrb {

struct PsArgs {
  int channels, height, width, pooled, output_dim, group_size, part_size, sample_per_part;
  int num_classes, channels_each_class, no_trans;
  float spatial_scale, trans_std;
};

struct PsBin {
  int n, ctop, ph, pw, batch, c, class_id, part_h, part_w;
  float wstart, hstart, sub_w, sub_h, roi_w, roi_h;
};

__device__ __forceinline__ PsBin ps_bin(int index, const PsArgs& a, const float* __restrict__ rois,
                                        const float* __restrict__ trans) {
  PsBin b;
  b.pw = index % a.pooled;
  b.ph = (index / a.pooled) % a.pooled;
  b.ctop = (index / a.pooled / a.pooled) % a.output_dim;
  b.n = index / a.pooled / a.pooled / a.output_dim;
  const float* r = rois + (size_t)b.n * 5;
  b.batch = (int)r[0];
  // deform_pool_kernel_cuda.cu:84-90
  const float rsw = roundf(r[1]) * a.spatial_scale - 0.5f, rsh = roundf(r[2]) * a.spatial_scale - 0.5f;
  const float rew = (roundf(r[3]) + 1.f) * a.spatial_scale - 0.5f, reh = (roundf(r[4]) + 1.f) * a.spatial_scale - 0.5f;
  b.roi_w = fmaxf(rew - rsw, 0.1f);
  b.roi_h = fmaxf(reh - rsh, 0.1f);
  const float bin_h = b.roi_h / (float)a.pooled, bin_w = b.roi_w / (float)a.pooled;
  b.sub_h = bin_h / (float)a.sample_per_part;
  b.sub_w = bin_w / (float)a.sample_per_part;
  b.part_h = (int)floorf((float)b.ph / a.pooled * a.part_size);
  b.part_w = (int)floorf((float)b.pw / a.pooled * a.part_size);
  b.class_id = b.ctop / a.channels_each_class;
  float tx = 0.f, ty = 0.f;
  if (!a.no_trans) {
    const size_t t0 = (((size_t)b.n * a.num_classes + b.class_id) * 2) * a.part_size;
    tx = trans[(t0 + b.part_h) * a.part_size + b.part_w] * a.trans_std;
    ty = trans[(t0 + a.part_size + b.part_h) * a.part_size + b.part_w] * a.trans_std;
  }
  b.wstart = (float)b.pw * bin_w + rsw + tx * b.roi_w;
  b.hstart = (float)b.ph * bin_h + rsh + ty * b.roi_h;
  int gw = (int)floorf((float)b.pw * a.group_size / a.pooled);
  int gh = (int)floorf((float)b.ph * a.group_size / a.pooled);
  gw = min(max(gw, 0), a.group_size - 1);
  gh = min(max(gh, 0), a.group_size - 1);
  b.c = (b.ctop * a.group_size + gh) * a.group_size + gw;
  return b;
}

__global__ void __launch_bounds__(256)
psroi_fwd_kernel(int count, PsArgs a, const float* __restrict__ data, const float* __restrict__ rois,
                 const float* __restrict__ trans, float* __restrict__ out, float* __restrict__ top_count) {
  for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < count; index += gridDim.x * blockDim.x) {
    const PsBin b = ps_bin(index, a, rois, trans);
    const float* __restrict__ plane = data + ((size_t)b.batch * a.channels + b.c) * a.height * a.width;
    float sum = 0.f;
    int cnt = 0;
    for (int ih = 0; ih < a.sample_per_part; ++ih)
      for (int iw = 0; iw < a.sample_per_part; ++iw) {
        float w = b.wstart + iw * b.sub_w, h = b.hstart + ih * b.sub_h;
        if (w < -0.5f || w > a.width - 0.5f || h < -0.5f || h > a.height - 0.5f) continue;
        w = fminf(fmaxf(w, 0.f), a.width - 1.f);
        h = fminf(fmaxf(h, 0.f), a.height - 1.f);
        const int x1 = (int)floorf(w), x2 = (int)ceilf(w), y1 = (int)floorf(h), y2 = (int)ceilf(h);
        const float dx = w - x1, dy = h - y1;
        sum += (1 - dx) * (1 - dy) * __ldg(plane + y1 * a.width + x1) + (1 - dx) * dy * __ldg(plane + y2 * a.width + x1) +
               dx * (1 - dy) * __ldg(plane + y1 * a.width + x2) + dx * dy * __ldg(plane + y2 * a.width + x2);
        ++cnt;
      }
    out[index] = cnt == 0 ? 0.f : sum / cnt;
    top_count[index] = (float)cnt;
  }
}

__global__ void __launch_bounds__(256)
psroi_bwd_kernel(int count, PsArgs a, const float* __restrict__ top_diff, const float* __restrict__ top_count,
                 const float* __restrict__ data, const float* __restrict__ rois, const float* __restrict__ trans,
                 float* __restrict__ data_diff, float* __restrict__ trans_diff) {
  for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < count; index += gridDim.x * blockDim.x) {
    const float tc = top_count[index];
    if (tc <= 0.f) continue;
    const PsBin b = ps_bin(index, a, rois, trans);
    const float diff_val = top_diff[index] / tc;
    const size_t base = ((size_t)b.batch * a.channels + b.c) * a.height * a.width;
    const float* __restrict__ plane = data + base;
    float* __restrict__ dplane = data_diff + base;
    float acc_x = 0.f, acc_y = 0.f;
    for (int ih = 0; ih < a.sample_per_part; ++ih)
      for (int iw = 0; iw < a.sample_per_part; ++iw) {
        float w = b.wstart + iw * b.sub_w, h = b.hstart + ih * b.sub_h;
        if (w < -0.5f || w > a.width - 0.5f || h < -0.5f || h > a.height - 0.5f) continue;
        w = fminf(fmaxf(w, 0.f), a.width - 1.f);
        h = fminf(fmaxf(h, 0.f), a.height - 1.f);
        const int x0 = (int)floorf(w), x1 = (int)ceilf(w), y0 = (int)floorf(h), y1 = (int)ceilf(h);
        const float dx = w - x0, dy = h - y0;
        atomicAdd(dplane + y0 * a.width + x0, (1 - dx) * (1 - dy) * diff_val);
        atomicAdd(dplane + y1 * a.width + x0, (1 - dx) * dy * diff_val);
        atomicAdd(dplane + y0 * a.width + x1, dx * (1 - dy) * diff_val);
        atomicAdd(dplane + y1 * a.width + x1, dx * dy * diff_val);
        if (a.no_trans) continue;
        const float u00 = __ldg(plane + y0 * a.width + x0), u01 = __ldg(plane + y1 * a.width + x0);
        const float u10 = __ldg(plane + y0 * a.width + x1), u11 = __ldg(plane + y1 * a.width + x1);
        acc_x += (u11 * dy + u10 * (1 - dy) - u01 * dy - u00 * (1 - dy)) * a.trans_std * diff_val * b.roi_w;
        acc_y += (u11 * dx + u01 * (1 - dx) - u10 * dx - u00 * (1 - dx)) * a.trans_std * diff_val * b.roi_h;
      }
    if (!a.no_trans) {
      // one atomic per bin and direction (the reference issues one per sample, :259-260)
      const size_t t0 = (((size_t)b.n * a.num_classes + b.class_id) * 2) * a.part_size;
      atomicAdd(trans_diff + (t0 + b.part_h) * a.part_size + b.part_w, acc_x);
      atomicAdd(trans_diff + (t0 + a.part_size + b.part_h) * a.part_size + b.part_w, acc_y);
    }
  }
}

static int ps_args(PsArgs& a, int channels, int height, int width, int no_trans, int channels_trans, float spatial_scale,
                   int output_dim, int group_size, int pooled_size, int part_size, int sample_per_part, float trans_std) {
  if (channels <= 0 || height <= 0 || width <= 0 || output_dim <= 0 || group_size <= 0 || pooled_size <= 0 ||
      part_size <= 0 || sample_per_part <= 0)
    return MRB_ERR_BAD_ARG;
  a.channels = channels; a.height = height; a.width = width; a.pooled = pooled_size; a.output_dim = output_dim;
  a.group_size = group_size; a.part_size = part_size; a.sample_per_part = sample_per_part; a.no_trans = no_trans;
  // deform_pool_kernel_cuda.cu:293-294
  a.num_classes = no_trans ? 1 : channels_trans / 2;
  if (a.num_classes <= 0) return MRB_ERR_BAD_ARG;
  a.channels_each_class = no_trans ? output_dim : output_dim / a.num_classes;
  if (a.channels_each_class <= 0) return MRB_ERR_BAD_ARG;
  a.spatial_scale = spatial_scale; a.trans_std = trans_std;
  return MRB_OK;
}
}  // namespace mrb
using namespace mrb;

MRB_API int mrb_deform_psroi_fwd(const float* data, const float* rois, const float* trans, float* out, float* top_count,
                                 int batch, int channels, int height, int width, int num_rois, int no_trans,
                                 int channels_trans, float spatial_scale, int output_dim, int group_size,
                                 int pooled_size, int part_size, int sample_per_part, float trans_std,
                                 mrb_stream_t stream) {
  if (num_rois == 0) return MRB_OK;
  PsArgs a;
  int rc = ps_args(a, channels, height, width, no_trans, channels_trans, spatial_scale, output_dim, group_size,
                   pooled_size, part_size, sample_per_part, trans_std);
  if (rc) return rc;
  if (!data || !rois || !out || !top_count || (!no_trans && !trans)) return MRB_ERR_BAD_ARG;
  const int64_t count = (int64_t)num_rois * output_dim * pooled_size * pooled_size;
  if (count >= (1ll << 31)) return MRB_ERR_UNSUPPORTED;
  psroi_fwd_kernel<<<grid_for(count, 256, 8, 8), 256, 0, (cudaStream_t)stream>>>((int)count, a, data, rois, trans, out, top_count);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_deform_psroi_bwd(const float* out_grad, const float* data, const float* rois, const float* trans,
                                 const float* top_count, float* in_grad, float* trans_grad, int batch, int channels,
                                 int height, int width, int num_rois, int no_trans, int channels_trans,
                                 float spatial_scale, int output_dim, int group_size, int pooled_size, int part_size,
                                 int sample_per_part, float trans_std, mrb_stream_t stream) {
  if (num_rois == 0) return MRB_OK;
  PsArgs a;
  int rc = ps_args(a, channels, height, width, no_trans, channels_trans, spatial_scale, output_dim, group_size,
                   pooled_size, part_size, sample_per_part, trans_std);
  if (rc) return rc;
  if (!out_grad || !data || !rois || !top_count || !in_grad || (!no_trans && (!trans || !trans_grad))) return MRB_ERR_BAD_ARG;
  const int64_t count = (int64_t)num_rois * output_dim * pooled_size * pooled_size;
  if (count >= (1ll << 31)) return MRB_ERR_UNSUPPORTED;
  psroi_bwd_kernel<<<grid_for(count, 256, 8, 8), 256, 0, (cudaStream_t)stream>>>((int)count, a, out_grad, top_count, data, rois,
                                                                               trans, in_grad, trans_grad);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
