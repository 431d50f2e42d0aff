// mask_targets.cu -- mask-head training targets on the GPU: project_masks_on_boxes (reference
// modeling/roi_heads/mask_head/loss.py:11-42) = for every positive proposal, crop its matched instance's segmentation to
// the proposal box, resize to M x M and rasterise.  The reference does this on the HOST, one proposal at a time
// (SegmentationMask.crop -> resize -> pycocotools frPyObjects/merge/decode, flagged as a bottleneck at loss.py:31-32),
// and uploads the result.  Here one CTA rasterises one proposal straight from the instances' polygon vertices.
//
//   polys      : packed vertex list, float2 (x, y) in image coordinates
//   poly_start : [num_polys + 1] vertex range of every polygon
//   inst_start : [num_instances + 1] polygon range of every instance (an instance = union of its polygons, as
//                mask_utils.merge)
//   rois       : [R, 4] proposal boxes (x1, y1, x2, y2), inst_of_roi : [R] matched instance index
//   out        : [R, M, M] fp32 in {0, 1}
// Rule: cell (i, j) is inside iff its CENTRE -- image point (x1 + (j + .5) (x2 - x1) / M, y1 + (i + .5) (y2 - y1) / M), the
// crop + resize of segmentation_mask.py:239-275 applied to the sampling point instead of the vertices -- lies inside the
// union of the polygons under the even-odd rule (crossing test with half-open edges).  pycocotools' rleFrPoly samples a
// 5x upsampled boundary instead and differs on cells the boundary passes through; it is not installed in this image, so
// that boundary rule is NOT pinned (tests compare with a numpy statement of THIS rule and with exact rectangles).
#include "common.cuh"

namespace mrb {

__global__ void __launch_bounds__(256)
mask_targets_poly_kernel(const float2* __restrict__ polys, const int* __restrict__ poly_start, const int* __restrict__ inst_start,
                         const float* __restrict__ rois, const int* __restrict__ inst_of_roi, float* __restrict__ out, int M) {
  const int r = blockIdx.x;
  const float x1 = rois[r * 4 + 0], y1 = rois[r * 4 + 1], x2 = rois[r * 4 + 2], y2 = rois[r * 4 + 3];
  const float sx = (x2 - x1) / (float)M, sy = (y2 - y1) / (float)M;
  const int inst = inst_of_roi[r];
  const int p0 = inst_start[inst], p1 = inst_start[inst + 1];
  for (int cell = threadIdx.x; cell < M * M; cell += blockDim.x) {
    const int i = cell / M, j = cell - i * M;
    const float px = x1 + ((float)j + 0.5f) * sx, py = y1 + ((float)i + 0.5f) * sy;
    bool inside_any = false;
    for (int p = p0; p < p1; ++p) {
      const int v0 = poly_start[p], v1 = poly_start[p + 1];
      if (v1 - v0 < 3) continue;
      bool in = false;
      float2 a = __ldg(polys + v1 - 1);
      for (int v = v0; v < v1; ++v) {
        const float2 b = __ldg(polys + v);
        if ((a.y > py) != (b.y > py)) {
          const float xi = (b.x - a.x) * (py - a.y) / (b.y - a.y) + a.x;
          if (px < xi) in = !in;
        }
        a = b;
      }
      inside_any |= in;
    }
    out[(size_t)r * M * M + cell] = inside_any ? 1.f : 0.f;
  }
}

// Rectangle instances (the synthetic benchmark's masks; also what a box-only dataset yields): cell (i, j) of ROI r is inside iff
// its centre, in the M x M frame of the ROI, lies within the matched ground-truth box scaled into that frame -- the closed-form
// special case of the rule above, in the operation order of the harness's PyTorch formulation (mrb_b200/model/roi_heads.py,
// MaskHead.mask_targets), so the two agree bit for bit.
__global__ void __launch_bounds__(256)
mask_targets_rect_kernel(const float4* __restrict__ gt, const float* __restrict__ rois, int roi_stride, float* __restrict__ out, int R,
                         int M) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * M * M) return;
  const int r = (int)(t / (M * M)), cell = (int)(t - (long long)r * M * M), i = cell / M, j = cell - i * M;
  const float* q = rois + (size_t)r * roi_stride;
  const float x1 = q[0], y1 = q[1], x2 = q[2], y2 = q[3];
  const float4 g = gt[r];
  const float sx = __fdiv_rn((float)M, fmaxf(__fsub_rn(x2, x1), 1e-6f)), sy = __fdiv_rn((float)M, fmaxf(__fsub_rn(y2, y1), 1e-6f));
  const float cx = (float)j + 0.5f, cy = (float)i + 0.5f;
  const bool inx = cx >= __fmul_rn(__fsub_rn(g.x, x1), sx) && cx <= __fmul_rn(__fsub_rn(g.z, x1), sx);
  const bool iny = cy >= __fmul_rn(__fsub_rn(g.y, y1), sy) && cy <= __fmul_rn(__fsub_rn(g.w, y1), sy);
  out[t] = (inx && iny) ? 1.f : 0.f;
}

}  // namespace mrb
using namespace mrb;

MRB_API int mrb_mask_targets_rect(const float* gt_boxes, const float* rois, int roi_stride, float* out, int num_rois, int mask_size,
                                  mrb_stream_t stream) {
  if (num_rois < 0 || mask_size <= 0 || roi_stride < 4) return MRB_ERR_BAD_ARG;
  if (num_rois == 0) return MRB_OK;
  if (!gt_boxes || !rois || !out || ((uintptr_t)gt_boxes & 15)) return MRB_ERR_BAD_ARG;
  const long long total = (long long)num_rois * mask_size * mask_size;
  mask_targets_rect_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)gt_boxes, rois, roi_stride, out, num_rois,
                                                                                   mask_size);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_mask_targets_polygons(const float* polys_xy, const int* poly_start, const int* inst_start, const float* rois,
                                      const int* inst_of_roi, float* out, int num_rois, int mask_size, mrb_stream_t stream) {
  if (num_rois < 0 || mask_size <= 0) return MRB_ERR_BAD_ARG;
  if (num_rois == 0) return MRB_OK;
  if (!polys_xy || !poly_start || !inst_start || !rois || !inst_of_roi || !out) return MRB_ERR_BAD_ARG;
  mask_targets_poly_kernel<<<num_rois, 256, 0, (cudaStream_t)stream>>>((const float2*)polys_xy, poly_start, inst_start, rois, inst_of_roi,
                                                                      out, mask_size);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
