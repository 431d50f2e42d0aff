/*
 * mrb_b200.h -- C ABI of libmrb_b200.so: the Blackwell (sm_100a) replacement for the
 * native surface of facebookresearch/maskrcnn-benchmark (`maskrcnn_benchmark._C`,
 * reference csrc/vision.cpp:9-25) plus the dense-conv engine behind
 * `maskrcnn_benchmark.layers.Conv2d` / `FrozenBatchNorm2d`.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`.
 *   - returns 0 (MRB_OK), a positive cudaError_t, or a negative MRB_ERR_* code.
 *     Never throws, never prints (the reference DCN kernels printf-and-continue,
 *     csrc/cuda/deform_conv_kernel_cuda.cu:279-283; here the binding raises).
 *   - explicit stream parameter (the reference launches NMS / DCN on the legacy
 *     default stream: csrc/cuda/nms.cu:94, deform_conv_kernel_cuda.cu:272).
 *   - no allocation and no host synchronisation inside; scratch is supplied by the
 *     caller and sized by the matching mrb_*_workspace_bytes() query.
 *   - there is NO CPU path.  On a machine without a Blackwell GPU every compute
 *     entry point fails with a CUDA error; nothing falls back.
 *
 * The reference-side binding (ctypes `_C` module) is
 * maskrcnn-benchmark_b200/maskrcnn_benchmark/_C.py; see INTEGRATION.md.
 */
#ifndef MRB_B200_H_
#define MRB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mrb_stream_t; /* cudaStream_t */

#define MRB_OK 0
#define MRB_ERR_BAD_ARG (-1)
#define MRB_ERR_UNSUPPORTED (-2)
#define MRB_ERR_WORKSPACE (-3)
#define MRB_ERR_DRIVER (-4)

/* memory layouts of a logical [N,C,H,W] activation */
#define MRB_LAYOUT_NCHW 0
#define MRB_LAYOUT_NHWC 1 /* torch.channels_last */

/* element types */
#define MRB_F32 0
#define MRB_BF16 1

int mrb_version(void);
/* static string for MRB_ERR_* and cudaError_t values */
const char* mrb_error_string(int code);

/* ------------------------------------------------------------------ ROIAlign
 * replaces ROIAlign_forward / ROIAlign_backward (csrc/ROIAlign.h:11-45;
 * CUDA csrc/cuda/ROIAlign_cuda.cu:257-346; CPU csrc/cpu/ROIAlign_cpu.cpp:221-257).
 * input  : logical [batch, channels, height, width] in `layout`, fp32
 * rois   : [num_rois, 5] fp32 (batch_idx, x1, y1, x2, y2) image coordinates
 * output : [num_rois, channels, pooled_h, pooled_w] fp32, contiguous (NCHW order)
 * Semantics: aligned=False, ROI coords not rounded, min ROI size 1, sampling_ratio<=0
 * => ceil(roi/pooled) adaptive grid.  Forward is bit-identical to the reference CPU
 * kernel (no FMA contraction, same operation order). */
int mrb_roi_align_fwd(const float* input, const float* rois, float* output, int num_rois, int batch,
                      int channels, int height, int width, int pooled_h, int pooled_w,
                      float spatial_scale, int sampling_ratio, int layout, mrb_stream_t stream);
/* grad_input : logical [batch, channels, height, width] in `layout`; fully overwritten
 * (zero-filled inside, as at::zeros at ROIAlign_cuda.cu:316). */
int mrb_roi_align_bwd(const float* grad_output, const float* rois, float* grad_input, int num_rois,
                      int batch, int channels, int height, int width, int pooled_h, int pooled_w,
                      float spatial_scale, int sampling_ratio, int layout, mrb_stream_t stream);

/* Multi-level (FPN) ROIAlign: the whole `Pooler.forward` loop (modeling/poolers.py:91-121:
 * LevelMapper + per-level nonzero/ROIAlign/index-scatter) in one launch.  Feature maps are NHWC,
 * bf16 or fp32 (`dtype`, also the output element type); each ROI picks its level in the kernel:
 * floor(canonical_level + log2(sqrt(area)/canonical_scale + 1e-6)) clamped to [k_min,k_max]
 * (poolers.py:31-42), levels[l] <-> k_min + l.
 *   feats_host / heights_host / widths_host / scales_host : HOST arrays, one entry per level
 *                  (the device pointers inside feats_host are device pointers)
 *   output       : [num_rois, channels, P, P] (out_nhwc == 0) or [num_rois, P, P, channels]
 * Backward accumulates (red.add.v4.f32) into caller-zeroed fp32 NHWC gradient maps. */
int mrb_roi_align_fpn_fwd(const void* const* feats_host, const int* heights_host, const int* widths_host,
                          const float* scales_host, int num_levels, const float* rois, void* output,
                          int num_rois, int batch, int channels, int pooled, int sampling_ratio,
                          int k_min, int k_max, float canonical_scale, int canonical_level, int dtype,
                          int out_nhwc, mrb_stream_t stream);
int mrb_roi_align_fpn_bwd(const void* grad_output, float* const* grad_feats_host, const int* heights_host,
                          const int* widths_host, const float* scales_host, int num_levels,
                          const float* rois, int num_rois, int batch, int channels, int pooled,
                          int sampling_ratio, int k_min, int k_max, float canonical_scale,
                          int canonical_level, int dtype, int out_nhwc, mrb_stream_t stream);

/* ------------------------------------------------------------------- ROIPool
 * replaces ROIPool_forward / ROIPool_backward (csrc/ROIPool.h:11-45;
 * csrc/cuda/ROIPool_cuda.cu:16-202).  NCHW fp32.  argmax: int32 offset in the H*W
 * plane, -1 for an empty bin. */
int mrb_roi_pool_fwd(const float* input, const float* rois, float* output, int32_t* argmax,
                     int num_rois, int batch, int channels, int height, int width, int pooled_h,
                     int pooled_w, float spatial_scale, mrb_stream_t stream);
int mrb_roi_pool_bwd(const float* grad_output, const float* rois, const int32_t* argmax,
                     float* grad_input, int num_rois, int batch, int channels, int height, int width,
                     int pooled_h, int pooled_w, mrb_stream_t stream);

/* ----------------------------------------------------------------------- NMS
 * replaces nms (csrc/nms.h:10-28; CPU csrc/cpu/nms_cpu.cpp:5-75; CUDA csrc/cuda/nms.cu:70-131).
 * Parity target is the CPU path: legacy "+1" areas, suppress when IoU >= threshold
 * (nms_cpu.cpp:60), kept indices returned ASCENDING BY INDEX (nms_cpu.cpp:64).
 * IoU is evaluated with round-to-nearest mul/add/sub/div and no FMA, so indices are
 * bit-exact with the reference for distinct scores; equal scores are ordered by
 * ascending index (the reference's non-stable sort leaves tie order unspecified).
 * Fully on device: no D2H mask copy, no host scan (cf. nms.cu:100-123).
 *   boxes [n,4] fp32 xyxy, scores [n] fp32
 *   keep  [n]   int64, first *num_keep entries valid
 *   num_keep    int32 (device)
 *   workspace   >= mrb_nms_workspace_bytes(n) bytes, 16-byte aligned */
size_t mrb_nms_workspace_bytes(int n);
int mrb_nms(const float* boxes, const float* scores, int n, float threshold, int64_t* keep,
            int32_t* num_keep, void* workspace, size_t workspace_bytes, mrb_stream_t stream);
/* Batched NMS over `num_problems` independent box sets stored back to back:
 * problem p owns rows [offsets_host[p], offsets_host[p+1]); keep rows are written at the
 * same offsets (indices relative to the problem), num_keep[p] per problem.  One
 * launch sequence for all (image, level) pairs of an RPN step
 * (modeling/rpn/inference.py:116-121 calls nms once per pair). */
size_t mrb_nms_batched_workspace_bytes(const int* offsets_host, int num_problems);
int mrb_nms_batched(const float* boxes, const float* scores, const int* offsets_host, int num_problems,
                    float threshold, int64_t* keep, int32_t* num_keep, void* workspace,
                    size_t workspace_bytes, mrb_stream_t stream);

/* The same for problems whose rows are ALREADY in descending-score order (the RPN's sorted top-k, inference.py:91-95): the
 * rank sort is skipped.  Same kept set as mrb_nms_batched on such input. */
int mrb_nms_batched_presorted(const float* boxes, const int* offsets_host, int num_problems, float threshold, int64_t* keep,
                              int32_t* num_keep, void* workspace, size_t workspace_bytes, mrb_stream_t stream);

/* ----------------------------------------------------------- SigmoidFocalLoss
 * replaces SigmoidFocalLoss_forward / _backward (csrc/SigmoidFocalLoss.h:10-41;
 * csrc/cuda/SigmoidFocalLoss_cuda.cu:20-188).  logits [A,num_classes] fp32, targets [A]
 * int32 (1..num_classes positive, 0 background, -1 ignore). */
int mrb_sigmoid_focal_fwd(const float* logits, const int32_t* targets, float* losses, int64_t num_anchors,
                          int num_classes, float gamma, float alpha, mrb_stream_t stream);
int mrb_sigmoid_focal_bwd(const float* logits, const int32_t* targets, const float* d_losses,
                          float* d_logits, int64_t num_anchors, int num_classes, float gamma,
                          float alpha, mrb_stream_t stream);

/* --------------------------------------------------- deformable convolution
 * replaces deform_conv_forward / _backward_input / _backward_parameters and
 * modulated_deform_conv_forward / _backward (csrc/deform_conv.h:11-191;
 * csrc/cuda/deform_conv_cuda.cu:158-691, deform_conv_kernel_cuda.cu:197-874).
 * All tensors NCHW fp32 contiguous.  `mask` == NULL selects DCNv1 (no modulation);
 * `bias` may be NULL.  No `columns` tensor is materialised in HBM beyond the caller's
 * workspace.  offset [N, dg*2*kh*kw, Ho, Wo]; mask [N, dg*kh*kw, Ho, Wo];
 * weight [Cout, Cin/groups, kh, kw]. */
typedef struct mrb_dcn_params {
  int batch, cin, height, width, cout, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int groups, deformable_groups;
} mrb_dcn_params;
size_t mrb_deform_conv_workspace_bytes(const mrb_dcn_params* p);
int mrb_deform_conv_fwd(const mrb_dcn_params* p, const float* input, const float* offset,
                        const float* mask, const float* weight, const float* bias, float* output,
                        void* workspace, size_t workspace_bytes, mrb_stream_t stream);
/* grad_input / grad_offset / grad_mask are overwritten; grad_weight / grad_bias are
 * ACCUMULATED into (+= scale * dW), matching deform_conv_backward_parameters' `scale`
 * and the modulated backward's accumulate-into-zeroed-buffers behaviour
 * (layers/dcn/deform_conv_func.py:97-104,218-224).  Any grad pointer may be NULL to skip. */
int mrb_deform_conv_bwd(const mrb_dcn_params* p, const float* input, const float* offset,
                        const float* mask, const float* weight, const float* grad_output,
                        float* grad_input, float* grad_offset, float* grad_mask, float* grad_weight,
                        float* grad_bias, float scale, void* workspace, size_t workspace_bytes,
                        mrb_stream_t stream);

/* ---------------------------------------------- deformable PS-ROI pooling
 * replaces deform_psroi_pooling_forward / _backward (csrc/deform_pool.h:11-70;
 * csrc/cuda/deform_pool_kernel_cuda.cu:53-365). NCHW fp32.  channels_trans = trans.size(1)
 * (2 * num_classes; ignored when no_trans).  The backward ACCUMULATES into in_grad /
 * trans_grad (the reference uses atomicAdd into caller-zeroed tensors). */
int mrb_deform_psroi_fwd(const float* data, const float* rois, const float* trans, float* out,
                         float* top_count, int batch, int channels, int height, int width, int num_rois,
                         int no_trans, int channels_trans, float spatial_scale, int output_dim,
                         int group_size, int pooled_size, int part_size, int sample_per_part,
                         float trans_std, mrb_stream_t stream);
int mrb_deform_psroi_bwd(const float* out_grad, const float* data, const float* rois, const float* trans,
                         const float* top_count, float* in_grad, float* trans_grad, int batch,
                         int channels, int height, int width, int num_rois, int no_trans,
                         int channels_trans, float spatial_scale, int output_dim, int group_size,
                         int pooled_size, int part_size, int sample_per_part, float trans_std,
                         mrb_stream_t stream);

/* --------------------------------------------------------- dense conv engine
 * replaces the ATen/cuDNN convolution behind layers.Conv2d (layers/misc.py:30-43) and the
 * FrozenBatchNorm2d / ReLU / residual-add elementwise passes that follow it in
 * Bottleneck.forward (modeling/backbone/resnet.py:324-344; layers/batch_norm.py:27-31).
 * Implicit GEMM on tcgen05 tensor cores: NHWC bf16 activations, KRSC bf16 weights
 * ([Cout, kh, kw, Cin]), fp32 accumulation in TMEM, fused epilogue
 *     y = act( acc * scale[c] + bias[c] + residual )        (scale/bias/residual optional)
 * TMA feeds shared memory directly from the NHWC tensor (im2col is folded into the TMA
 * box coordinates; no column matrix exists anywhere). */
#define MRB_CONV_PAD_W 1              /* mrb_conv_params.flags: `pad_w` holds the width padding (else pad_w == pad) */
/* mrb_conv_params.flags: grouped convolution with cin == cout, (cin / groups) dividing 64 (ResNeXt 32x4d/32x8d, reference
 * modeling/backbone/resnet.py:302-311 `groups=num_groups`), expressed as block-diagonal 64-channel super-groups: output
 * channels [64b, 64b+64) read input channels [64b, 64b+64) only.  The weight operand is the EXPANDED filter
 * [Cout][kh][kw][64] (zero outside each group's own channels; mrb_b200/ops.py expand_grouped_weight); dgrad takes the
 * prepared (flipped, transposed within the block) [Cin][kh][kw][64] through mrb_conv2d_dgrad_prepared; wgrad returns
 * [Cout][kh][kw][128] -- for every output channel the gradient against the 128 input channels of its Cout tile -- from
 * which the caller keeps the group's own columns.  Stride 1 only. */
#define MRB_CONV_GROUPED64 2
typedef struct mrb_conv_params {
  int batch, height, width, cin;   /* input  NHWC */
  int cout, kh, kw;                /* filter KRSC (rectangular kernels allowed) */
  int stride, pad;                 /* pad: height (and, by default, width) padding; dilation 1; groups 1; stride 2 only
                                      for 1x1 */
  int relu;                        /* fused ReLU in the epilogue */
  int out_dtype;                   /* MRB_BF16 or MRB_F32 */
  int out_h, out_w;                /* 0 = (in + 2*pad - k)/stride + 1; a smaller explicit size computes only the
                                      top-left out_h x out_w outputs (used by the space-to-depth stem, whose 4x4
                                      kernel needs padding 2 on the left/top but 1 on the right/bottom) */
  int pad_w, flags;                /* see MRB_CONV_PAD_W */
  /* Optional element strides {image, row, pixel} (channel stride is always 1) of the input-side tensor (`input` of
   * fwd/wgrad, `grad_input`/`add`/`relu_mask` of dgrad) and of the output-side tensor (`output`/`residual` of fwd,
   * `grad_output` of dgrad/wgrad).  All zero = dense NHWC.  Multiples of 8 elements (TMA: 16 B); stride 1 only.
   * This is how a convolution reads or writes a strided window of a larger tensor in place: the 2x2/stride-2
   * transposed convolution of the mask head writes its sub-pixel planes straight into the [N,2H,2W,C] result, and
   * the 7x7/2 stem reads overlapping 4-pixel windows of its space-to-depth input as 64-channel pixels. */
  long long x_pitch[3];
  long long y_pitch[3];
} mrb_conv_params;
/* Mask-head training targets from polygon segmentations: project_masks_on_boxes (modeling/roi_heads/mask_head/loss.py:11-42,
 * the host-side per-proposal crop/resize/rasterise loop) as one launch.  polys_xy: packed (x, y) vertices, image
 * coordinates; poly_start[num_polys + 1], inst_start[num_instances + 1]: CSR ranges (instance = union of its polygons);
 * rois [R,4] xyxy; inst_of_roi [R]; out [R, M, M] fp32 in {0,1}.  Cell-centre, even-odd rule (see csrc/mask_targets.cu). */
int mrb_mask_targets_polygons(const float* polys_xy, const int* poly_start, const int* inst_start, const float* rois,
                              const int* inst_of_roi, float* out, int num_rois, int mask_size, mrb_stream_t stream);

/* ------------------------------------------------------------- detection glue (csrc/detect_glue.cu)
 * The stages between the tensor-core phases of a train step that the reference runs as Python over BoxList objects
 * (hundreds of small launches, several host synchronisations): one launch each, fixed shapes, no host sync.
 *
 * mrb_rpn_decode: RPNPostProcessor.forward_for_single_feature_map after its top-k (modeling/rpn/inference.py:91-111) for one
 * level: gathers the k selected anchors' deltas, BoxCoder.decode (modeling/box_coder.py:52-95), clip_to_image
 * (structures/bounding_box.py:198-212) and the sigmoid of the k logits.  logits [N, A], deltas [N, A, 4], anchors [A, 4],
 * topk_idx [N, k] int64, image_w / image_h [N] fp32 -> boxes [N, k, 4], scores [N, k].  weights_host: 4 floats. */
int mrb_rpn_decode(const float* logits, const float* deltas, const float* anchors, const int64_t* topk_idx, const float* image_w,
                   const float* image_h, float* boxes, float* scores, int num_images, int num_anchors, int k,
                   const float* weights_host, float xform_clip, mrb_stream_t stream);
/* mrb_rpn_topk_decode: objectness.topk(pre_nms_top_n, sorted=True) (inference.py:91-95) AND mrb_rpn_decode_packed in one
 * launch: one thread-block cluster per image selects the k largest logits of the level (radix select across the cluster's
 * CTAs), sorts them (descending logit, ascending anchor among equals) and decodes.  boxes [N, k, 4], scores [N, k]. */
int mrb_rpn_topk_decode(const float* head_output, int anchors_per_location, int pixel_stride, const float* anchors,
                        const float* image_w, const float* image_h, float* boxes, float* scores, int num_images, int num_anchors,
                        int k, const float* weights_host, float xform_clip, mrb_stream_t stream);
/* mrb_rpn_collect: what follows the NMS (inference.py:116-123, select_over_all_levels :154-181, add_gt_proposals :53-74).
 * boxes / scores / keep: the (level-major, image-minor) problems of mrb_nms_batched back to back, level l holding
 * num_images x k_per_level[l] rows; num_keep [num_levels * num_images].  Per image the candidates are the first
 * min(num_keep, post_nms_top_n) survivors of every level; of those the fpn_post_nms_top_n best are kept -- over the whole
 * batch (per_batch, training: original order preserved) or per image (sorted by score).  Output rows [N, W + gmax] with
 * W = mrb_rpn_collect_width(...): the kept boxes, then zero rows (valid = 0, score = -1), then the gmax ground-truth slots
 * (score 1, valid for the first gt_count[i]); gmax = 0 appends nothing.  Ties at the cut are taken in (image, slot) order. */
int mrb_rpn_collect(const float* boxes, const float* scores, const int64_t* keep, const int32_t* num_keep,
                    const int* k_per_level_host, int num_levels, int num_images, int post_nms_top_n, int fpn_post_nms_top_n,
                    int per_batch, int sorted, const float* gt_boxes, const int32_t* gt_count, int gmax, float* out_boxes,
                    float* out_scores, uint8_t* out_valid, mrb_stream_t stream);
int mrb_rpn_collect_width(const int* k_per_level_host, int num_levels, int num_images, int post_nms_top_n,
                          int fpn_post_nms_top_n, int per_batch);
/* mrb_roi_assign_sample: FastRCNNLossComputation.subsample (modeling/roi_heads/box_head/loss.py:41-118) for the batch: IoU
 * with the ground truth (structures/boxlist_ops.py:53-89), Matcher without the low-quality pass (modeling/matcher.py:42-81),
 * labels, BalancedPositiveNegativeSampler (balanced_positive_negative_sampler.py:19-68) driven by rand_keys [N, P] (iid
 * uniform: the int(batch * fraction) smallest keys among the positives, the rest from the negatives == randperm[:n]),
 * BoxCoder.encode (box_coder.py:22-50).  boxes [N, P, 4], valid [N, P] u8, gt_boxes [N, gmax, 4], gt_labels [N, gmax] int64,
 * gt_count [N].  Outputs, S = batch_size_per_image rows per image, sampled rows first in proposal order, then unsampled
 * rows as padding with label -1: out_rois [N*S, 5] (image index, box), out_labels [N*S] int64, out_reg_targets [N*S, 4],
 * out_gt_index [N*S] int64.  With mask_rois_per_image = M > 0 also the mask branch's list (keep_only_positive_boxes,
 * mask_head/mask_head.py:11-32, at fixed width): the positives among the S rows first: mask_rois [N*M, 5], mask_labels
 * [N*M] int64 (0 on padding), mask_weight [N*M] (1 / 0), mask_gt_index [N*M] int64. */
int mrb_roi_assign_sample(const float* boxes, const uint8_t* valid, const float* rand_keys, const float* gt_boxes,
                          const int64_t* gt_labels, const int32_t* gt_count, int num_images, int num_proposals, int gmax,
                          int batch_size_per_image, float positive_fraction, float fg_iou, float bg_iou,
                          const float* weights_host, int mask_rois_per_image, float* out_rois, int64_t* out_labels,
                          float* out_reg_targets, int64_t* out_gt_index, float* mask_rois, int64_t* mask_labels,
                          float* mask_weight, int64_t* mask_gt_index, int64_t* out_proposal_index /* [N*S] or NULL */,
                          mrb_stream_t stream);
/* Box-head post-processing (PostProcessor.forward, modeling/roi_heads/box_head/inference.py:45-149), fixed shapes, no host
 * synchronisation.  mrb_box_post_decode: softmax over the C class logits of `outputs` [N*P, ld] (C logits, then 4C regression
 * outputs per row), per-class BoxCoder.decode of `proposals` [N*P, 4] + clip_to_image, score threshold -> one NMS problem per
 * (image, class > 0): boxes / scores [(N * (C-1)) * P] in the layout of mrb_nms_batched (score -1 = no candidate).
 * mrb_box_post_select, after that NMS: the detections_per_img best survivors of every image over all classes, in (class,
 * proposal) order -> out_boxes [N, D, 4], out_scores [N, D], out_labels [N, D] int64, out_count [N] (rows beyond are zero). */
int mrb_box_post_decode(const float* outputs, int ld, int num_classes, const float* proposals, const uint8_t* valid,
                        const float* image_w, const float* image_h, int num_images, int proposals_per_image, float score_thresh,
                        const float* weights_host, float xform_clip, float* boxes, float* scores, mrb_stream_t stream);
int mrb_box_post_select(const float* boxes, const float* scores, const int64_t* keep, const int32_t* num_keep, int num_images,
                        int proposals_per_image, int num_classes, int detections_per_img, float* out_boxes, float* out_scores,
                        int64_t* out_labels, int32_t* out_count, mrb_stream_t stream);
/* mrb_rpn_anchor_match: RPNLossComputation.match_targets_to_anchors + the labelling of prepare_targets
 * (modeling/rpn/loss.py:40-90): IoU of every anchor with the ground truth, Matcher with allow_low_quality_matches
 * (matcher.py:42-112), label 1 / 0 / -1 (between thresholds, or outside the image by more than straddle_thresh;
 * straddle_thresh < 0 disables the visibility test, anchor_generator.py:100-116).  anchors [A, 4] (shared by the images),
 * labels [N, A] fp32, matched_gt [N, A] int32 (arg-max ground truth of every anchor). */
size_t mrb_rpn_anchor_match_workspace_bytes(int num_images, int num_anchors, int gmax);
/* mrb_rpn_sample: BalancedPositiveNegativeSampler over the labelled anchors of every image (balanced_positive_negative_sampler.py:
 * 19-68; "n random elements" = the n smallest of the iid rand_keys [N, A]) + BoxCoder.encode of the sampled positives against
 * their matched ground truth (rpn/loss.py:86-90), one launch (a cluster of CTAs per image).  P = int(batch * fraction).
 * pos_idx / pos_ok / reg_targets [N, P]; sel_idx / sel_label / sel_weight [N, P + batch]: the P positive slots, then the
 * negative slots (weight 0 = unused slot) -- the inputs of mrb_rpn_loss_fwd. */
int mrb_rpn_sample(const float* labels, const int32_t* matched_gt, const float* rand_keys, const float* anchors,
                   const float* gt_boxes, int num_images, int num_anchors, int gmax, int batch_size_per_image,
                   float positive_fraction, const float* weights_host, int64_t* pos_idx, uint8_t* pos_ok, float* reg_targets,
                   int64_t* sel_idx, float* sel_label, float* sel_weight, mrb_stream_t stream);
int mrb_rpn_anchor_match(const float* anchors, const float* gt_boxes, const int32_t* gt_count, const float* image_w,
                         const float* image_h, int num_images, int num_anchors, int gmax, float fg_iou, float bg_iou,
                         float straddle_thresh, float* labels, int32_t* matched_gt, void* workspace, size_t workspace_bytes,
                         mrb_stream_t stream);

/* ------------------------------------------------------------- loss stages (csrc/loss_glue.cu)
 * Forward + backward of the three loss computations of the train step, each a launch or two, deterministic reductions.
 * `result` buffers are small fp32 device arrays written by the forward and read by the backward; grad_* are the upstream
 * gradients of the scalar losses as DEVICE scalars.
 *
 * RPN (modeling/rpn/loss.py:92-131): head_outputs_host[l] = level l's head output [N, locations[l], pixel_stride] fp32 (A
 * logits, then 4A deltas, then padding up to pixel_stride per location: the NHWC output of the fused cls+bbox 1x1 conv).  sel_* [N, num_sel]: the sampled anchors
 * (index into the concatenated anchor list, label 1/0, weight 1 / 0 = padding); pos_* [N, num_pos]: the sampled positives
 * with their regression targets.  result[3] = {objectness loss, box loss (smooth-L1 beta), #sampled}.  The backward writes
 * ONLY the sampled entries of grad_outputs_host[l] (same shapes), which the caller has zeroed. */
int mrb_rpn_loss_fwd(void* const* head_outputs_host, const int* locations_host, int num_levels, int num_images,
                     int anchors_per_location, int pixel_stride, const int64_t* sel_idx, const float* sel_label, const float* sel_weight, int num_sel,
                     const int64_t* pos_idx, const uint8_t* pos_ok, const float* reg_targets, int num_pos, float beta,
                     float* result, mrb_stream_t stream);
int mrb_rpn_loss_bwd(void* const* head_outputs_host, void* const* grad_outputs_host, const int* locations_host, int num_levels,
                     int num_images, int anchors_per_location, int pixel_stride, const int64_t* sel_idx, const float* sel_label,
                     const float* sel_weight, int num_sel, const int64_t* pos_idx, const uint8_t* pos_ok, const float* reg_targets,
                     int num_pos, float beta, const float* result, const float* grad_objectness, const float* grad_box,
                     mrb_stream_t stream);
/* Box head (modeling/roi_heads/box_head/loss.py:120-167): outputs [R, ld] fp32 = class logits in columns [0, C), box
 * regression in [C + 4c, C + 4c + 4); labels [R] int64 (-1 = row not sampled), reg_targets [R, 4].
 * result[3] = {cross-entropy (mean over sampled rows), smooth-L1(beta 1) sum over positives / #sampled, #sampled}.
 * The backward writes the whole [R, ld] gradient. */
int mrb_box_loss_fwd(const float* outputs, int ld, int num_classes, const int64_t* labels, const float* reg_targets, int num_rois,
                     float* row_scratch /* [3 R] */, float* result, mrb_stream_t stream);
int mrb_box_loss_bwd(const float* outputs, int ld, int num_classes, const int64_t* labels, const float* reg_targets, int num_rois,
                     const float* result, const float* grad_cls, const float* grad_box, float* grad_outputs, mrb_stream_t stream);
/* Mask head (modeling/roi_heads/mask_head/loss.py:100-133): logits [R, mask_pixels, channels] bf16 (NHWC), labels [R] int64
 * (class whose plane is read, loss.py:120-126), targets [R, mask_pixels] fp32, weights [R] (1 positive / 0 padding).
 * row_losses [R] scratch; result[2] = {sum_r w_r mean_pix BCE / max(sum w, 1), max(sum w, 1)}.  The backward writes the whole
 * bf16 gradient [R, mask_pixels, channels] (channels % 8 == 0). */
int mrb_mask_loss_fwd(const void* logits_bf16, int channels, int mask_pixels, const int64_t* labels, const float* targets,
                      const float* weights, int num_rois, float* row_losses, float* result, mrb_stream_t stream);
int mrb_mask_loss_bwd(const void* logits_bf16, int channels, int mask_pixels, const int64_t* labels, const float* targets,
                      const float* weights, int num_rois, const float* result, const float* grad_loss, void* grad_logits_bf16,
                      mrb_stream_t stream);
/* mrb_rpn_decode reading logits and deltas in place from the head output [N, locations, A + 4A] (see mrb_rpn_loss_fwd). */
int mrb_rpn_decode_packed(const float* head_output, int anchors_per_location, int pixel_stride, const float* anchors,
                          const int64_t* topk_idx,
                          const float* image_w, const float* image_h, float* boxes, float* scores, int num_images,
                          int num_anchors, int k, const float* weights_host, float xform_clip, mrb_stream_t stream);

/* The same for rectangle instances (matched ground-truth box per ROI): gt_boxes [R, 4], rois [R, roi_stride >= 4] with the box
 * in its first four floats... see csrc/mask_targets.cu.  out [R, M, M] fp32 in {0,1}. */
int mrb_mask_targets_rect(const float* gt_boxes, const float* rois, int roi_stride, float* out, int num_rois, int mask_size,
                          mrb_stream_t stream);

/* Operand preparation for MRB_CONV_GROUPED64 (csrc/grouped_prep.cu).  weight: the grouped filter [C][taps][C/groups] bf16
 * (KRSC).  w_exp / wd_exp: [C][taps][64] bf16 -- the forward operand and the (flipped, per-Cout scaled) data-gradient operand
 * of mrb_conv2d_fwd / mrb_conv2d_dgrad_prepared; either may be NULL.  mrb_grouped_collapse_wgrad folds the [C][taps][128]
 * fp32 result of mrb_conv2d_wgrad(MRB_CONV_GROUPED64) back into the grouped layout [C, C/groups, kh, kw] given by its
 * element strides (co, cg, tap), overwriting or accumulating. */
int mrb_grouped_expand_weights(const void* weight_bf16, const float* scale, void* w_exp_bf16, void* wd_exp_bf16,
                               int channels, int taps, int groups, mrb_stream_t stream);
int mrb_grouped_collapse_wgrad(const float* grad_expanded128, float* grad_weight, int channels, int taps, int groups,
                               long long stride_co, long long stride_cg, long long stride_tap, int accumulate,
                               mrb_stream_t stream);

/* Deformable convolution on the tensor-core path (NHWC bf16): bilinear sampler producing the GEMM's A operand and its
 * backward.  Replaces deformable_im2col / col2im / col2im_coord and the modulated twins
 * (csrc/cuda/deform_conv_kernel_cuda.cu:197-874) for the model's DFConv2d layers (layers/misc.py:114-203); the GEMMs
 * around them are mrb_conv2d_fwd / _dgrad_prepared / _wgrad over K = kh*kw*C (see csrc/dcn_nhwc.cu).
 *   input        : [N,H,W,C] bf16 (C % 8 == 0)
 *   offset_mask  : [N,Ho,Wo,oc_pitch] fp32: channels (2t, 2t+1) = (dh, dw) of tap t; modulated != 0: channel
 *                  2*kh*kw + t = mask LOGIT of tap t (the sigmoid of DFConv2d.forward is applied inside)
 *   columns      : [N*Ho*Wo][kh*kw*C] bf16, K ordered (tap, channel) == the KRSC filter's memory order
 *   grad_input   : [N,H,W,C] fp32, ACCUMULATED into (caller zero-fills; may be NULL)
 *   grad_offset_mask : like offset_mask, fully written for the channels in use (mask-logit gradient through the sigmoid)
 * deformable_groups == 1. */
int mrb_dcn_sample_nhwc(const void* input_bf16, const float* offset_mask, void* columns_bf16, int batch, int height,
                        int width, int channels, int out_h, int out_w, int kh, int kw, int stride, int pad, int dilation,
                        int oc_pitch, int modulated, mrb_stream_t stream);
int mrb_dcn_backward_nhwc(const void* input_bf16, const float* offset_mask, const void* grad_columns_bf16,
                          float* grad_input_f32, float* grad_offset_mask, int batch, int height, int width, int channels,
                          int out_h, int out_w, int kh, int kw, int stride, int pad, int dilation, int oc_pitch,
                          int modulated, mrb_stream_t stream);
int mrb_conv2d_fwd(const mrb_conv_params* p, const void* input_bf16, const void* weight_bf16,
                   const float* scale, const float* bias, const void* residual, void* output,
                   mrb_stream_t stream);
/* dgrad: grad_input[N,H,W,Cin] = conv_transpose(grad_output[N,Ho,Wo,Cout] (bf16), weight) computed by the
 * same implicit-GEMM kernel on flipped/transposed weights that are prepared into `workspace`
 * (>= mrb_conv2d_dgrad_workspace_bytes).  `scale` (optional, per Cout) folds the frozen-BN scale of the
 * forward epilogue into the weights; `add` (optional, bf16, shaped like grad_input) is summed in (the
 * other branch of a residual join); `relu_mask` (optional, bf16, shaped like grad_input) zeroes the
 * result where the saved forward activation is <= 0 (ReLU backward of the producer layer).
 * p->out_dtype selects the grad_input element type.  stride 2 is supported for 1x1 kernels only:
 * grad_input is zero-filled, then written at the even positions; there `add` must be grad_input itself
 * (in-place accumulation of a second stride-2 branch, no zero fill). */
size_t mrb_conv2d_dgrad_workspace_bytes(const mrb_conv_params* p);
int mrb_conv2d_dgrad(const mrb_conv_params* p, const void* grad_output_bf16, const void* weight_bf16,
                     const float* scale, const void* add, const void* relu_mask, void* grad_input,
                     void* workspace, size_t workspace_bytes, mrb_stream_t stream);

/* The flipped/transposed (and BN-scaled) weights of many layers in ONE launch, for use with
 * mrb_conv2d_dgrad_prepared: prepared[l] must hold Cout*taps*Cin bf16 ([Cin][taps][Cout] layout).  All arrays are
 * HOST arrays of num_layers entries (scales_host may be NULL, or hold NULL entries). */
int mrb_conv2d_prepare_dgrad_weights(int num_layers, const void* const* weights_host, const float* const* scales_host,
                                     void* const* prepared_host, const int* couts_host, const int* taps_host,
                                     const int* cins_host, mrb_stream_t stream);
int mrb_conv2d_dgrad_prepared(const mrb_conv_params* p, const void* grad_output_bf16, const void* prepared_weight,
                              const void* add, const void* relu_mask, void* grad_input, mrb_stream_t stream);

/* wgrad: grad_weight[Cout][kh][kw][Cin] (fp32, KRSC == torch channels_last of [Cout,Cin,kh,kw]) =
 * sum over N,Ho,Wo of grad_output (x) input, on tcgen05 with MN-major operands; the pixel axis is split
 * across CTAs and reduced with red.global.add.f32 into the buffer, which is zeroed inside.  input and
 * grad_output are NHWC bf16. */
int mrb_conv2d_wgrad(const mrb_conv_params* p, const void* input_bf16, const void* grad_output_bf16,
                     const float* scale /* optional [Cout] */, float* grad_weight, mrb_stream_t stream);
/* Forward with the residual given at HALF resolution [N, ceil(Ho/2), ceil(Wo/2), Cout]: the epilogue reads it at
 * (h>>1, w>>1), i.e. FPN's `lateral(C_i) + nearest_upsample_2x(P_{i+1})` (modeling/backbone/fpn.py:59-64)
 * without materialising the upsampled map. */
int mrb_conv2d_fwd_up2(const mrb_conv_params* p, const void* input_bf16, const void* weight_bf16,
                       const float* scale, const float* bias, const void* residual_half, void* output,
                       mrb_stream_t stream);
/* grad_bias[c] = sum over pixels of an NHWC bf16 gradient [pixels, channels] (fp32, zeroed inside). */
int mrb_bias_grad(const void* grad_bf16_nhwc, float* grad_bias, long long pixels, int channels,
                  mrb_stream_t stream);
/* The same two reductions WITHOUT the zero-fill: the result is red.add-ed into the caller's fp32 gradient
 * accumulator (what autograd's AccumulateGrad node does to `param.grad`, torch/csrc/autograd/functions/
 * accumulate_grad.h, minus the temporary, its fill and the add pass). */
int mrb_conv2d_wgrad_accumulate(const mrb_conv_params* p, const void* input_bf16, const void* grad_output_bf16,
                                const float* scale /* optional [Cout] */, float* grad_weight, mrb_stream_t stream);
int mrb_bias_grad_accumulate(const void* grad_bf16_nhwc, float* grad_bias, long long pixels, int channels,
                             mrb_stream_t stream);

/* ---- pooling passes of the ResNet+FPN path (NHWC bf16, channels % 8 == 0) ---------------------------------------
 * mrb_max_pool_nhwc: F.max_pool2d(kernel, stride, pad) forward without an index tensor -- the stem pool of
 * modeling/backbone/resnet.py:307 (3, 2, 1; frozen, never differentiated) and FPN's P6 subsample (1, 2, 0;
 * modeling/backbone/fpn.py:77-79).  output is [N, (H+2p-k)/s+1, (W+2p-k)/s+1, C].
 * mrb_sum_pool2x2_nhwc: out[n, i, j, :] = sum of the (up to) 2x2 block of grad at (2i.., 2j..): the backward of the
 * nearest-2x upsample in FPN's top-down path (fpn.py:59-64).  out is [N, ceil(H/2), ceil(W/2), C]. */
int mrb_max_pool_nhwc(const void* input_bf16, void* output_bf16, int batch, int height, int width, int channels,
                      int kernel, int stride, int pad, mrb_stream_t stream);
int mrb_sum_pool2x2_nhwc(const void* grad_bf16, void* out_bf16, int batch, int height, int width, int channels,
                         mrb_stream_t stream);

/* ---- parameter update of the train step (reference solver/build.py:7-20 -> torch.optim.SGD) ----------------
 * One streaming pass over n contiguous fp32 parameters:
 *     d = grad * grad_scale + weight_decay * param;  m = momentum * m + d;  param -= lr * m
 * and, in the same pass, param_bf16 (optional) <- bf16(param)  -- the operand copy the conv engine reads --
 * and grad <- 0 when zero_grad != 0 (the accumulate entry points above add into it next step).
 * param/grad/momentum_buf 16-byte aligned, param_bf16 8-byte aligned.  26 B of HBM traffic per parameter. */
int mrb_sgd_momentum_step(float* param, float* grad, float* momentum_buf, void* param_bf16 /* optional */, long long n,
                          float lr, float momentum, float weight_decay, float grad_scale, int zero_grad,
                          mrb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MRB_B200_H_ */
