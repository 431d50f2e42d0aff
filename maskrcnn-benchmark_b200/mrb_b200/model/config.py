"""The handful of reference config values the R-50-FPN configs use (config/defaults.py and
configs/e2e_{faster,mask}_rcnn_R_50_FPN_1x.yaml), as a plain dataclass."""
from dataclasses import dataclass, field
from typing import Tuple


@dataclass
class RCNNConfig:
    mask_on: bool = True                      # MODEL.MASK_ON (e2e_mask_rcnn_R_50_FPN_1x.yaml:4)
    num_classes: int = 81                     # ROI_BOX_HEAD.NUM_CLASSES (defaults.py:229)
    # backbone (R-50-FPN): defaults.py:97-108, 245-278
    stage_blocks: Tuple[int, ...] = (3, 4, 6, 3)
    stem_out: int = 64
    res2_out: int = 256
    width_per_group: int = 64
    num_groups: int = 1                       # RESNETS.NUM_GROUPS (32 for X-101-32x8d, with width_per_group 8)
    stride_in_1x1: bool = True
    stage_with_dcn: Tuple[bool, ...] = (False, False, False, False)   # RESNETS.STAGE_WITH_DCN (configs/dcn: F,T,T,T)
    with_modulated_dcn: bool = False          # RESNETS.WITH_MODULATED_DCN
    freeze_at: int = 2                        # BACKBONE.FREEZE_CONV_BODY_AT
    fpn_out: int = 256                        # RESNETS.BACKBONE_OUT_CHANNELS
    # RPN: defaults.py:128-175, yaml:9-13
    anchor_sizes: Tuple[int, ...] = (32, 64, 128, 256, 512)
    anchor_strides: Tuple[int, ...] = (4, 8, 16, 32, 64)
    aspect_ratios: Tuple[float, ...] = (0.5, 1.0, 2.0)
    straddle_thresh: int = 0
    rpn_fg_iou: float = 0.7
    rpn_bg_iou: float = 0.3
    rpn_batch_size: int = 256
    rpn_positive_fraction: float = 0.5
    pre_nms_top_n_train: int = 2000
    pre_nms_top_n_test: int = 1000
    post_nms_top_n_train: int = 2000
    post_nms_top_n_test: int = 1000
    fpn_post_nms_top_n_train: int = 2000
    fpn_post_nms_top_n_test: int = 1000       # yaml:13 (e2e_mask: 1000)
    fpn_post_nms_per_batch: bool = True
    rpn_nms_thresh: float = 0.7
    rpn_min_size: int = 0
    # ROI heads: defaults.py:181-241, yaml:14-31
    roi_fg_iou: float = 0.5
    roi_bg_iou: float = 0.5
    bbox_reg_weights: Tuple[float, ...] = (10.0, 10.0, 5.0, 5.0)
    roi_batch_size: int = 512
    roi_positive_fraction: float = 0.25
    score_thresh: float = 0.05
    roi_nms: float = 0.5
    detections_per_img: int = 100
    pooler_scales: Tuple[float, ...] = (0.25, 0.125, 0.0625, 0.03125)
    box_resolution: int = 7
    box_sampling_ratio: int = 2
    mlp_head_dim: int = 1024
    mask_resolution_pool: int = 14
    mask_sampling_ratio: int = 2
    mask_conv_layers: Tuple[int, ...] = (256, 256, 256, 256)
    mask_resolution: int = 28
    # 0: the mask head runs on the gathered positive ROIs (one `nonzero` host sync per step, as the reference's
    # keep_only_positive_boxes).  > 0: fixed-shape variant -- the first `mask_rois_per_image` sampled ROIs of each
    # image in positives-first order, non-positives weighted 0 in the loss (identical loss value; the sampler never
    # yields more than int(roi_batch_size * roi_positive_fraction) = 128 positives per image) -- which makes the
    # whole train step free of host synchronisation and therefore CUDA-graph capturable.
    mask_rois_per_image: int = 0
    parallel_heads: bool = False   # issue the mask branch on its own CUDA stream (overlaps the box branch, fwd and bwd)
    size_divisibility: int = 32               # DATALOADER.SIZE_DIVISIBILITY (yaml:35-36)
    pixel_mean: Tuple[float, ...] = field(default=(102.9801, 115.9465, 122.7717))
