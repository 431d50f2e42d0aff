// oracle/ref_cuda/roi_nms_focal.cu -- see shim_common.h
#include "shim_common.h"
#include "cuda/ROIAlign_cuda.cu"
#include "cuda/ROIPool_cuda.cu"
#include "cuda/SigmoidFocalLoss_cuda.cu"
#include "cuda/nms.cu"
