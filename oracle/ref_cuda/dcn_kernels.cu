// oracle/ref_cuda/dcn_kernels.cu -- see shim_common.h
#include "shim_common.h"
#include "cuda/deform_conv_kernel_cuda.cu"
