// oracle/ref_cuda/dpool_kernels.cu -- see shim_common.h
#include "shim_common.h"
#include "cuda/deform_pool_kernel_cuda.cu"
