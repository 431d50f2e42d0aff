// oracle/ref_cuda/dpool_host.cu -- see shim_common.h
#include "shim_common.h"
#include "cuda/deform_pool_cuda.cu"
