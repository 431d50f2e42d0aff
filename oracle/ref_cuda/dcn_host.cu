// oracle/ref_cuda/dcn_host.cu -- see shim_common.h
#include "shim_common.h"
#include "cuda/deform_conv_cuda.cu"
