// api_misc.cu -- version / error strings of the C ABI.
#include "common.cuh"

MRB_API int mrb_version(void) { return 100; }  // 0.1.0

MRB_API const char* mrb_error_string(int code) {
  switch (code) {
    case MRB_OK: return "ok";
    case MRB_ERR_BAD_ARG: return "mrb: bad argument";
    case MRB_ERR_UNSUPPORTED: return "mrb: unsupported shape/dtype/layout for the sm_100a kernels";
    case MRB_ERR_WORKSPACE: return "mrb: workspace too small";
    case MRB_ERR_DRIVER: return "mrb: CUDA driver entry point unavailable (cuTensorMapEncodeTiled)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "mrb: unknown error";
}
