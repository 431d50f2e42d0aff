// oracle/ref_cuda: build units that compile the reference's OWN CUDA sources (csrc/cuda/*.cu) where they lie under
// $MRB_REFERENCE -- nothing is copied into this repository and no kernel arithmetic is touched.  Adaptations to a current
// PyTorch, all in this header / oracle/thc_compat:
//   * AT_DISPATCH_* re-pointed at an overload accepting `tensor.type()` (the DeprecatedTypeProperties overload is gone),
//   * AT_CHECK -> TORCH_CHECK (renamed), <THC/...> -> oracle/thc_compat,
//   * at::globalContext().lazyInitCUDA() returned a THCState* (csrc/cuda/nms.cu:83); it returns void now.
// The result (oracle/_ref/mrb_ref_cuda.so) is the GPU-side CHECKER and the "reference CUDA kernels on this box" baseline of
// tools/bench_ops.py.  Test / baseline infrastructure only.
#pragma once
#include <torch/extension.h>
#include <ATen/Dispatch.h>
#include <ATen/cuda/CUDAContext.h>
#include <THC/THC.h>

namespace mrbref {
inline at::ScalarType st(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
inline at::ScalarType st(at::ScalarType t) { return t; }
struct Ctx {
  THCState* lazyInitCUDA() { at::globalContext().lazyInitDevice(c10::DeviceType::CUDA); return nullptr; }
};
}  // namespace mrbref
namespace at { inline ::mrbref::Ctx mrbref_ctx() { return {}; } }

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::mrbref::st(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
#undef AT_DISPATCH_FLOATING_TYPES_AND_HALF
#define AT_DISPATCH_FLOATING_TYPES_AND_HALF(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::mrbref::st(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES_AND(at::ScalarType::Half, __VA_ARGS__))
#ifndef AT_CHECK
#define AT_CHECK TORCH_CHECK
#endif
#define globalContext() mrbref_ctx()
