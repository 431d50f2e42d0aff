// oracle/_ref build unit: compiles the reference's OWN CPU sources where they
// lie under $MRB_REFERENCE (default /root/reference) -- nothing is copied into
// this repository.  The only adaptation is a macro shim: torch >= 2.x removed
// the AT_DISPATCH overload that accepted `tensor.type()`
// (csrc/cpu/ROIAlign_cpu.cpp:242, csrc/cpu/nms_cpu.cpp:71), so the dispatch
// macro is re-pointed at an overload that extracts the ScalarType.  The
// arithmetic of the reference kernels is untouched.
#include <torch/extension.h>

namespace mrbref {
inline at::ScalarType st(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
inline at::ScalarType st(at::ScalarType t) { return t; }
}  // namespace mrbref

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::mrbref::st(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))

// resolved through -I $MRB_REFERENCE/maskrcnn_benchmark/csrc (see build_ref.py)
#include "cpu/ROIAlign_cpu.cpp"
#include "cpu/nms_cpu.cpp"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("nms", &nms_cpu, "reference csrc/cpu/nms_cpu.cpp");
  m.def("roi_align_forward", &ROIAlign_forward_cpu, "reference csrc/cpu/ROIAlign_cpu.cpp");
}
