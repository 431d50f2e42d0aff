// oracle/ref_cuda/bind.cpp -- python names of the reference's CUDA entry points (csrc/cuda/vision.h)
#include <torch/extension.h>
#include "cuda/vision.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("nms", &nms_cuda, "csrc/cuda/nms.cu (boxes [N,5] sorted by score)");
  m.def("roi_align_forward", &ROIAlign_forward_cuda);
  m.def("roi_align_backward", &ROIAlign_backward_cuda);
  m.def("roi_pool_forward", &ROIPool_forward_cuda);
  m.def("roi_pool_backward", &ROIPool_backward_cuda);
  m.def("sigmoid_focalloss_forward", &SigmoidFocalLoss_forward_cuda);
  m.def("sigmoid_focalloss_backward", &SigmoidFocalLoss_backward_cuda);
  m.def("deform_conv_forward", &deform_conv_forward_cuda);
  m.def("deform_conv_backward_input", &deform_conv_backward_input_cuda);
  m.def("deform_conv_backward_parameters", &deform_conv_backward_parameters_cuda);
  m.def("modulated_deform_conv_forward", &modulated_deform_conv_cuda_forward);
  m.def("modulated_deform_conv_backward", &modulated_deform_conv_cuda_backward);
  m.def("deform_psroi_pooling_forward", &deform_psroi_pooling_cuda_forward);
  m.def("deform_psroi_pooling_backward", &deform_psroi_pooling_cuda_backward);
}
