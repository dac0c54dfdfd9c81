// TEST INFRASTRUCTURE.  pybind shim around the REFERENCE's own CPU implementation of the rotated BEV IoU
// (/root/reference/pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:222-252 boxes_iou_bev_cpu), compiled from where it lies by
// oracle/build_oracle.py into oracle/_ref/ (git-ignored).  The reference binds the same function in
// iou3d_nms_api.cpp:11-17 next to its CUDA entry points, which cannot be built without a GPU toolchain target here.
#include <torch/extension.h>

int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) { m.def("boxes_iou_bev_cpu", &boxes_iou_bev_cpu, "oriented boxes iou (reference CPU path)"); }
