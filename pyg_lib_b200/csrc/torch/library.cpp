// `pyg::cuda_version` — same schema as the reference (pyg_lib/csrc/library.cpp:19-29).
#include "common.h"

namespace pyg {

int64_t cuda_version() { return pygb200_cuda_version(); }

// process-wide count of kernels launched by libpyg_b200.so (bench.py's `gpu_launches`)
int64_t b200_kernel_launches() { return pygb200_kernel_launches(); }

TORCH_LIBRARY_FRAGMENT(pyg, m) {
  m.def("cuda_version", &cuda_version);
  m.def("b200_kernel_launches", &b200_kernel_launches);
}

}  // namespace pyg
