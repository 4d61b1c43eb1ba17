// C++ caller of libpyg.so's PYGB200_API entry points (pyg_lib_b200/csrc/torch/api.h): argument checks raise c10::Error
// before any kernel is reached; valid calls go through the dispatcher — to the clear "no CPU fallback" error with CPU
// tensors, to the CUDA kernels with `--gpu`.  Built and run by tests/test_cpp_api.py.
#include <torch/torch.h>

#include <cstring>
#include <iostream>

#include "api.h"

template <typename F>
static bool throws(F f, const char* needle) {
  try { f(); } catch (const c10::Error& e) { return std::strstr(e.what(), needle) != nullptr; }
  return false;
}

int main(int argc, char** argv) {
  const bool gpu = argc > 1 && std::strcmp(argv[1], "--gpu") == 0;
  int bad = 0;
  auto x = torch::randn({8, 16}), w = torch::randn({2, 16, 32});
  auto ptr = torch::tensor({0, 5, 8}, torch::kLong);
  bad += !throws([&] { pyg::ops::segment_matmul(x, ptr, w.to(torch::kDouble)); }, "same dtype");
  bad += !throws([&] { pyg::ops::segment_matmul(x, ptr, w.slice(1, 0, 8)); }, "other.size(1) == input.size(-1)");
  bad += !throws([&] { pyg::ops::segment_matmul(x, torch::tensor({0, 8}, torch::kLong), w); }, "ptr.numel()");
  bad += !throws([&] { pyg::ops::segment_matmul(x, ptr, at::Tensor()); }, "undefined");
  bad += !throws([&] { pyg::ops::grouped_matmul({x, x}, {w[0]}); }, "must match number");
  bad += !throws([&] { pyg::ops::grouped_matmul({x}, {torch::randn({15, 4})}); }, "size(0)");
  auto rowptr = torch::tensor({0, 2, 4, 6}, torch::kLong), col = torch::tensor({1, 2, 0, 2, 0, 1}, torch::kLong);
  bad += !throws([&] { pyg::sampler::neighbor_sample(rowptr, col, torch::tensor({0}, torch::kInt), {2}); }, "same dtype");
  bad += !throws([&] { pyg::sampler::neighbor_sample(rowptr, at::Tensor(), torch::tensor({0}, torch::kLong), {2}); }, "undefined");
  if (!gpu) {   // valid arguments, CPU tensors: the dispatcher is reached and says what this build is
    bad += !throws([&] { pyg::ops::segment_matmul(x, ptr, w); }, "no CPU fallback");
    bad += !throws([&] { pyg::sampler::neighbor_sample(rowptr, col, torch::tensor({0}, torch::kLong), {2}); }, "no CPU fallback");
  } else {
    auto dev = torch::Device(torch::kCUDA, 0);
    auto out = pyg::ops::segment_matmul(x.to(dev), ptr.to(dev), w.to(dev)).cpu();
    auto ref = torch::cat({x.slice(0, 0, 5).matmul(w[0]), x.slice(0, 5, 8).matmul(w[1])});
    bad += !torch::allclose(out, ref, 1e-4, 1e-4);
    auto outs = pyg::ops::grouped_matmul({x.to(dev), x.to(dev)}, {w[0].to(dev), w[1].to(dev)});
    bad += !torch::allclose(outs[1].cpu(), x.matmul(w[1]), 1e-4, 1e-4);
    auto xg = x.to(dev).requires_grad_(true);   // the Autograd kernel is behind the same entry point
    pyg::ops::segment_matmul(xg, ptr.to(dev), w.to(dev)).sum().backward();
    bad += !xg.grad().defined();
    auto r = pyg::sampler::neighbor_sample(rowptr.to(dev), col.to(dev), torch::tensor({0}, torch::kLong).to(dev), {-1, -1});
    bad += !(std::get<2>(r).numel() == 3 && std::get<4>(r).size() == 3);
  }
  std::cout << (bad ? "CPP_API_FAIL " : "CPP_API_OK ") << bad << std::endl;
  return bad ? 1 : 0;
}
