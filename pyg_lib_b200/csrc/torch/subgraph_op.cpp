// torch.ops.pyg.subgraph on CUDA tensors — schema verbatim from pyg_lib/csrc/sampler/subgraph.cpp:28-32
// (the reference registers a CPU kernel only, cpu/subgraph_kernel.cpp:93-96; here the CUDA key, no CPU fallback).
#include <map>
#include <mutex>

#include "common.h"

namespace pyg {
namespace sampler {
namespace {

// one persistent workspace per (device, stream)
pygb200_subgraph* get_handle(int device, cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, pygb200_subgraph*> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(device, stream);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  pygb200_subgraph* h = nullptr;
  PYGB_TORCH_CALL(pygb200_subgraph_create(&h));
  cache[key] = h;
  return h;
}

std::tuple<at::Tensor, at::Tensor, std::optional<at::Tensor>> subgraph_cuda(const at::Tensor& rowptr, const at::Tensor& col,
                                                                            const at::Tensor& nodes, bool return_edge_id) {
  TORCH_CHECK(rowptr.is_cuda() && col.is_cuda() && nodes.is_cuda(),
              "pyg_lib_b200: subgraph expects CUDA tensors (no CPU fallback)");
  const auto st = nodes.scalar_type();
  TORCH_CHECK(st == at::kLong || st == at::kInt, "subgraph: index tensors must be int64 or int32");
  // at::checkAllSameType of the reference's wrapper (subgraph.cpp:18-20) + what the kernel assumes
  TORCH_CHECK(rowptr.scalar_type() == st && col.scalar_type() == st, "subgraph: expected 'rowptr', 'col' and 'nodes' to have the same dtype");
  TORCH_CHECK(rowptr.device() == nodes.device() && col.device() == nodes.device(), "subgraph: tensors must live on one device");
  TORCH_CHECK(rowptr.dim() == 1 && col.dim() == 1 && nodes.dim() == 1, "subgraph: expected one-dimensional tensors");
  TORCH_CHECK(rowptr.is_contiguous() && col.is_contiguous() && nodes.is_contiguous(), "subgraph: expected contiguous tensors");
  TORCH_CHECK(rowptr.numel() >= 1, "'rowptr' must have at least one element");

  c10::cuda::CUDAGuard guard(nodes.device());
  cudaStream_t stream = at::cuda::getCurrentCUDAStream();
  pygb200_subgraph* h = get_handle(nodes.device().index(), stream);
  const int idx32 = st == at::kInt;
  const int64_t n = nodes.numel();
  at::Tensor out_rowptr = rowptr.new_empty({n + 1});
  int64_t n_edges = 0;
  PYGB_TORCH_CALL(pygb200_subgraph_count(h, rowptr.data_ptr(), col.data_ptr(), rowptr.numel() - 1, nodes.data_ptr(), n, idx32,
                                         out_rowptr.data_ptr(), &n_edges, stream));
  at::Tensor out_col = col.new_empty({n_edges});
  std::optional<at::Tensor> out_eid = std::nullopt;
  if (return_edge_id) out_eid = col.new_empty({n_edges});
  PYGB_TORCH_CALL(pygb200_subgraph_fill(h, rowptr.data_ptr(), col.data_ptr(), nodes.data_ptr(), n, idx32, out_col.data_ptr(),
                                        return_edge_id ? out_eid->data_ptr() : nullptr, stream));
  return std::make_tuple(out_rowptr, out_col, out_eid);
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(pyg, m) {
  m.def(TORCH_SELECTIVE_SCHEMA(
      "pyg::subgraph(Tensor rowptr, Tensor col, Tensor "
      "nodes, bool return_edge_id) -> (Tensor, Tensor, Tensor?)"));
}

TORCH_LIBRARY_IMPL(pyg, CUDA, m) {
  m.impl(TORCH_SELECTIVE_NAME("pyg::subgraph"), TORCH_FN(subgraph_cuda));
}

}  // namespace sampler
}  // namespace pyg
