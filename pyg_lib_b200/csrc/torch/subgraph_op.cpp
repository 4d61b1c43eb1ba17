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

// pyg::relabel_neighborhood (relabel<disjoint>, cpu/dist_relabel_kernel.cpp:30-95) on CUDA tensors
std::tuple<at::Tensor, at::Tensor> relabel_neighborhood_cuda(const at::Tensor& seed, const at::Tensor& sampled_nodes_with_duplicates,
                                                             const std::vector<int64_t>& num_sampled_neighbors_per_node,
                                                             int64_t num_nodes, const std::optional<at::Tensor>& batch, bool csc,
                                                             bool disjoint) {
  const at::Tensor& sampled = sampled_nodes_with_duplicates;
  TORCH_CHECK(seed.is_cuda() && sampled.is_cuda(), "pyg_lib_b200: relabel_neighborhood expects CUDA tensors (no CPU fallback)");
  const auto st = seed.scalar_type();
  TORCH_CHECK(st == at::kLong || st == at::kInt, "relabel_neighborhood: index tensors must be int64 or int32");
  TORCH_CHECK(sampled.scalar_type() == st, "relabel_neighborhood: expected 'seed' and 'sampled_nodes_with_duplicates' to have the same dtype");
  if (disjoint) {   // dist_relabel_kernel.cpp:37-43
    TORCH_CHECK(batch.has_value(), "Batch needs to be specified to create disjoint subgraphs");
    TORCH_CHECK(batch->is_contiguous(), "Non-contiguous 'batch'");
    TORCH_CHECK(batch->numel() == sampled.numel(), "Each node must belong to a subgraph");
    TORCH_CHECK(batch->scalar_type() == st && batch->device() == seed.device(), "relabel_neighborhood: 'batch' must match 'seed' in dtype and device");
  }
  TORCH_CHECK(seed.is_contiguous(), "Non-contiguous 'seed'");
  TORCH_CHECK(sampled.is_contiguous(), "Non-contiguous 'sampled_nodes_with_duplicates'");
  TORCH_CHECK(sampled.device() == seed.device(), "relabel_neighborhood: tensors must live on one device");
  (void)num_nodes;  // (the reference only sizes its mapper with it)
  int64_t total = 0;
  for (const int64_t c : num_sampled_neighbors_per_node) {
    TORCH_CHECK(c >= 0, "relabel_neighborhood: negative neighbour count");
    total += c;
  }
  TORCH_CHECK(total <= sampled.numel(), "relabel_neighborhood: more neighbours counted than sampled nodes given");

  c10::cuda::CUDAGuard guard(seed.device());
  cudaStream_t stream = at::cuda::getCurrentCUDAStream();
  pygb200_subgraph* h = get_handle(seed.device().index(), stream);
  at::Tensor row = at::empty({total}, seed.options()), colv = at::empty({total}, seed.options());
  PYGB_TORCH_CALL(pygb200_relabel_neighborhood(h, seed.data_ptr(), seed.numel(), sampled.data_ptr(), disjoint ? batch->data_ptr() : nullptr,
                                               total, num_sampled_neighbors_per_node.data(), (int64_t)num_sampled_neighbors_per_node.size(),
                                               st == at::kInt, row.data_ptr(), colv.data_ptr(), stream));
  if (csc) std::swap(row, colv);   // get_sampled_edges, dist_relabel_kernel.cpp:16-27
  return std::make_tuple(row, colv);
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(pyg, m) {
  m.def(TORCH_SELECTIVE_SCHEMA(
      "pyg::subgraph(Tensor rowptr, Tensor col, Tensor "
      "nodes, bool return_edge_id) -> (Tensor, Tensor, Tensor?)"));
  m.def(TORCH_SELECTIVE_SCHEMA(   // pyg_lib/csrc/sampler/dist_relabel.cpp:71-76
      "pyg::relabel_neighborhood(Tensor seed, Tensor "
      "sampled_nodes_with_duplicates, int[] num_sampled_neighbors_per_node, "
      "int "
      "num_nodes, Tensor? batch = None, bool csc = False, bool disjoint = "
      "False) -> (Tensor, Tensor)"));
}

TORCH_LIBRARY_IMPL(pyg, CUDA, m) {
  m.impl(TORCH_SELECTIVE_NAME("pyg::subgraph"), TORCH_FN(subgraph_cuda));
  m.impl(TORCH_SELECTIVE_NAME("pyg::relabel_neighborhood"), TORCH_FN(relabel_neighborhood_cuda));
}

}  // namespace sampler
}  // namespace pyg
