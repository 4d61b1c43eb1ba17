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

// pyg::hetero_relabel_neighborhood (hetero relabel<disjoint>, cpu/dist_relabel_kernel.cpp:97-273).
// Every destination type's sampled list is consumed front to back (`slice_dict`, `:206-235`), so a node's id is its
// rank among the first occurrences of [that type's seeds | that type's sampled list] — one pygb200_relabel_ids per
// node type.  What remains is the schedule (which positions belong to which layer / edge type / source node, and the
// source node's local index), pure bookkeeping over the int[][] count lists: done here on the host exactly in the
// reference's loop order, then one pygb200_relabel_expand per edge type.
typedef std::string node_type;
typedef std::string rel_type;
typedef std::tuple<std::string, std::string, std::string> edge_type;

std::tuple<c10::Dict<rel_type, at::Tensor>, c10::Dict<rel_type, at::Tensor>> hetero_relabel_neighborhood_cuda(
    const std::vector<node_type>& node_types, const std::vector<edge_type>& edge_types, const c10::Dict<node_type, at::Tensor>& seed_dict,
    const c10::Dict<node_type, at::Tensor>& sampled_nodes_with_duplicates_dict,
    const c10::Dict<rel_type, std::vector<std::vector<int64_t>>>& num_sampled_neighbors_per_node_dict,
    const c10::Dict<node_type, int64_t>& num_nodes_dict, const std::optional<c10::Dict<node_type, at::Tensor>>& batch_dict, bool csc,
    bool disjoint) {
  TORCH_CHECK(seed_dict.size() > 0, "hetero_relabel_neighborhood: empty 'seed_dict'");
  TORCH_CHECK(!edge_types.empty(), "hetero_relabel_neighborhood: empty 'edge_types'");
  const at::Tensor& first = seed_dict.begin()->value();
  TORCH_CHECK(first.is_cuda(), "pyg_lib_b200: hetero_relabel_neighborhood expects CUDA tensors (no CPU fallback)");
  const auto st = first.scalar_type();
  const auto dev = first.device();
  TORCH_CHECK(st == at::kLong || st == at::kInt, "hetero_relabel_neighborhood: index tensors must be int64 or int32");
  if (disjoint) TORCH_CHECK(batch_dict.has_value(), "Batch needs to be specified to create disjoint subgraphs");
  (void)num_nodes_dict;
  auto rel_of = [](const edge_type& k) { return std::get<0>(k) + "__" + std::get<1>(k) + "__" + std::get<2>(k); };
  auto check = [&](const at::Tensor& t, const char* what) {
    TORCH_CHECK(t.is_contiguous(), "Non-contiguous '", what, "'");
    TORCH_CHECK(t.scalar_type() == st && t.device() == dev && t.dim() == 1, "hetero_relabel_neighborhood: '", what,
                "' must be one-dimensional and match the seeds in dtype and device");
  };

  c10::cuda::CUDAGuard guard(dev);
  cudaStream_t stream = at::cuda::getCurrentCUDAStream();
  pygb200_subgraph* h = get_handle(dev.index(), stream);
  const int idx32 = st == at::kInt;

  // ---- ids per node type; disjoint batch ids of the seeds run over seed_dict in its order (`:180-193`)
  std::map<node_type, int64_t> batch0;
  int64_t nb = 0;
  for (const auto& kv : seed_dict) { batch0[kv.key()] = nb; nb += kv.value().numel(); }
  std::map<node_type, at::Tensor> ids;
  for (const auto& t : node_types) {
    TORCH_CHECK(sampled_nodes_with_duplicates_dict.contains(t), "hetero_relabel_neighborhood: no sampled nodes for node type '", t, "'");
    const at::Tensor sampled = sampled_nodes_with_duplicates_dict.at(t);
    check(sampled, "sampled_nodes_with_duplicates");
    at::Tensor seed;
    if (seed_dict.contains(t)) { seed = seed_dict.at(t); check(seed, "seed"); }
    at::Tensor batch;
    if (disjoint) {
      TORCH_CHECK(batch_dict->contains(t), "hetero_relabel_neighborhood: no batch vector for node type '", t, "'");
      batch = batch_dict->at(t);
      check(batch, "batch");
      TORCH_CHECK(batch.numel() == sampled.numel(), "Each node must belong to a subgraph");
    }
    at::Tensor out = at::empty({sampled.numel()}, sampled.options());
    PYGB_TORCH_CALL(pygb200_relabel_ids(h, seed.defined() ? seed.data_ptr() : nullptr, seed.defined() ? seed.numel() : 0,
                                        batch0.count(t) ? batch0[t] : 0, sampled.data_ptr(), disjoint ? batch.data_ptr() : nullptr,
                                        sampled.numel(), idx32, out.data_ptr(), stream));
    ids[t] = out;
  }

  // ---- the schedule (`:195-259`)
  const size_t R = edge_types.size();
  std::vector<const std::vector<std::vector<int64_t>>*> counts(R);
  std::vector<std::vector<std::vector<int64_t>>> counts_store(R);
  for (size_t r = 0; r < R; ++r) {
    const rel_type rk = rel_of(edge_types[r]);
    TORCH_CHECK(num_sampled_neighbors_per_node_dict.contains(rk), "hetero_relabel_neighborhood: no neighbour counts for '", rk, "'");
    counts_store[r] = num_sampled_neighbors_per_node_dict.at(rk);
    counts[r] = &counts_store[r];
    TORCH_CHECK(!counts[r]->empty(), "hetero_relabel_neighborhood: empty layer list for '", rk, "'");
  }
  const size_t L = counts[0]->size();
  for (size_t r = 0; r < R; ++r) TORCH_CHECK(counts[r]->size() >= L, "hetero_relabel_neighborhood: layer lists differ in length");
  std::map<node_type, int64_t> cursor, src_off;
  for (const auto& t : node_types) { cursor[t] = 0; src_off[t] = 0; }
  std::vector<std::pair<int64_t, int64_t>> src_slice(R);
  for (size_t r = 0; r < R; ++r) src_slice[r] = {0, (int64_t)(*counts[r])[0].size()};
  std::vector<std::vector<int64_t>> seg_count(R), seg_row(R), seg_pos(R);
  for (size_t ell = 0; ell < L; ++ell) {
    for (size_t r = 0; r < R; ++r) {
      const auto& k = edge_types[r];
      const node_type& dst = !csc ? std::get<2>(k) : std::get<0>(k);
      TORCH_CHECK(cursor.count(dst), "hetero_relabel_neighborhood: edge type with unknown node type '", dst, "'");
      const auto& cl = (*counts[r])[ell];
      const int64_t begin_i = src_slice[r].first, end_i = src_slice[r].second;
      TORCH_CHECK((int64_t)cl.size() >= end_i - begin_i, "hetero_relabel_neighborhood: too few neighbour counts in a layer");
      for (int64_t i = begin_i; i < end_i; ++i) {
        const int64_t c = cl[(size_t)(i - begin_i)];
        TORCH_CHECK(c >= 0, "hetero_relabel_neighborhood: negative neighbour count");
        if (c > 0) { seg_count[r].push_back(c); seg_row[r].push_back(i); seg_pos[r].push_back(cursor[dst]); }
        cursor[dst] += c;
      }
    }
    if (ell + 1 < L) {
      for (size_t r = 0; r < R; ++r) {
        const node_type& src = !csc ? std::get<0>(edge_types[r]) : std::get<2>(edge_types[r]);
        src_off[src] = std::max(src_off[src], src_slice[r].second);
      }
      for (size_t r = 0; r < R; ++r) {
        const node_type& src = !csc ? std::get<0>(edge_types[r]) : std::get<2>(edge_types[r]);
        src_slice[r] = {src_off[src], src_off[src] + (int64_t)(*counts[r])[ell + 1].size()};
      }
    }
  }
  for (const auto& t : node_types)
    TORCH_CHECK(cursor[t] <= ids[t].numel(), "hetero_relabel_neighborhood: more neighbours counted than sampled nodes given for '", t, "'");

  // ---- rows / cols per edge type
  c10::Dict<rel_type, at::Tensor> out_row, out_col;
  for (size_t r = 0; r < R; ++r) {
    const auto& k = edge_types[r];
    const node_type& dst = !csc ? std::get<2>(k) : std::get<0>(k);
    int64_t total = 0;
    for (const int64_t c : seg_count[r]) total += c;
    at::Tensor row = at::empty({total}, first.options()), colv = at::empty({total}, first.options());
    PYGB_TORCH_CALL(pygb200_relabel_expand(h, seg_count[r].data(), seg_row[r].data(), seg_pos[r].data(), (int64_t)seg_count[r].size(),
                                           ids[dst].data_ptr(), ids[dst].numel(), idx32, row.data_ptr(), colv.data_ptr(), stream));
    if (csc) std::swap(row, colv);   // get_sampled_edges, `:16-27`
    out_row.insert(rel_of(k), row);
    out_col.insert(rel_of(k), colv);
  }
  return std::make_tuple(out_row, out_col);
}

// pyg::merge_sampler_outputs (merge_outputs<disjoint>, cpu/dist_merge_outputs_kernel.cpp:15-137) on CUDA tensors: the
// plan is host arithmetic over the int[] / int[][] arguments (pygb200_merge_plan), the data moves in three segment
// gathers (node ids, edge ids, batch broadcast).  The reference pads every seed's slot with -1 and strips all -1
// afterwards; ids are never -1, so the result is the plain concatenation.
std::tuple<at::Tensor, at::Tensor, std::optional<at::Tensor>, std::vector<int64_t>> merge_sampler_outputs_cuda(
    const std::vector<at::Tensor>& node_ids, const std::vector<at::Tensor>& edge_ids,
    const std::vector<std::vector<int64_t>>& cumsum_neighbors_per_node, const std::vector<int64_t>& partition_ids,
    const std::vector<int64_t>& partition_orders, int64_t num_partitions, int64_t num_neighbors, const std::optional<at::Tensor>& batch,
    bool disjoint) {
  (void)num_neighbors;   // (only sizes the reference's padded scratch)
  TORCH_CHECK(partition_ids.size() == partition_orders.size(), "Every partition ID must be assigned a sampling order");
  if (disjoint) TORCH_CHECK(batch.has_value(), "Disjoint sampling requires 'batch' to be specified");
  const int64_t P = num_partitions;
  TORCH_CHECK(P >= 1 && (int64_t)node_ids.size() >= P && (int64_t)edge_ids.size() >= P && (int64_t)cumsum_neighbors_per_node.size() >= P,
              "merge_sampler_outputs: expected one node / edge tensor and one cumulative count list per partition");
  TORCH_CHECK(node_ids[0].is_cuda(), "pyg_lib_b200: merge_sampler_outputs expects CUDA tensors (no CPU fallback)");
  const auto st = node_ids[0].scalar_type();
  const auto dev = node_ids[0].device();
  TORCH_CHECK(st == at::kLong || st == at::kInt, "merge_sampler_outputs: index tensors must be int64 or int32");
  auto check = [&](const at::Tensor& t, const char* what) {
    TORCH_CHECK(t.defined() && t.is_contiguous() && t.dim() == 1 && t.scalar_type() == st && t.device() == dev, "merge_sampler_outputs: '",
                what, "' must be contiguous, one-dimensional, of one dtype and on one device");
  };
  const int64_t n = (int64_t)partition_ids.size();
  std::vector<int64_t> flat, off(P + 1, 0), node_numel(P), edge_numel(P);
  std::vector<const void*> node_ptrs(P), edge_ptrs(P);
  for (int64_t p = 0; p < P; ++p) {
    check(node_ids[p], "node_ids"); check(edge_ids[p], "edge_ids");
    node_numel[p] = node_ids[p].numel(); edge_numel[p] = edge_ids[p].numel();
    node_ptrs[p] = node_ids[p].data_ptr(); edge_ptrs[p] = edge_ids[p].data_ptr();
    flat.insert(flat.end(), cumsum_neighbors_per_node[p].begin(), cumsum_neighbors_per_node[p].end());
    off[p + 1] = (int64_t)flat.size();
  }
  std::vector<int64_t> nb(n), eb(n), cnt(n);
  PYGB_TORCH_CALL(pygb200_merge_plan(flat.data(), off.data(), P, partition_ids.data(), partition_orders.data(), n, node_numel.data(),
                                     edge_numel.data(), nb.data(), eb.data(), cnt.data()));
  int64_t total = 0;
  for (const int64_t c : cnt) total += c;

  c10::cuda::CUDAGuard guard(dev);
  cudaStream_t stream = at::cuda::getCurrentCUDAStream();
  pygb200_subgraph* h = get_handle(dev.index(), stream);
  const int idx32 = st == at::kInt;
  at::Tensor out_node = at::empty({total}, node_ids[0].options()), out_edge = at::empty({total}, node_ids[0].options());
  PYGB_TORCH_CALL(pygb200_segment_gather(h, node_ptrs.data(), P, partition_ids.data(), nb.data(), cnt.data(), n, 0, idx32, out_node.data_ptr(), stream));
  PYGB_TORCH_CALL(pygb200_segment_gather(h, edge_ptrs.data(), P, partition_ids.data(), eb.data(), cnt.data(), n, 0, idx32, out_edge.data_ptr(), stream));
  std::optional<at::Tensor> out_batch = std::nullopt;
  if (disjoint) {
    check(*batch, "batch");
    TORCH_CHECK(batch->numel() >= n, "merge_sampler_outputs: 'batch' needs one entry per seed");
    out_batch = at::empty({total}, node_ids[0].options());
    std::vector<int64_t> zero(n, 0), pos(n);
    for (int64_t j = 0; j < n; ++j) pos[j] = j;
    const void* bp[1] = {batch->data_ptr()};
    PYGB_TORCH_CALL(pygb200_segment_gather(h, bp, 1, zero.data(), pos.data(), cnt.data(), n, 1, idx32, out_batch->data_ptr(), stream));
  }
  return std::make_tuple(out_node, out_edge, out_batch, cnt);
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
  m.def(TORCH_SELECTIVE_SCHEMA(   // pyg_lib/csrc/sampler/dist_relabel.cpp:77-83
      "pyg::hetero_relabel_neighborhood(str[] node_types, (str, str, str)[] "
      "edge_types, Dict(str, Tensor) seed_dict, Dict(str, Tensor) "
      "sampled_nodes_with_duplicates_dict, Dict(str, int[][]) "
      "num_sampled_neighbors_per_node_dict, Dict(str, int) num_nodes_dict, "
      "Dict(str, Tensor)? batch_dict = None, bool csc = False, bool disjoint = "
      "False) -> (Dict(str, Tensor), Dict(str, Tensor))"));
  m.def(TORCH_SELECTIVE_SCHEMA(   // pyg_lib/csrc/sampler/dist_merge_outputs.cpp:51-55
      "pyg::merge_sampler_outputs(Tensor[] node_ids, Tensor[] edge_ids, "
      "int[][] cumsum_neighbors_per_node, int[] partition_ids, int[] "
      "partition_orders, int num_partitions, int num_neighbors, Tensor? "
      "batch, bool disjoint) -> (Tensor, Tensor, Tensor?, int[])"));
}

TORCH_LIBRARY_IMPL(pyg, CUDA, m) {
  m.impl(TORCH_SELECTIVE_NAME("pyg::subgraph"), TORCH_FN(subgraph_cuda));
  m.impl(TORCH_SELECTIVE_NAME("pyg::relabel_neighborhood"), TORCH_FN(relabel_neighborhood_cuda));
}

// dict arguments carry no backend key (as for hetero_neighbor_sample): BackendSelect, the kernel checks devices itself
TORCH_LIBRARY_IMPL(pyg, BackendSelect, m) {
  m.impl(TORCH_SELECTIVE_NAME("pyg::hetero_relabel_neighborhood"), TORCH_FN(hetero_relabel_neighborhood_cuda));
  m.impl(TORCH_SELECTIVE_NAME("pyg::merge_sampler_outputs"), TORCH_FN(merge_sampler_outputs_cuda));
}

}  // namespace sampler
}  // namespace pyg
