"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pyg_lib_b200) never imports this.

The functions mirror the reference's Python API (pyg_lib/sampler/__init__.py:11-200,
pyg_lib/ops/__init__.py:99-172) on CPU tensors and consume / advance torch's default CPU
generator exactly like the reference does (whole 128-word blocks of at::randint, see
pyg_lib/csrc/random/cpu/rand_engine.h:80-92), so `torch.manual_seed(s)` followed by the same
call sequence gives the same outputs as the reference.
"""
import ctypes as C
import os
import os.path as osp
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

_HERE = osp.dirname(osp.abspath(__file__))
_LIB = None


class MTState(C.Structure):
    _fields_ = [('state', C.c_uint32 * 624), ('left', C.c_int32), ('next', C.c_int32)]


class HomoResult(C.Structure):
    _fields_ = [('n_edges', C.c_int64), ('n_nodes', C.c_int64),
                ('row', C.POINTER(C.c_int64)), ('col', C.POINTER(C.c_int64)),
                ('edge_id', C.POINTER(C.c_int64)), ('node_id', C.POINTER(C.c_int64)),
                ('node_batch', C.POINTER(C.c_int64)), ('nodes_per_hop', C.POINTER(C.c_int64)),
                ('edges_per_hop', C.POINTER(C.c_int64)), ('rng_blocks', C.c_int64)]


class HeteroResult(C.Structure):
    _fields_ = [('T', C.c_int64), ('R', C.c_int64), ('L', C.c_int64),
                ('n_nodes', C.POINTER(C.c_int64)),
                ('node_id', C.POINTER(C.POINTER(C.c_int64))),
                ('node_batch', C.POINTER(C.POINTER(C.c_int64))),
                ('nodes_per_hop', C.POINTER(C.c_int64)),
                ('n_edges', C.POINTER(C.c_int64)),
                ('row', C.POINTER(C.POINTER(C.c_int64))),
                ('col', C.POINTER(C.POINTER(C.c_int64))),
                ('edge_id', C.POINTER(C.POINTER(C.c_int64))),
                ('edges_per_hop', C.POINTER(C.c_int64)), ('rng_blocks', C.c_int64)]


def build(force: bool = False) -> str:
    path = osp.join(_HERE, 'liboracle.so')
    srcs = [osp.join(_HERE, f) for f in ('sampler_oracle.c', 'matmul_oracle.c')]
    if force or not osp.exists(path) or any(osp.getmtime(s) > osp.getmtime(path) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, 'liboracle.so'], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return path


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_mt_seed.argtypes = [C.POINTER(MTState), C.c_uint64]
        _LIB.oracle_randint_words.argtypes = [C.POINTER(MTState), C.c_void_p, C.c_int64]
        _LIB.oracle_neighbor_sample.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                                C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                C.POINTER(MTState), C.POINTER(HomoResult)]
        _LIB.oracle_neighbor_sample_temporal.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                                         C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(MTState),
                                                         C.POINTER(HomoResult), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.oracle_hetero_neighbor_sample_temporal.argtypes = [
            C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(MTState), C.POINTER(HeteroResult),
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.oracle_homo_free.argtypes = [C.POINTER(HomoResult)]
        _LIB.oracle_hetero_neighbor_sample.argtypes = [
            C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(MTState), C.POINTER(HeteroResult)]
        _LIB.oracle_hetero_free.argtypes = [C.POINTER(HeteroResult)]
        _LIB.oracle_segment_matmul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_int64, C.c_int64, C.c_int]
        _LIB.oracle_num_threads.restype = C.c_int
    return _LIB


# --------------------------------------------------------------------------- RNG state plumbing
# torch.get_rng_state() layout (CPUGeneratorImplStateLegacy): u64 seed | i32 left | i32 seeded |
# u64 next | u64 state[624] | normal cache ...   (torch/include/ATen/CPUGeneratorImpl.h)
_OFF_LEFT, _OFF_NEXT, _OFF_STATE = 8, 16, 24


def mt_from_torch(gen: Optional[torch.Generator] = None) -> MTState:
    raw = (torch.get_rng_state() if gen is None else gen.get_state()).numpy()
    mt = MTState()
    st = np.frombuffer(raw[_OFF_STATE:_OFF_STATE + 624 * 8].tobytes(), dtype=np.uint64).astype(np.uint32)
    C.memmove(mt.state, st.ctypes.data, 624 * 4)
    mt.left = int(np.frombuffer(raw[_OFF_LEFT:_OFF_LEFT + 4].tobytes(), dtype=np.int32)[0])
    mt.next = int(np.frombuffer(raw[_OFF_NEXT:_OFF_NEXT + 8].tobytes(), dtype=np.uint64)[0])
    return mt


def mt_to_torch(mt: MTState, gen: Optional[torch.Generator] = None) -> None:
    raw = (torch.get_rng_state() if gen is None else gen.get_state()).clone()
    arr = raw.numpy()
    st = np.ctypeslib.as_array(mt.state).astype(np.uint64)
    arr[_OFF_STATE:_OFF_STATE + 624 * 8] = np.frombuffer(st.tobytes(), dtype=np.uint8)
    arr[_OFF_LEFT:_OFF_LEFT + 4] = np.frombuffer(np.int32(mt.left).tobytes(), dtype=np.uint8)
    arr[_OFF_NEXT:_OFF_NEXT + 8] = np.frombuffer(np.uint64(mt.next).tobytes(), dtype=np.uint8)
    if gen is None:
        torch.set_rng_state(raw)
    else:
        gen.set_state(raw)


def mt_seed(seed: int) -> MTState:
    mt = MTState()
    lib().oracle_mt_seed(C.byref(mt), seed)
    return mt


def randint_words(mt: MTState, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint64)
    lib().oracle_randint_words(C.byref(mt), out.ctypes.data, n)
    return out


def _i64(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.int64, copy=False))


def _take(ptr, n: int, dtype: torch.dtype) -> torch.Tensor:
    if n == 0:
        return torch.empty(0, dtype=dtype)
    return torch.from_numpy(np.ctypeslib.as_array(ptr, shape=(n,)).copy()).to(dtype)


# --------------------------------------------------------------------------- sampler
def neighbor_sample(rowptr, col, seed, num_neighbors: List[int], node_time=None, edge_time=None,
                    seed_time=None, edge_weight=None, csc: bool = False, replace: bool = False,
                    directed: bool = True, disjoint: bool = False, temporal_strategy: str = 'uniform',
                    return_edge_id: bool = True, mt: Optional[MTState] = None):
    """Oracle for pyg_lib.sampler.neighbor_sample (uniform and temporal; weighted not restated)."""
    assert edge_weight is None
    if node_time is not None or edge_time is not None:
        if not disjoint:
            raise RuntimeError('Temporal sampling needs to create disjoint subgraphs')
        if edge_time is not None and seed_time is None:
            raise RuntimeError('Seed time needs to be specified')
    own_mt = mt is None
    if own_mt:
        mt = mt_from_torch()
    dt = seed.dtype
    rp, cl, sd = _i64(rowptr), _i64(col), _i64(seed)
    nn = np.asarray(num_neighbors, dtype=np.int64)
    res = HomoResult()
    nt = _i64(node_time) if node_time is not None else None
    et = _i64(edge_time) if edge_time is not None else None
    stt = _i64(seed_time) if seed_time is not None else None
    lib().oracle_neighbor_sample_temporal(rp.ctypes.data, rp.size - 1, cl.ctypes.data, sd.ctypes.data, sd.size,
                                          nn.ctypes.data, len(num_neighbors), int(replace), int(disjoint),
                                          C.byref(mt), C.byref(res), nt.ctypes.data if nt is not None else None,
                                          et.ctypes.data if et is not None else None,
                                          stt.ctypes.data if stt is not None else None, int(temporal_strategy == 'last'))
    L = len(num_neighbors)
    row = _take(res.row, res.n_edges, dt)
    colv = _take(res.col, res.n_edges, dt)
    eid = _take(res.edge_id, res.n_edges, dt)
    node = _take(res.node_id, res.n_nodes, dt)
    if disjoint:
        batch = _take(res.node_batch, res.n_nodes, dt)
        node = torch.stack([batch, node], dim=1)
    nph = [int(res.nodes_per_hop[i]) for i in range(L + 1)]
    eph = [int(res.edges_per_hop[i]) for i in range(L)]
    lib().oracle_homo_free(C.byref(res))
    if own_mt:
        mt_to_torch(mt)
    if not directed:
        raise RuntimeError('Undirected subgraphs not yet supported')
    if csc:
        row, colv = colv, row
    return row, colv, node, (eid if return_edge_id else None), nph, eph


def hetero_neighbor_sample(node_types: List[str], edge_types: List[Tuple[str, str, str]],
                           rowptr_dict: Dict[str, torch.Tensor], col_dict: Dict[str, torch.Tensor],
                           seed_dict: Dict[str, torch.Tensor], num_neighbors_dict: Dict[str, List[int]],
                           csc: bool = False, replace: bool = False, directed: bool = True,
                           disjoint: bool = False, return_edge_id: bool = True,
                           mt: Optional[MTState] = None, node_time_dict=None, edge_time_dict=None, seed_time_dict=None,
                           temporal_strategy: str = 'uniform'):
    """Oracle for torch.ops.pyg.hetero_neighbor_sample with ONE ATen thread; dict keys are
    'src__rel__dst' strings exactly like the operator (neighbor.cpp:137-147)."""
    own_mt = mt is None
    if own_mt:
        mt = mt_from_torch()
    rel = ['__'.join(k) for k in edge_types]
    # node type order: seed_dict order first (batch ids in disjoint mode follow it), then the rest
    types = list(seed_dict.keys()) + [t for t in node_types if t not in seed_dict]
    tix = {t: i for i, t in enumerate(types)}
    T, R = len(types), len(edge_types)
    L = max(len(num_neighbors_dict[r]) for r in rel) if R else 0
    dt = next(iter(seed_dict.values())).dtype
    src = np.array([tix[k[0] if not csc else k[2]] for k in edge_types], dtype=np.int64)
    dst = np.array([tix[k[2] if not csc else k[0]] for k in edge_types], dtype=np.int64)
    rps = [_i64(rowptr_dict[r]) for r in rel]
    cls = [_i64(col_dict[r]) for r in rel]
    sds = [_i64(seed_dict[t]) if t in seed_dict else np.empty(0, dtype=np.int64) for t in types]
    PP = C.c_void_p * max(R, 1)
    PT = C.c_void_p * max(T, 1)
    rp_arr = PP(*[a.ctypes.data for a in rps])
    cl_arr = PP(*[a.ctypes.data for a in cls])
    sd_arr = PT(*[a.ctypes.data for a in sds])
    nsd = np.array([a.size for a in sds], dtype=np.int64)
    nn = np.array([list(num_neighbors_dict[r]) for r in rel], dtype=np.int64).reshape(R, L)
    res = HeteroResult()
    keep = []

    def ptr_array(d, keys, n):
        if d is None:
            return None
        arrs = [(_i64(d[k]) if k in d else None) for k in keys]
        keep.append(arrs)
        return (C.c_void_p * max(n, 1))(*[(a.ctypes.data if a is not None else None) for a in arrs])
    nt_arr = ptr_array(node_time_dict, types, T)
    et_arr = ptr_array(edge_time_dict, rel, R)
    st_arr = ptr_array(seed_time_dict, types, T)
    lib().oracle_hetero_neighbor_sample_temporal(T, R, src.ctypes.data, dst.ctypes.data, rp_arr, cl_arr, sd_arr,
                                                 nsd.ctypes.data, nn.ctypes.data, L, int(replace), int(disjoint),
                                                 C.byref(mt), C.byref(res), nt_arr, et_arr, st_arr,
                                                 int(temporal_strategy == 'last'))
    row_d, col_d, eid_d, node_d, nph_d, eph_d = {}, {}, {}, {}, {}, {}
    for t in node_types:
        i = tix[t]
        n = int(res.n_nodes[i])
        node = _take(res.node_id[i], n, dt)
        if disjoint:
            node = torch.stack([_take(res.node_batch[i], n, dt), node], dim=1) if n else node.view(0, 2)
        node_d[t] = node
        nph_d[t] = [int(res.nodes_per_hop[i * (L + 1) + j]) for j in range(L + 1)]
    for k, r in enumerate(rel):
        n = int(res.n_edges[k])
        a, b = _take(res.row[k], n, dt), _take(res.col[k], n, dt)
        row_d[r], col_d[r] = (a, b) if not csc else (b, a)
        eid_d[r] = _take(res.edge_id[k], n, dt)
        eph_d[r] = [int(res.edges_per_hop[k * L + j]) for j in range(L)]
    lib().oracle_hetero_free(C.byref(res))
    if own_mt:
        mt_to_torch(mt)
    if not directed:
        raise RuntimeError('Undirected heterogeneous graphs not yet supported')
    return row_d, col_d, node_d, (eid_d if return_edge_id else None), nph_d, eph_d


# --------------------------------------------------------------------------- matmul
_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _raw(t: torch.Tensor) -> torch.Tensor:
    t = t.detach().cpu().contiguous()
    return t if t.dtype == torch.float32 else t.view(torch.int16)


def segment_matmul(inputs: torch.Tensor, ptr: torch.Tensor, other: torch.Tensor,
                   bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Oracle for pyg_lib.ops.segment_matmul (pyg_lib/ops/__init__.py:137-172)."""
    x, w = _raw(inputs), _raw(other)
    p = np.ascontiguousarray(ptr.detach().cpu().numpy().astype(np.int64))
    N, K = inputs.shape
    B, _, M = other.shape
    out = torch.zeros(N, M, dtype=inputs.dtype)
    o = out if out.dtype == torch.float32 else out.view(torch.int16)
    lib().oracle_segment_matmul(x.data_ptr(), p.ctypes.data, w.data_ptr(), o.data_ptr(), K, M, B,
                                _DT[inputs.dtype])
    if bias is not None:
        for i in range(B):
            out[int(p[i]):int(p[i + 1])] += bias[i].cpu()
    return out


def grouped_matmul(inputs: List[torch.Tensor], others: List[torch.Tensor],
                   biases: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """Oracle for pyg_lib.ops.grouped_matmul (pyg_lib/ops/__init__.py:99-134)."""
    outs = []
    for i, (a, b) in enumerate(zip(inputs, others)):
        ptr = torch.tensor([0, a.size(0)])
        o = segment_matmul(a.contiguous(), ptr, b.contiguous().unsqueeze(0))
        if biases is not None:
            o = o + biases[i].cpu()
        outs.append(o)
    return outs


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def subgraph(rowptr: torch.Tensor, col: torch.Tensor, nodes: torch.Tensor,
             return_edge_id: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Oracle for pyg_lib.sampler.subgraph (numpy restatement of pyg_lib/csrc/sampler/cpu/subgraph_kernel.cpp:13-89).

    `nodes` are numbered by `Mapper::fill` (mapper.h:49-53,29-46): dense ids in first-occurrence order, a
    repeated node keeps the id of its first occurrence but still gets its own output row.  Row i lists, in CSR
    order, the neighbours of nodes[i] that are in the set (`:41-52,66-82`), mapped to those ids; `edge_id` is the
    position of the kept edge in `col`."""
    dt = nodes.dtype
    rp, cl, nd = _i64(rowptr), _i64(col), _i64(nodes)
    ids: Dict[int, int] = {}
    for v in nd.tolist():
        if v not in ids:
            ids[v] = len(ids)
    lut = np.full(int(rp.shape[0]) - 1, -1, dtype=np.int64)
    for v, i in ids.items():
        lut[v] = i
    out_rowptr = np.zeros(nd.shape[0] + 1, dtype=np.int64)
    cols, eids = [], []
    for i, v in enumerate(nd.tolist()):
        j = np.arange(rp[v], rp[v + 1], dtype=np.int64)
        w = lut[cl[j]]
        keep = w >= 0
        cols.append(w[keep]); eids.append(j[keep])
        out_rowptr[i + 1] = out_rowptr[i] + int(keep.sum())
    out_col = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    out_eid = np.concatenate(eids) if eids else np.zeros(0, dtype=np.int64)
    return (torch.from_numpy(out_rowptr).to(dt), torch.from_numpy(out_col).to(dt),
            torch.from_numpy(out_eid).to(dt) if return_edge_id else None)


def dist_neighbor_sample(rowptr, col, seed, num_neighbors: int, node_time=None, edge_time=None, seed_time=None,
                         edge_weight=None, csc: bool = False, replace: bool = False, directed: bool = True,
                         disjoint: bool = False, temporal_strategy: str = 'uniform', mt: Optional[MTState] = None):
    """Oracle for pyg::dist_neighbor_sample (dist_neighbor_sample_kernel, neighbor_kernel.cpp:957-978).

    The distributed variant is the ordinary one-hop `sample<>` with `distributed=true`: same frontier order, same
    draws, but `add` pushes every sampled neighbour unmapped (`:296-303`) and the loop records
    cumsum_neighbors_per_node (`:386-388,446-492`).  So it is restated on top of the one-hop oracle: edges come out in
    the same order, `col[edge_id[p]]` is the global id of edge p's neighbour, `row[p]` the seed it belongs to."""
    r = neighbor_sample(rowptr, col, seed, [int(num_neighbors)], node_time, edge_time, seed_time, edge_weight, False, replace,
                        directed, disjoint, temporal_strategy, True, mt)
    row, eid = r[0], r[3]
    S = int(seed.numel())
    glob = col[eid.to(torch.int64)].to(seed.dtype)      # global id of every sampled neighbour (edge id = position in col)
    if disjoint:                                        # (batch, node): a neighbour inherits the batch = index of its seed
        nodes = torch.stack([torch.cat([torch.arange(S, dtype=seed.dtype), row.to(seed.dtype)]), torch.cat([seed, glob])], 1)
    else:
        nodes = torch.cat([seed, glob])
    per_seed = torch.bincount(row.to(torch.int64), minlength=S)
    cumsum = [S] + (S + torch.cumsum(per_seed, 0)).tolist()
    return nodes, eid, cumsum


def relabel_neighborhood(seed: torch.Tensor, sampled_nodes_with_duplicates: torch.Tensor, num_sampled_neighbors_per_node: List[int],
                         num_nodes: int, batch: Optional[torch.Tensor] = None, csc: bool = False,
                         disjoint: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Oracle for pyg::relabel_neighborhood (relabel<disjoint>, pyg_lib/csrc/sampler/cpu/dist_relabel_kernel.cpp:30-95):
    the mapper is filled with the seeds (`:67-74`; disjoint: key (i, seed_i)), then every sampled node is inserted in
    order (`:76-92`); row = index of the source node the neighbour was counted for, col = the mapper's id."""
    ids: Dict = {}
    for i, v in enumerate(seed.tolist()):
        k = (i, v) if disjoint else v
        if k not in ids:
            ids[k] = len(ids)
    nodes = sampled_nodes_with_duplicates.tolist()
    bt = batch.tolist() if disjoint else None
    rows, cols = [], []
    j = 0
    for i, c in enumerate(num_sampled_neighbors_per_node):
        for _ in range(int(c)):
            k = (bt[j], nodes[j]) if disjoint else nodes[j]
            if k not in ids:
                ids[k] = len(ids)
            rows.append(i); cols.append(ids[k])
            j += 1
    row = torch.tensor(rows, dtype=seed.dtype); col = torch.tensor(cols, dtype=seed.dtype)
    return (col, row) if csc else (row, col)


def hetero_relabel_neighborhood(node_types: List[str], edge_types: List[Tuple[str, str, str]], seed_dict: Dict[str, torch.Tensor],
                                sampled_dict: Dict[str, torch.Tensor], counts_dict: Dict[str, List[List[int]]],
                                num_nodes_dict: Dict[str, int], batch_dict: Optional[Dict[str, torch.Tensor]] = None,
                                csc: bool = False, disjoint: bool = False):
    """Oracle for pyg::hetero_relabel_neighborhood (hetero relabel<disjoint>, dist_relabel_kernel.cpp:97-273), 1 thread.

    One mapper per node type, filled with that type's seeds (disjoint: keys (batch, node), the batch counter running
    over all seeds in seed_dict order, `:180-193`).  Layer by layer, edge type by edge type, source node by source
    node, the next `count` entries of the DST type's sampled list are inserted (`:206-235`); row = the source node's
    local index, which per edge type starts at 0 and continues after the largest index any edge type with the same
    source node type used in the previous layer (`:240-258`)."""
    dt = next(iter(seed_dict.values())).dtype
    rel = lambda k: '__'.join(k)  # noqa: E731
    ids = {t: {} for t in node_types}
    cursor = {t: 0 for t in node_types}
    b = 0
    for t, sd in seed_dict.items():
        for v in sd.tolist():
            key = (b, v) if disjoint else v
            if disjoint:
                b += 1
            if key not in ids[t]:
                ids[t][key] = len(ids[t])
    sampled = {t: sampled_dict[t].tolist() for t in node_types}
    batches = {t: batch_dict[t].tolist() for t in node_types} if disjoint else None
    rows = {k: [] for k in edge_types}
    cols = {k: [] for k in edge_types}
    L = len(counts_dict[rel(edge_types[0])])
    src_slice = {k: (0, len(counts_dict[rel(k)][0])) for k in edge_types}
    src_off = {t: 0 for t in node_types}
    for ell in range(L):
        for k in edge_types:
            dst = k[0] if csc else k[2]
            begin_i, end_i = src_slice[k]
            for i in range(begin_i, end_i):
                c = int(counts_dict[rel(k)][ell][i - begin_i])
                for j in range(cursor[dst], cursor[dst] + c):
                    key = (batches[dst][j], sampled[dst][j]) if disjoint else sampled[dst][j]
                    if key not in ids[dst]:
                        ids[dst][key] = len(ids[dst])
                    rows[k].append(i); cols[k].append(ids[dst][key])
                cursor[dst] += c
        if ell < L - 1:
            for k in edge_types:
                src = k[2] if csc else k[0]
                src_off[src] = max(src_off[src], src_slice[k][1])
            for k in edge_types:
                src = k[2] if csc else k[0]
                src_slice[k] = (src_off[src], src_off[src] + len(counts_dict[rel(k)][ell + 1]))
    out_row, out_col = {}, {}
    for k in edge_types:
        r, c = torch.tensor(rows[k], dtype=dt), torch.tensor(cols[k], dtype=dt)
        out_row[rel(k)], out_col[rel(k)] = (c, r) if csc else (r, c)
    return out_row, out_col


def merge_sampler_outputs(node_ids: List[torch.Tensor], edge_ids: List[torch.Tensor], cumsum_neighbors_per_node: List[List[int]],
                          partition_ids: List[int], partition_orders: List[int], num_partitions: int, num_neighbors: int,
                          batch: Optional[torch.Tensor] = None, disjoint: bool = False):
    """Oracle for pyg::merge_sampler_outputs (merge_outputs<disjoint>, cpu/dist_merge_outputs_kernel.cpp:15-137): seed j's
    neighbours are node_ids[p][c[o]:c[o+1]] and edge_ids[p][c[o]-c[0]:c[o+1]-c[0]] with p = partition_ids[j],
    o = partition_orders[j], c = cumsum_neighbors_per_node[p] (`:79-92`); the outputs are their concatenation in seed
    order (the reference pads every seed to `offset` entries with -1 and removes the -1 again, `:60-64,107-122`)."""
    nodes, edges, batches, counts = [], [], [], []
    for j, (p, o) in enumerate(zip(partition_ids, partition_orders)):
        c = cumsum_neighbors_per_node[p]
        b, e = c[o], c[o + 1]
        nodes.append(node_ids[p][b:e]); edges.append(edge_ids[p][b - c[0]:e - c[0]])
        if disjoint:
            batches.append(batch[j].repeat(e - b))
        counts.append(e - b)
    dt = node_ids[0].dtype
    cat = lambda xs: torch.cat(xs) if xs else torch.zeros(0, dtype=dt)  # noqa: E731
    return cat(nodes), cat(edges), (cat(batches) if disjoint else None), counts
