"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's *biased* (edge_weight) neighbour sampling.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does.

The reference's biased path is "call these ATen functions on the CPU generator" (neighbor_kernel.cpp:245-285): per
frontier node either the whole neighbourhood (count < 0 or, without replacement, count >= degree; :259-265),
at::multinomial(weight, count, replace=True) (:267-271) or the weighted-reservoir keys rand.log() / weight with
rand = at::empty_like(weight).uniform_() and their top-`count` (:272-278), then add() of row_start + index in that order
(:279-283).  The restatement makes the same ATen calls in the same order on torch's default CPU generator, so it is
bit-faithful by construction (and is pinned to the reference built by oracle/build_ref.sh in tests/test_weighted_oracle.py);
the loops around them follow neighbor_kernel.cpp:383-470 (homogeneous) and :591-823 (heterogeneous, single-threaded order).
The RandintEngine the reference constructs (and never uses on this path) prefetches 128 int64 at construction
(rand_engine.h:28,80-85): the same torch.randint call is made here so the generator ends in the same state.
"""
from typing import Dict, List, Optional, Tuple

import torch

I64_MIN, I64_MAX = -(1 << 63), (1 << 63) - 1


def _engine_prefetch():
    torch.randint(I64_MIN, I64_MAX, (128,), dtype=torch.int64)   # PrefetchedRandint::prefetch (rand_engine.h:80-85)


def _biased_indices(weight: torch.Tensor, count: int, replace: bool) -> Optional[List[int]]:
    """neighbor_kernel.cpp:257-278: None = the full neighbourhood, else the chosen in-row indices in add() order."""
    population = weight.numel()
    if count < 0 or (not replace and count >= population):
        return None
    if replace:
        return torch.multinomial(weight, count, True).tolist()
    rand = torch.empty_like(weight).uniform_()
    key = rand.log() / weight
    return key.topk(count)[1].tolist()


class _Mapper:   # mapper.h:29-46 (the hash-map variant; ids in first-insertion order)
    def __init__(self):
        self.ids = {}

    def insert(self, node):
        n = len(self.ids)
        got = self.ids.setdefault(node, n)
        return got, got == n


def neighbor_sample(rowptr: torch.Tensor, col: torch.Tensor, seed: torch.Tensor, num_neighbors: List[int],
                    edge_weight: torch.Tensor, replace: bool = False, csc: bool = False, disjoint: bool = False):
    """(row, col, node_id, edge_id, num_sampled_nodes_per_hop, num_sampled_edges_per_hop) like pyg::neighbor_sample with
    edge_weight given (neighbor_kernel.cpp:383-503); the default CPU generator is advanced exactly as the reference does."""
    rp, cl, sd = rowptr.tolist(), col.tolist(), seed.tolist()
    _engine_prefetch()
    mapper = _Mapper()
    nodes = []
    for i, v in enumerate(sd):   # :409-416 (duplicate seeds are listed twice but mapped once, like the reference)
        node = (i, v) if disjoint else v
        nodes.append(node)
        mapper.insert(node)
    rows, cols, eids, nph, eph = [], [], [], [len(sd)], []
    begin, end = 0, len(sd)
    for count in num_neighbors:
        n_e = 0
        for i in range(begin, end):   # :436-448
            src = nodes[i]
            v = src[1] if disjoint else src
            rs, re = rp[v], rp[v + 1]
            if re - rs == 0 or count == 0:
                continue
            idx = _biased_indices(edge_weight[rs:re], count, replace)
            edges = range(rs, re) if idx is None else [rs + j for j in idx]
            for e in edges:   # add(): :287-317
                dst = (src[0], cl[e]) if disjoint else cl[e]
                local, new = mapper.insert(dst)
                if new:
                    nodes.append(dst)
                rows.append(i); cols.append(local); eids.append(e); n_e += 1
        begin, end = end, len(nodes)
        nph.append(end - begin); eph.append(n_e)
    dt = seed.dtype
    node_t = torch.tensor(nodes, dtype=dt).reshape(-1, 2) if disjoint else torch.tensor(nodes, dtype=dt)
    row_t, col_t = torch.tensor(rows, dtype=dt), torch.tensor(cols, dtype=dt)
    if csc:
        row_t, col_t = col_t, row_t
    return row_t, col_t, node_t, torch.tensor(eids, dtype=dt), nph, eph


def hetero_neighbor_sample(node_types: List[str], edge_types: List[Tuple[str, str, str]], rowptr_dict: Dict[str, torch.Tensor],
                           col_dict: Dict[str, torch.Tensor], seed_dict: Dict[str, torch.Tensor],
                           num_neighbors_dict: Dict[str, List[int]], edge_weight_dict: Dict[str, torch.Tensor],
                           replace: bool = False, csc: bool = False):
    """Heterogeneous, every relation weighted, non-disjoint, one thread (neighbor_kernel.cpp:591-823): per hop the relations
    in edge_types order sample from the frontier slice their source type had at the START of the hop; new nodes join their
    type's list at once (single-threaded: dst_sampled_nodes == sampled_nodes_dict, :722-724)."""
    _engine_prefetch()
    L = max(len(v) for v in num_neighbors_dict.values()) if num_neighbors_dict else 0
    nodes = {t: [] for t in node_types}
    mapper = {t: _Mapper() for t in node_types}
    nph = {t: [] for t in node_types}
    for t, s in seed_dict.items():   # :669-704
        for v in s.tolist():
            nodes[t].append(v); mapper[t].insert(v)
    sl = {t: (0, len(nodes[t])) for t in node_types}
    for t in node_types:
        nph[t].append(len(nodes[t]))
    rels = ['__'.join(k) for k in edge_types]
    rows = {r: [] for r in rels}; cols = {r: [] for r in rels}; eids = {r: [] for r in rels}; eph = {r: [] for r in rels}
    for ell in range(L):
        for k, r in zip(edge_types, rels):
            src, dst = (k[0], k[2]) if not csc else (k[2], k[0])
            count = num_neighbors_dict[r][ell]
            rp, cl, w = rowptr_dict[r].tolist(), col_dict[r].tolist(), edge_weight_dict[r]
            b, e_ = sl[src]
            n_e = 0
            for i in range(b, e_):
                v = nodes[src][i]
                rs, re = rp[v], rp[v + 1]
                if re - rs == 0 or count == 0:
                    continue
                idx = _biased_indices(w[rs:re], count, replace)
                edges = range(rs, re) if idx is None else [rs + j for j in idx]
                for e in edges:
                    local, new = mapper[dst].insert(cl[e])
                    if new:
                        nodes[dst].append(cl[e])
                    rows[r].append(i); cols[r].append(local); eids[r].append(e); n_e += 1
            eph[r].append(n_e)
        for t in node_types:   # :807-812
            sl[t] = (sl[t][1], len(nodes[t]))
            nph[t].append(sl[t][1] - sl[t][0])
    dt = next(iter(seed_dict.values())).dtype
    T = lambda x: torch.tensor(x, dtype=dt)   # noqa: E731
    row_d = {r: T(rows[r]) for r in rels}; col_d = {r: T(cols[r]) for r in rels}
    if csc:
        row_d, col_d = col_d, row_d
    return row_d, col_d, {t: T(nodes[t]) for t in node_types}, {r: T(eids[r]) for r in rels}, nph, eph
