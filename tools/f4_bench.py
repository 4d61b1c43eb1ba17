"""CUDA-event timings and algorithmic-bytes rooflines of the SURVEY 8(f)(4) ops at ogbn-products scale:
pyg::subgraph, pyg::dist_neighbor_sample, pyg::relabel_neighborhood, pyg::merge_sampler_outputs.

    python tools/f4_bench.py > gpurun_out/f4_bench.json        (one GPU; prints one JSON object)

Algorithmic bytes (int64): what the op must read and write once —
  subgraph               n rowptr pairs (16 B) + every col entry of the selected rows (8 B) + 16 B per kept edge (col, edge id) + 8 B per node
  dist_neighbor_sample   16 B rowptr pair per seed + 8 B col gather + 16 B (node, edge id) written per sampled neighbour
  relabel_neighborhood   8 B read + 16 B (row, col) written per sampled neighbour, 8 B per seed
  merge_sampler_outputs  16 B read + 16 B written per neighbour (node id, edge id)
The fraction of the measured HBM peak is reported next to the time; these ops are latency-chains of small launches at
this size (a few MB), like C2 — the time per call is the number that matters."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import pyg_lib_b200 as P  # noqa: E402
from graphs import lognormal_csr  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
dev = 'cuda:0'
N, E = 2_449_029, 123_718_280
peak = 6577.0
try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:  # noqa
    pass
rowptr, col = lognormal_csr(N, E, seed=1, device=dev)
perm = torch.randperm(N, generator=torch.Generator().manual_seed(2)).to(dev)


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, out   # us


res = {'graph': 'ogbn-products-shaped CSR (2,449,029 nodes / 123,718,280 edges, int64)', 'hbm_peak_gbs': peak, 'ops': {}}


def add(name, us, nbytes, **kw):
    res['ops'][name] = dict(us_per_call=us, algorithmic_bytes=nbytes, gbs=nbytes / us / 1e3, frac_of_measured_hbm_peak=nbytes / us / 1e3 / peak, **kw)


# ---- subgraph on the node set of a sampled 2-hop neighbourhood (what PyG's subgraph loaders pass) and on a big random set
torch.manual_seed(1)
for label, nodes in (('subgraph_2hop_nodeset', P.sampler.neighbor_sample(rowptr, col, perm[:1024], [15, 10])[2]),
                     ('subgraph_200k_random_nodes', perm[:200_000].contiguous())):
    us, out = timed(lambda: torch.ops.pyg.subgraph(rowptr, col, nodes, True))
    deg = (rowptr[nodes + 1] - rowptr[nodes]).sum().item()
    kept = out[1].numel()
    add(label, us, 16 * nodes.numel() + 8 * deg + 16 * kept + 8 * nodes.numel(), nodes=nodes.numel(), scanned_edges=deg, kept_edges=kept)

# ---- one distributed hop: dist_neighbor_sample -> relabel_neighborhood (== one hop of neighbor_sample)
for S, k in ((1024, 15), (16384, 10), (65536, 15)):
    seed = perm[:S].contiguous()
    torch.manual_seed(3)
    us, (node, eid, cum) = timed(lambda: torch.ops.pyg.dist_neighbor_sample(rowptr, col, seed, k, None, None, None, None, False, False, True, False, 'uniform'))
    n = eid.numel()
    add(f'dist_neighbor_sample_{S}x{k}', us, 16 * S + 8 * n + 16 * n, seeds=S, sampled=n,
        note='includes the D2H of the cumulative counts (seeds + 1 values): the reference API returns them as a host list')
    counts = (torch.tensor(cum[1:]) - torch.tensor(cum[:-1])).tolist()
    sampled = node[S:].contiguous()
    us2, _ = timed(lambda: torch.ops.pyg.relabel_neighborhood(seed, sampled, counts, N, None, False, False))
    add(f'relabel_neighborhood_{S}x{k}', us2, 8 * S + 8 * n + 16 * n, seeds=S, sampled=n,
        note='the per-node counts are a host int[] in the reference API: their upload is part of the call')

# ---- merge_sampler_outputs: 4 partitions of a 16384-seed hop
S, k, parts = 16384, 10, 4
seed = perm[:S].contiguous()
part = torch.randint(0, parts, (S,), generator=torch.Generator().manual_seed(4))
node_ids, edge_ids, cums, orders = [], [], [], [0] * S
for p_ in range(parts):
    idx = (part == p_).nonzero().flatten()
    for o, j in enumerate(idx.tolist()):
        orders[j] = o
    torch.manual_seed(5 + p_)
    n_, e_, c_ = torch.ops.pyg.dist_neighbor_sample(rowptr, col, seed[idx.to(dev)], k, None, None, None, None, False, False, True, False, 'uniform')
    node_ids.append(n_); edge_ids.append(e_); cums.append(c_)
pids = part.tolist()
us, out = timed(lambda: torch.ops.pyg.merge_sampler_outputs(node_ids, edge_ids, cums, pids, orders, parts, k, None, False))
n = out[1].numel()
add(f'merge_sampler_outputs_{parts}x{S // parts}x{k}', us, 32 * n, seeds=S, merged=n,
    note='cumulative counts, partition ids and orders are host lists in the reference API: the plan is host arithmetic')
print(json.dumps(res))
