"""GPU parity: pyg::merge_sampler_outputs (host plan + segment gathers) vs fixtures made by the reference
(tests/golden/merge_outputs.npz).  Bit-exact.  (Collected last: the whole dist pipeline — sample per partition,
merge, relabel — is exercised at the end.)"""
import os.path as osp

import numpy as np
import pytest
import torch

from graphs import MERGE_CASES, build_merge, lognormal_csr
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
HERE = osp.dirname(osp.abspath(__file__))


@pytest.fixture(scope='module')
def lib():
    import pyg_lib_b200
    return pyg_lib_b200


@pytest.mark.parametrize('name', list(MERGE_CASES))
@pytest.mark.parametrize('dtype', [torch.int64, torch.int32])
def test_merge_sampler_outputs_golden(lib, name, dtype):
    G = np.load(osp.join(HERE, 'golden', 'merge_outputs.npz'))
    node_ids, edge_ids, cums, pids, pords, P, k, batch, disjoint = build_merge(MERGE_CASES[name], O.dist_neighbor_sample)
    out = torch.ops.pyg.merge_sampler_outputs([t.to(DEV, dtype) for t in node_ids], [t.to(DEV, dtype) for t in edge_ids], cums, pids,
                                              pords, P, k, None if batch is None else batch.to(DEV, dtype), disjoint)
    assert out[0].dtype == dtype and out[1].dtype == dtype
    assert np.array_equal(out[0].cpu().numpy(), G[f'{name}/node']) and np.array_equal(out[1].cpu().numpy(), G[f'{name}/edge'])
    assert list(out[3]) == G[f'{name}/counts'].tolist()
    if disjoint:
        assert np.array_equal(out[2].cpu().numpy(), G[f'{name}/batch'])
    else:
        assert out[2] is None


def test_partitioned_sample_merge_relabel_equals_neighbor_sample(lib):
    """The distributed pipeline on one GPU: seeds dealt to 4 partitions, one dist_neighbor_sample per partition (each
    from a fresh generator state, as separate workers would), merge, relabel -> the same local graph as relabelling
    one dist_neighbor_sample over all seeds drawn with the same per-seed neighbourhoods is not defined (the draws
    differ per call order), so the check is structural: merged rows follow the seed order, merged neighbours are
    neighbours of their seed, and relabel reproduces first-occurrence ids of [seeds | merged]."""
    rowptr, col = lognormal_csr(100_000, 3_000_000, seed=6)
    d_rowptr, d_col = rowptr.to(DEV), col.to(DEV)
    g = torch.Generator().manual_seed(12)
    seed = torch.randperm(100_000, generator=g)[:3000]
    part = torch.randint(0, 4, (3000,), generator=g)
    node_ids, edge_ids, cums, orders = [], [], [], [0] * 3000
    for p in range(4):
        idx = (part == p).nonzero().flatten()
        for o, j in enumerate(idx.tolist()):
            orders[j] = o
        torch.manual_seed(100 + p)
        n, e, c = torch.ops.pyg.dist_neighbor_sample(d_rowptr, d_col, seed[idx].to(DEV), 6, None, None, None, None, False, False, True,
                                                     False, 'uniform')
        node_ids.append(n); edge_ids.append(e); cums.append(list(c))
    nodes, edges, _, counts = torch.ops.pyg.merge_sampler_outputs(node_ids, edge_ids, cums, part.tolist(), orders, 4, 6, None, False)
    assert sum(counts) == nodes.numel() == edges.numel()
    assert torch.equal(d_col[edges], nodes)                                    # edge ids point at the merged neighbours
    src = torch.repeat_interleave(seed.to(DEV), torch.tensor(counts, device=DEV))
    assert bool(((edges >= d_rowptr[src]) & (edges < d_rowptr[src + 1])).all())   # ... inside their seed's row
    row, colv = torch.ops.pyg.relabel_neighborhood(seed.to(DEV), nodes, counts, 100_000, None, False, False)
    exp = O.relabel_neighborhood(seed, nodes.cpu(), counts, 100_000)
    assert torch.equal(row.cpu(), exp[0]) and torch.equal(colv.cpu(), exp[1])
