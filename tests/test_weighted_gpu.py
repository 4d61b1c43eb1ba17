"""GPU parity of BIASED (edge_weight) sampling: torch.ops.pyg.neighbor_sample / hetero_neighbor_sample with edge weights
-> pygb200_sampler_run_weighted, against fixtures produced by the reference itself (tests/golden/weighted_outputs.npz) and
against the oracle (oracle/weighted.py) on bigger inputs.  Bit-exact: every index tensor, every count, and the state the call
leaves the CPU generator in."""
import os

import numpy as np
import pytest
import torch

from graphs import (HETERO_WEIGHTED_CASES, WEIGHTED_CASES, build_hetero_weighted, build_weighted, build_weights, random_csr)
from oracle import oracle as O
from oracle import weighted as W

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'weighted_outputs.npz'))


@pytest.fixture(scope='module')
def lib():
    import pyg_lib_b200  # noqa: F401
    return pyg_lib_b200


@pytest.fixture(scope='module')
def live_oracle():
    """Tests that run oracle/weighted.py on THIS host compare the GPU with this host's torch.log (MKL).  The device reproduces the
    MKL the fixtures were made with (csrc/mkl_logf_table.inc); should a host's MKL take another code path, those tests say so
    instead of failing — the fixture-based tests above remain the parity statement."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tools'))
    from make_logf_table import read_inc
    t = read_inc()
    k = np.unique(np.concatenate([(t >> 1).astype(np.int64), np.random.RandomState(0).randint(1, 1 << 24, 200000)]))
    u = k.astype(np.float32) * np.float32(2.0 ** -24)
    bits = np.log(u.astype(np.float64)).astype(np.float32).view(np.int32).copy()
    pos = np.searchsorted(k, (t >> 1).astype(np.int64))
    bits[pos] += np.where(t & 1, 1, -1).astype(np.int32)
    if not np.array_equal(torch.log(torch.from_numpy(u)).numpy().view(np.int32), bits):
        pytest.skip("this host's torch.log (MKL) differs from the table the reference fixtures were made with")
    return True


def _rng():
    return torch.get_rng_state().numpy()[:24 + 624 * 8].copy()


def _cmp(out, exp):
    assert list(out[4]) == list(exp[4]) and list(out[5]) == list(exp[5])
    for k in (2, 0, 1, 3):
        assert torch.equal(out[k].cpu().long(), exp[k].long()), ('node', 'row', 'col', 'eid')[k - 2 if k == 2 else k]


@pytest.mark.parametrize('name', list(WEIGHTED_CASES))
@pytest.mark.parametrize('dtype', [torch.int64, torch.int32])
def test_weighted_golden(lib, name, dtype):
    case = WEIGHTED_CASES[name]
    rowptr, col, seed, w = build_weighted(case)
    torch.manual_seed(case['rng_seed'])
    out = lib.sampler.neighbor_sample(rowptr.to(DEV, dtype), col.to(DEV, dtype), seed.to(DEV, dtype), case['num_neighbors'],
                                      edge_weight=w.to(DEV), csc=case.get('csc', False), replace=case.get('replace', False),
                                      disjoint=case.get('disjoint', False))
    p = f'homo/{name}/'
    assert out[0].dtype == dtype and out[2].dtype == dtype
    assert out[4] == GOLD[p + 'nph'].tolist() and out[5] == GOLD[p + 'eph'].tolist()
    for i, k in ((2, 'node'), (0, 'row'), (1, 'col'), (3, 'eid')):
        assert np.array_equal(out[i].cpu().numpy(), GOLD[p + k]), k
    assert np.array_equal(_rng(), GOLD[p + 'rng_after'])


@pytest.mark.parametrize('name', list(HETERO_WEIGHTED_CASES))
def test_hetero_weighted_golden(lib, name):
    case = HETERO_WEIGHTED_CASES[name]
    nt, et, rp, cl, sd, nn, wd = build_hetero_weighted(case)
    dv = lambda d: {k: v.to(DEV) for k, v in d.items()}  # noqa: E731
    torch.manual_seed(case['rng_seed'])
    out = torch.ops.pyg.hetero_neighbor_sample(nt, et, dv(rp), dv(cl), dv(sd), nn, None, None, None, dv(wd), case.get('csc', False),
                                               case.get('replace', False), True, case.get('disjoint', False), 'uniform', True)
    p = f'hetero/{name}/'
    for i, key in enumerate(('row', 'col', 'node', 'eid')):
        for k, v in out[i].items():
            assert np.array_equal(v.cpu().numpy(), GOLD[p + f'{key}/{k}']), (key, k)
    for k, v in out[4].items():
        assert list(v) == GOLD[p + f'nph/{k}'].tolist()
    for k, v in out[5].items():
        assert list(v) == GOLD[p + f'eph/{k}'].tolist()
    assert np.array_equal(_rng(), GOLD[p + 'rng_after'])


def test_reference_known_answer(lib):
    """BiasedNeighborTest, test/csrc/sampler/test_neighbor.cpp:300-329: cycle graph, only the even edges carry weight."""
    n = 6
    rowptr = torch.arange(0, 2 * n + 1, 2)
    col = torch.stack([(torch.arange(n) - 1) % n, (torch.arange(n) + 1) % n], 1).view(-1)
    w = torch.tensor([1.0, 0.0] * n)
    out = lib.sampler.neighbor_sample(rowptr.to(DEV), col.to(DEV), torch.arange(0, 2).to(DEV), [1], edge_weight=w.to(DEV))
    assert out[0].tolist() == [0, 1] and out[1].tolist() == [2, 0] and out[2].tolist() == [0, 1, 5] and out[3].tolist() == [0, 2]


@pytest.mark.parametrize('replace', [False, True])
@pytest.mark.parametrize('weights', ['uniform', 'masked_pos', 'quantized_pos'])
def test_bigger_graph_against_the_oracle(lib, live_oracle, replace, weights):
    rowptr, col = random_csr(20000, 30, 5, big=[(11, 40000), (12, 3000)])
    if weights == 'quantized_pos':
        w = build_weights('quantized', rowptr, 3) + (torch.arange(col.numel()) % 7 == 0).float()
    else:
        w = build_weights(weights, rowptr, 3)
    if replace:   # at::multinomial wants a positive sum in every sampled row
        w[rowptr[:-1][rowptr[:-1] < rowptr[1:]]] += 0.125
    seed = torch.cat([torch.tensor([11, 12]), torch.randperm(20000, generator=torch.Generator().manual_seed(9))[:1500]])
    nn = [12, 6]
    torch.manual_seed(99)
    out = lib.sampler.neighbor_sample(rowptr.to(DEV), col.to(DEV), seed.to(DEV), nn, edge_weight=w.to(DEV), replace=replace)
    after = _rng()
    torch.manual_seed(99)
    exp = W.neighbor_sample(rowptr, col, seed, nn, w, replace=replace)
    _cmp(out, exp)
    assert np.array_equal(after, _rng())


def test_stream_continues_across_calls_and_kinds(lib, live_oracle):
    """Biased and uniform calls in a row without reseeding: every call starts where the previous one left the CPU generator
    (the device keeps a persistent raw stream; biased passes consume a data-dependent number of outputs)."""
    rowptr, col = random_csr(3000, 20, 21)
    w = build_weights('masked', rowptr, 21)
    wpos = build_weights('masked_pos', rowptr, 22)
    seed = torch.randperm(3000, generator=torch.Generator().manual_seed(4))[:200]
    d = lambda t: t.to(DEV)  # noqa: E731
    torch.manual_seed(5)
    got = [lib.sampler.neighbor_sample(d(rowptr), d(col), d(seed), [6, 4], edge_weight=d(w)),
           lib.sampler.neighbor_sample(d(rowptr), d(col), d(seed), [6, 4]),
           lib.sampler.neighbor_sample(d(rowptr), d(col), d(seed), [5, 5], edge_weight=d(wpos), replace=True),
           lib.sampler.neighbor_sample(d(rowptr), d(col), d(seed), [6, 4], edge_weight=d(w))]
    after = _rng()
    torch.manual_seed(5)
    exp = [W.neighbor_sample(rowptr, col, seed, [6, 4], w),
           O.neighbor_sample(rowptr, col, seed, [6, 4]),
           W.neighbor_sample(rowptr, col, seed, [5, 5], wpos, replace=True),
           W.neighbor_sample(rowptr, col, seed, [6, 4], w)]
    for g, e in zip(got, exp):
        _cmp(g, e)
    assert np.array_equal(after, _rng())


def test_pass_larger_than_the_default_stream_buffer(lib, live_oracle):
    """One hop that consumes 10 M engine outputs (200 rows of 50,000 neighbours): the raw-stream buffer grows mid-run."""
    n_hub, deg = 200, 50000
    rowptr = torch.zeros(302, dtype=torch.int64)
    rowptr[1:n_hub + 1] = deg
    rowptr[n_hub + 1:] = 3
    rowptr = rowptr.cumsum(0)
    g = torch.Generator().manual_seed(1)
    col = torch.randint(0, 301, (int(rowptr[-1]),), generator=g)
    w = torch.rand(col.numel(), generator=g)
    seed = torch.arange(0, n_hub)
    torch.manual_seed(3)
    out = lib.sampler.neighbor_sample(rowptr.to(DEV), col.to(DEV), seed.to(DEV), [5, 2], edge_weight=w.to(DEV))
    after = _rng()
    torch.manual_seed(3)
    exp = W.neighbor_sample(rowptr, col, seed, [5, 2], w)
    _cmp(out, exp)
    assert np.array_equal(after, _rng())


def test_full_neighbourhood_fanouts_ignore_the_weights(lib, live_oracle):
    rowptr, col = random_csr(500, 6, 2)
    seed = torch.arange(0, 20)
    w = torch.rand(col.numel())
    torch.manual_seed(1)
    a = lib.sampler.neighbor_sample(rowptr.to(DEV), col.to(DEV), seed.to(DEV), [-1, -1], edge_weight=w.to(DEV))
    ra = _rng()
    torch.manual_seed(1)
    b = W.neighbor_sample(rowptr, col, seed, [-1, -1], w)
    _cmp(a, b)
    assert np.array_equal(ra, _rng())


def test_errors(lib, live_oracle):
    rowptr, col = random_csr(300, 8, 3)
    seed = torch.arange(0, 30).to(DEV)
    rp, cl = rowptr.to(DEV), col.to(DEV)
    w = torch.rand(col.numel())
    with pytest.raises(RuntimeError, match='float32'):
        lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], edge_weight=w.double().to(DEV))
    with pytest.raises(RuntimeError, match='one entry per edge'):
        lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], edge_weight=w[:-1].to(DEV))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], edge_weight=w)
    with pytest.raises(RuntimeError, match='fan-out of 1'):
        lib.sampler.neighbor_sample(rp, cl, seed, [3, 1], edge_weight=w.to(DEV), replace=True)
    with pytest.raises(RuntimeError, match='Biased node temporal'):
        lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], node_time=torch.zeros(300, dtype=torch.long, device=DEV), edge_weight=w.to(DEV),
                                    disjoint=True)
    # at::multinomial's checks (replacement only; the reservoir keys take any float)
    wneg = w.clone(); wneg[int(rowptr[5])] = -1.0
    with pytest.raises(RuntimeError, match='invalid multinomial distribution'):
        lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], edge_weight=wneg.to(DEV), replace=True)
    wzero = torch.zeros_like(w)
    with pytest.raises(RuntimeError, match='sum of probabilities'):
        lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], edge_weight=wzero.to(DEV), replace=True)
    # ... and the sampler still works afterwards (the aborted run is cleaned up)
    torch.manual_seed(2)
    out = lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], edge_weight=w.to(DEV))
    torch.manual_seed(2)
    _cmp(out, W.neighbor_sample(rowptr, col, seed.cpu(), [3, 2], w))
    # negative / NaN weights without replacement flow through the key arithmetic like in the reference
    wodd = w.clone(); wodd[::5] = -wodd[::5]; wodd[3::11] = float('nan'); wodd[7::13] = float('inf')
    wodd[rowptr[:-1][(rowptr[:-1] < rowptr[1:]) & (torch.arange(300) % 3 == 0)]] = float('nan')   # NaN key at in-row index 0
    torch.manual_seed(4)
    out = lib.sampler.neighbor_sample(rp, cl, seed, [3, 2], edge_weight=wodd.to(DEV))
    torch.manual_seed(4)
    _cmp(out, W.neighbor_sample(rowptr, col, seed.cpu(), [3, 2], wodd))


@pytest.mark.parametrize('replace', [False, True])
@pytest.mark.parametrize('disjoint', [False, True])
def test_weighted_dist_neighbor_sample_is_one_hop_of_neighbor_sample(lib, live_oracle, replace, disjoint):
    """pyg::dist_neighbor_sample with edge_weight (biased_sample with distributed = true, neighbor_kernel.cpp:296-303,436-448): the
    same draws as the first hop of neighbor_sample, nothing mapped — global ids, edge ids, cumulative counts per seed."""
    rowptr, col = random_csr(4000, 18, 12, big=[(3, 9000)])
    w = build_weights('masked_pos' if replace else 'masked', rowptr, 5)
    seed = torch.cat([torch.tensor([3]), torch.randperm(4000, generator=torch.Generator().manual_seed(2))[:300]])
    k = 7
    torch.manual_seed(8)
    node, eid, cum = torch.ops.pyg.dist_neighbor_sample(rowptr.to(DEV), col.to(DEV), seed.to(DEV), k, None, None, None, w.to(DEV), True,
                                                        replace, True, disjoint, 'uniform')
    after = _rng()
    torch.manual_seed(8)
    exp = W.neighbor_sample(rowptr, col, seed, [k], w, replace=replace, disjoint=disjoint)
    assert np.array_equal(after, _rng())
    assert torch.equal(eid.cpu(), exp[3])
    S = seed.numel()
    dst = col[exp[3]]
    counts = torch.bincount(exp[0], minlength=S)
    assert list(cum) == [S] + (S + torch.cumsum(counts, 0)).tolist()   # neighbor_kernel.cpp:386,447
    if not disjoint:
        assert torch.equal(node.cpu(), torch.cat([seed, dst]))
    else:
        assert torch.equal(node.cpu(), torch.stack([torch.cat([torch.arange(S), exp[0]]), torch.cat([seed, dst])], 1))
