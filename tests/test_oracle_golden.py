"""CPU: pin the oracle restatement (oracle/*.c) to (1) the reference's own known-answer vectors and
(2) outputs of the reference itself (tests/golden/reference_outputs.npz, made by make_golden.py)."""
import numpy as np
import pytest
import torch

from graphs import (HETERO_CASES, HOMO_CASES, MATMUL_CASES, build_hetero, build_homo, build_matmul, build_temporal,
                    cycle_graph)
from oracle import oracle as O


def test_randint_stream_matches_torch():
    """oracle mt19937 + randint transform == at::randint(INT64_MIN, INT64_MAX) (rand_engine.h:80-85)."""
    torch.manual_seed(123456)
    mt = O.mt_from_torch()
    w = O.randint_words(mt, 1000)
    t = torch.randint(-2 ** 63, 2 ** 63 - 1, (1000,), dtype=torch.int64).numpy().view(np.uint64)
    assert (w == t).all()
    mt2 = O.mt_from_torch()
    assert mt.left == mt2.left and mt.next == mt2.next
    assert (np.ctypeslib.as_array(mt.state) == np.ctypeslib.as_array(mt2.state)).all()


def test_mt_seed_matches_manual_seed():
    torch.manual_seed(987654321)
    a, b = O.mt_from_torch(), O.mt_seed(987654321)
    assert a.left == b.left and a.next == b.next
    assert (np.ctypeslib.as_array(a.state) == np.ctypeslib.as_array(b.state)).all()


# ---- reference known-answer vectors, test/csrc/sampler/test_neighbor.cpp -------------------------
def test_kat_full_neighborhood():  # :8-31
    rowptr, col = cycle_graph(6)
    row, colv, node, eid, nph, eph = O.neighbor_sample(rowptr, col, torch.arange(2, 4), [-1, -1])
    assert row.tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    assert colv.tolist() == [2, 1, 0, 3, 4, 0, 1, 5]
    assert node.tolist() == [2, 3, 1, 4, 0, 5]
    assert eid.tolist() == [4, 5, 6, 7, 2, 3, 8, 9]
    assert nph == [2, 2, 2] and eph == [4, 4]


def test_kat_zero_degree():  # :33-57
    rowptr, col = torch.zeros(6, dtype=torch.int64), torch.zeros(0, dtype=torch.int64)
    row, colv, node, eid, nph, eph = O.neighbor_sample(rowptr, col, torch.arange(0, 5), [-1, -1])
    assert row.numel() == 0 and colv.numel() == 0 and eid.numel() == 0
    assert node.tolist() == [0, 1, 2, 3, 4] and nph == [5, 0, 0] and eph == [0, 0]


@pytest.mark.parametrize('replace', [False, True])
def test_kat_seeded(replace):  # :59-113
    rowptr, col = cycle_graph(6)
    torch.manual_seed(123456)
    row, colv, node, eid, _, _ = O.neighbor_sample(rowptr, col, torch.arange(2, 4), [1, 1], replace=replace)
    assert row.tolist() == [0, 1, 2, 3] and colv.tolist() == [2, 3, 0, 4]
    assert node.tolist() == [2, 3, 1, 4, 5] and eid.tolist() == [4, 7, 3, 9]


def test_kat_disjoint():  # :115-144
    rowptr, col = cycle_graph(6)
    row, colv, node, eid, _, _ = O.neighbor_sample(rowptr, col, torch.arange(2, 4), [2, 2], disjoint=True)
    assert row.tolist() == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5]
    assert colv.tolist() == [2, 3, 4, 5, 6, 0, 0, 7, 8, 1, 1, 9]
    assert node.flatten().tolist() == [0, 2, 1, 3, 0, 1, 0, 3, 1, 2, 1, 4, 0, 0, 0, 4, 1, 1, 1, 5]
    assert eid.tolist() == [4, 5, 6, 7, 2, 3, 6, 7, 4, 5, 8, 9]


def test_kat_hetero_single_relation():  # :259-298
    nt, et, rp, cl, sd, nn = build_hetero(HETERO_CASES['cycle_single_rel'])
    row, colv, node, eid, nph, eph = O.hetero_neighbor_sample(nt, et, rp, cl, sd, nn)
    k = 'paper__to__paper'
    assert row[k].tolist() == [0, 0, 1, 1, 2, 2, 3, 3] and colv[k].tolist() == [2, 1, 0, 3, 4, 0, 1, 5]
    assert node['paper'].tolist() == [2, 3, 1, 4, 0, 5] and eid[k].tolist() == [4, 5, 6, 7, 2, 3, 8, 9]
    assert nph['paper'] == [2, 2, 2] and eph[k] == [4, 4]


# ---- fixtures produced by the reference itself ---------------------------------------------------
def _rng_prefix():
    return torch.get_rng_state().numpy()[:24 + 624 * 8]


@pytest.mark.parametrize('name', list(HOMO_CASES))
def test_homo_matches_reference(name, golden):
    case = HOMO_CASES[name]
    rowptr, col, seed = build_homo(case)
    nt = et = stt = None
    if 'temporal' in case:
        col, nt, et, stt = build_temporal(case, rowptr, col, seed)
    torch.manual_seed(case['rng_seed'])
    row, colv, node, eid, nph, eph = O.neighbor_sample(rowptr, col, seed, case['num_neighbors'], node_time=nt,
                                                        edge_time=et, seed_time=stt, csc=case.get('csc', False),
                                                        replace=case.get('replace', False),
                                                        disjoint=case.get('disjoint', False),
                                                        temporal_strategy=case.get('strategy', 'uniform'))
    p = f'homo/{name}/'
    assert np.array_equal(row.numpy(), golden[p + 'row'])
    assert np.array_equal(colv.numpy(), golden[p + 'col'])
    assert np.array_equal(node.numpy(), golden[p + 'node'])
    assert np.array_equal(eid.numpy(), golden[p + 'eid'])
    assert nph == golden[p + 'nph'].tolist() and eph == golden[p + 'eph'].tolist()
    # the default CPU generator must be left exactly where the reference leaves it
    assert np.array_equal(_rng_prefix(), golden[p + 'rng_after'])


@pytest.mark.parametrize('name', list(HETERO_CASES))
def test_hetero_matches_reference(name, golden):
    case = HETERO_CASES[name]
    nt, et, rp, cl, sd, nn = build_hetero(case)
    torch.manual_seed(case['rng_seed'])
    row, colv, node, eid, nph, eph = O.hetero_neighbor_sample(nt, et, rp, cl, sd, nn,
                                                               csc=case.get('csc', False),
                                                               replace=case.get('replace', False),
                                                               disjoint=case.get('disjoint', False))
    p = f'hetero/{name}/'
    for k in rp:
        assert np.array_equal(row[k].numpy(), golden[p + 'row/' + k]), k
        assert np.array_equal(colv[k].numpy(), golden[p + 'col/' + k]), k
        assert np.array_equal(eid[k].numpy(), golden[p + 'eid/' + k]), k
        assert eph[k] == golden[p + 'eph/' + k].tolist()
    for t in nt:
        assert np.array_equal(node[t].numpy(), golden[p + 'node/' + t]), t
        assert nph[t] == golden[p + 'nph/' + t].tolist()
    assert np.array_equal(_rng_prefix(), golden[p + 'rng_after'])


@pytest.mark.parametrize('name', list(MATMUL_CASES))
def test_matmul_matches_reference(name, golden):
    case = MATMUL_CASES[name]
    x, ptr, w = build_matmul(case)
    out = O.segment_matmul(x, ptr, w).float().numpy()
    ref = golden[f'matmul/{name}/out']
    if case['dtype'] == 'float32':
        assert np.allclose(out, ref, atol=1e-5, rtol=1e-5)
    else:  # one storage-dtype ulp elementwise + 1e-3 relative Frobenius
        assert np.linalg.norm(out - ref) <= 1e-3 * np.linalg.norm(ref)
        ulp = 2.0 ** -7 if case['dtype'] == 'bfloat16' else 2.0 ** -10
        assert (np.abs(out - ref) <= ulp * np.maximum(np.abs(ref), 2.0 ** -14) + 1e-30).all()
    # and against a plain fp32 matmul (the reference's own test oracle, test_matmul.py:38-44)
    for i in range(ptr.numel() - 1):
        a, b = int(ptr[i]), int(ptr[i + 1])
        exp = (x[a:b].float() @ w[i].float()).numpy()
        tol = 1e-5 if case['dtype'] == 'float32' else 2e-2
        assert np.allclose(out[a:b], exp, atol=tol, rtol=tol)


# ---- temporal known-answer vectors, test/csrc/sampler/test_neighbor.cpp:146-257 --------------------------
def test_kat_node_temporal():
    rowptr, col = cycle_graph(6)
    node_time = torch.arange(6)
    col = torch.sort(col.view(-1, 2), dim=1).values.flatten()
    out1 = O.neighbor_sample(rowptr, col, torch.arange(2, 4), [2, 2], node_time=node_time, disjoint=True)
    assert out1[0].tolist() == [0, 1, 2, 2, 3, 3] and out1[1].tolist() == [2, 3, 4, 0, 5, 1]
    assert out1[2].flatten().tolist() == [0, 2, 1, 3, 0, 1, 1, 2, 0, 0, 1, 1]
    assert out1[3].tolist() == [4, 6, 2, 3, 4, 5]
    out2 = O.neighbor_sample(rowptr, col, torch.arange(2, 4), [1, 2], node_time=node_time, disjoint=True,
                             temporal_strategy='last')
    for a, b in zip(out1[:4], out2[:4]):
        assert torch.equal(a, b)


def test_kat_edge_temporal():
    rowptr, col = cycle_graph(6)
    edge_time = torch.arange(col.numel())
    out = O.neighbor_sample(rowptr, col, torch.arange(2, 4), [2, 2], edge_time=edge_time, seed_time=torch.arange(5, 7),
                            disjoint=True)
    assert out[0].tolist() == [0, 0, 1, 2, 2, 4, 4] and out[1].tolist() == [2, 3, 4, 5, 0, 6, 1]
    assert out[2].flatten().tolist() == [0, 2, 1, 3, 0, 1, 0, 3, 1, 2, 0, 0, 1, 1]
    assert out[3].tolist() == [4, 5, 6, 2, 3, 4, 5]
    out2 = O.neighbor_sample(rowptr, col, torch.arange(2, 4), [1, 1], edge_time=edge_time,
                             seed_time=torch.tensor([-1, -1]), disjoint=True, replace=True)
    assert out2[0].numel() == 0 and out2[2].flatten().tolist() == [0, 2, 1, 3]


# ------------------------------------------------------------------------------------ subgraph
import os.path as osp  # noqa: E402
from graphs import SUBGRAPH_CASES, build_subgraph  # noqa: E402


@pytest.fixture(scope='module')
def subgraph_golden():
    return np.load(osp.join(osp.dirname(osp.abspath(__file__)), 'golden', 'subgraph_outputs.npz'))


def test_subgraph_kat():
    """test/csrc/sampler/test_subgraph.cpp:7-24."""
    rowptr, col, nodes = build_subgraph(SUBGRAPH_CASES['cycle_kat'])
    r = O.subgraph(rowptr, col, nodes)
    assert r[0].tolist() == [0, 1, 3, 5, 6] and r[1].tolist() == [1, 0, 2, 1, 3, 2] and r[2].tolist() == [3, 4, 5, 6, 7, 8]


@pytest.mark.parametrize('name', list(SUBGRAPH_CASES))
def test_subgraph_oracle_matches_reference_fixture(subgraph_golden, name):
    rowptr, col, nodes = build_subgraph(SUBGRAPH_CASES[name])
    r = O.subgraph(rowptr, col, nodes)
    assert np.array_equal(r[0].numpy(), subgraph_golden[f'{name}/rowptr'])
    assert np.array_equal(r[1].numpy(), subgraph_golden[f'{name}/col'])
    assert np.array_equal(r[2].numpy(), subgraph_golden[f'{name}/eid'])
    assert O.subgraph(rowptr, col, nodes, return_edge_id=False)[2] is None


# ------------------------------------------------------------------------------------ dist_neighbor_sample
from graphs import DIST_CASES, build_dist  # noqa: E402


@pytest.fixture(scope='module')
def dist_golden():
    return np.load(osp.join(osp.dirname(osp.abspath(__file__)), 'golden', 'dist_outputs.npz'))


def test_dist_neighbor_sample_kats():
    """test/csrc/sampler/test_dist_neighbor.cpp:8-77 (full, without / with replacement)."""
    rowptr, col = cycle_graph(6)
    r = O.dist_neighbor_sample(rowptr, col, torch.arange(2, 4), -1)
    assert r[0].tolist() == [2, 3, 1, 3, 2, 4] and r[1].tolist() == [4, 5, 6, 7] and r[2] == [2, 4, 6]
    torch.manual_seed(123456)
    r = O.dist_neighbor_sample(rowptr, col, torch.arange(2, 4), 1)
    assert r[0].tolist() == [2, 3, 1, 4] and r[1].tolist() == [4, 7] and r[2] == [2, 3, 4]
    torch.manual_seed(123456)
    r = O.dist_neighbor_sample(rowptr, col, torch.arange(2, 4), 2, replace=True)
    assert r[0].tolist() == [2, 3, 1, 3, 4, 4] and r[1].tolist() == [4, 5, 7, 7] and r[2] == [2, 4, 6]


@pytest.mark.parametrize('name', list(DIST_CASES))
def test_dist_oracle_matches_reference_fixture(dist_golden, name):
    case = DIST_CASES[name]
    rowptr, col, seed = build_dist(case)
    torch.manual_seed(case['rng_seed'])
    r = O.dist_neighbor_sample(rowptr, col, seed, case['k'], replace=case.get('replace', False), disjoint=case.get('disjoint', False))
    assert np.array_equal(r[0].numpy(), dist_golden[f'{name}/node'])
    assert np.array_equal(r[1].numpy(), dist_golden[f'{name}/eid'])
    assert list(r[2]) == dist_golden[f'{name}/cumsum'].tolist()
    assert np.array_equal(torch.get_rng_state().numpy()[:24 + 624 * 8], dist_golden[f'{name}/rng_after'])


# ------------------------------------------------------------------------------------ relabel_neighborhood
from graphs import RELABEL_CASES, build_relabel  # noqa: E402


@pytest.mark.parametrize('name', list(RELABEL_CASES))
def test_relabel_oracle_matches_reference_fixture(name):
    """KATs test/csrc/sampler/test_dist_relabel.cpp:9-60 + cases derived from one-hop distributed samples."""
    G = np.load(osp.join(osp.dirname(osp.abspath(__file__)), 'golden', 'relabel_outputs.npz'))
    seed, sampled, counts, num_nodes, batch, csc, disjoint = build_relabel(RELABEL_CASES[name], O.dist_neighbor_sample)
    r = O.relabel_neighborhood(seed, sampled, counts, num_nodes, batch, csc, disjoint)
    assert np.array_equal(r[0].numpy(), G[f'{name}/row']) and np.array_equal(r[1].numpy(), G[f'{name}/col'])
    if name == 'kat':
        assert r[0].tolist() == [0, 0, 1, 1] and r[1].tolist() == [2, 1, 0, 3]
    if name == 'kat_disjoint':
        assert r[0].tolist() == [0, 0, 1, 1] and r[1].tolist() == [2, 3, 4, 5]


# ------------------------------------------------------------------------------------ hetero_relabel_neighborhood
from graphs import HETERO_RELABEL_CASES, build_hetero_relabel  # noqa: E402


@pytest.mark.parametrize('name', list(HETERO_RELABEL_CASES))
def test_hetero_relabel_oracle_matches_reference_fixture(name):
    G = np.load(osp.join(osp.dirname(osp.abspath(__file__)), 'golden', 'hetero_relabel_outputs.npz'))
    r = O.hetero_relabel_neighborhood(*build_hetero_relabel(HETERO_RELABEL_CASES[name]))
    for k in r[0]:
        assert np.array_equal(r[0][k].numpy(), G[f'{name}/row/{k}']) and np.array_equal(r[1][k].numpy(), G[f'{name}/col/{k}']), k


# ------------------------------------------------------------------------------------ merge_sampler_outputs
from graphs import MERGE_CASES, build_merge  # noqa: E402


@pytest.mark.parametrize('name', list(MERGE_CASES))
def test_merge_oracle_and_host_plan_match_reference_fixture(name):
    """The oracle restatement AND the product's host-side plan (pygb200_merge_plan: pure arithmetic, callable without a
    GPU; the gathers it drives are emulated with numpy here) against fixtures made by the reference
    (KATs test/csrc/sampler/test_dist_merge_outputs.cpp:7-91 + partitioned random cases)."""
    import ctypes as C
    G = np.load(osp.join(osp.dirname(osp.abspath(__file__)), 'golden', 'merge_outputs.npz'))
    a = build_merge(MERGE_CASES[name], O.dist_neighbor_sample)
    r = O.merge_sampler_outputs(*a)
    assert np.array_equal(r[0].numpy(), G[f'{name}/node']) and np.array_equal(r[1].numpy(), G[f'{name}/edge'])
    assert list(r[3]) == G[f'{name}/counts'].tolist()
    if a[8]:
        assert np.array_equal(r[2].numpy(), G[f'{name}/batch'])
    # ---- host plan of the product
    from pyg_lib_b200.build import build
    build(verbose=False)
    lib = C.CDLL(osp.join(osp.dirname(osp.dirname(osp.abspath(__file__))), 'pyg_lib_b200', 'libpyg_b200.so'))
    lib.pygb200_last_error.restype = C.c_char_p
    node_ids, edge_ids, cums, pids, pords, P = a[0], a[1], a[2], a[3], a[4], a[5]
    flat = [x for c in cums for x in c]
    off = np.cumsum([0] + [len(c) for c in cums]).astype(np.int64)
    n = len(pids)
    arr = lambda x: (C.c_int64 * max(len(x), 1))(*x)  # noqa: E731
    nb, eb, cnt = (C.c_int64 * max(n, 1))(), (C.c_int64 * max(n, 1))(), (C.c_int64 * max(n, 1))()
    rc = lib.pygb200_merge_plan(arr(flat), arr(off.tolist()), C.c_int64(P), arr(pids), arr(pords), C.c_int64(n),
                                arr([t.numel() for t in node_ids]), arr([t.numel() for t in edge_ids]), nb, eb, cnt)
    assert rc == 0, lib.pygb200_last_error()
    assert list(cnt)[:n] == G[f'{name}/counts'].tolist()
    nodes = np.concatenate([node_ids[pids[j]].numpy()[nb[j]:nb[j] + cnt[j]] for j in range(n)]) if n else np.zeros(0, np.int64)
    edges = np.concatenate([edge_ids[pids[j]].numpy()[eb[j]:eb[j] + cnt[j]] for j in range(n)]) if n else np.zeros(0, np.int64)
    assert np.array_equal(nodes, G[f'{name}/node']) and np.array_equal(edges, G[f'{name}/edge'])
    bad = arr([p + 100 for p in pids])
    assert lib.pygb200_merge_plan(arr(flat), arr(off.tolist()), C.c_int64(P), bad, arr(pords), C.c_int64(n),
                                  arr([t.numel() for t in node_ids]), arr([t.numel() for t in edge_ids]), nb, eb, cnt) != 0
