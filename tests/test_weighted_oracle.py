"""CPU: the biased (edge_weight) sampling oracle and the word-level model of the CUDA algorithm against fixtures produced by the
reference itself (tests/golden/make_golden_weighted.py), plus the pieces the model is made of against ATen on this host."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import weighted_model as M  # noqa: E402
from graphs import (HETERO_WEIGHTED_CASES, WEIGHTED_CASES, build_hetero_weighted, build_weighted)  # noqa: E402
from oracle import weighted as W  # noqa: E402

GOLD = np.load(os.path.join(HERE, 'golden', 'weighted_outputs.npz'))


def rng_prefix():
    return torch.get_rng_state().numpy()[:24 + 624 * 8]


@pytest.mark.parametrize('name', list(WEIGHTED_CASES))
def test_oracle_matches_reference_fixture(name):
    case = WEIGHTED_CASES[name]
    rowptr, col, seed, w = build_weighted(case)
    torch.manual_seed(case['rng_seed'])
    o = W.neighbor_sample(rowptr, col, seed, case['num_neighbors'], w, replace=case.get('replace', False), csc=case.get('csc', False),
                          disjoint=case.get('disjoint', False))
    for k, v in zip(('row', 'col', 'node', 'eid'), o[:4]):
        assert np.array_equal(v.numpy(), GOLD[f'homo/{name}/{k}']), k
    assert o[4] == GOLD[f'homo/{name}/nph'].tolist() and o[5] == GOLD[f'homo/{name}/eph'].tolist()
    assert np.array_equal(rng_prefix(), GOLD[f'homo/{name}/rng_after'])


@pytest.mark.parametrize('name', [n for n, c in HETERO_WEIGHTED_CASES.items() if not c.get('disjoint') and 'weighted_rels' not in c])
def test_hetero_oracle_matches_reference_fixture(name):
    case = HETERO_WEIGHTED_CASES[name]
    nt, et, rp, cl, sd, nn, wd = build_hetero_weighted(case)
    torch.manual_seed(case['rng_seed'])
    o = W.hetero_neighbor_sample(nt, et, rp, cl, sd, nn, wd, replace=case.get('replace', False), csc=case.get('csc', False))
    for i, key in enumerate(('row', 'col', 'node', 'eid')):
        for k, v in o[i].items():
            assert np.array_equal(v.numpy(), GOLD[f'hetero/{name}/{key}/{k}']), (key, k)
    for k, v in o[4].items():
        assert v == GOLD[f'hetero/{name}/nph/{k}'].tolist()
    for k, v in o[5].items():
        assert v == GOLD[f'hetero/{name}/eph/{k}'].tolist()
    assert np.array_equal(rng_prefix(), GOLD[f'hetero/{name}/rng_after'])


@pytest.mark.parametrize('name', [n for n, c in WEIGHTED_CASES.items() if 'hub' not in n and not c.get('disjoint')])
def test_word_level_model_matches_reference_fixture(name):
    """The algorithm the kernels implement (raw engine words, log table, parallel top-k + libstdc++ replay on ties,
    float32 running sums) gives the reference's result."""
    case = WEIGHTED_CASES[name]
    rowptr, col, seed, w = build_weighted(case)
    torch.manual_seed(case['rng_seed'])
    stats = {}
    o = M.neighbor_sample(rowptr, col, seed, case['num_neighbors'], w, replace=case.get('replace', False), stats=stats)
    row, colv = (o[1], o[0]) if case.get('csc', False) else (o[0], o[1])
    for k, v in zip(('row', 'col', 'node', 'eid'), (row, colv, o[2], o[3])):
        assert np.array_equal(v.numpy(), GOLD[f'homo/{name}/{k}']), k
    if case['weights'] in ('masked', 'quantized') and not case.get('replace', False):
        assert stats.get('ties', 0) > 0   # the libstdc++ replay is exercised


def test_topk_restatement_matches_aten_on_ties():
    rng = np.random.RandomState(1)
    for trial in range(400):
        n = int(rng.choice([2, 3, 4, 5, 8, 17, 40, 100, 300, 1000, 3000]))
        k = int(rng.randint(1, min(n, 40)))
        if trial % 4 == 0:
            n = max(n, 64 * k + int(rng.randint(0, 50)))   # std::partial_sort branch
        kind = trial % 5
        if kind == 0: v = rng.rand(n)
        elif kind == 1: v = rng.randint(0, 4, n).astype(np.float64)
        elif kind == 2: v = np.where(rng.rand(n) < 0.7, -np.inf, rng.rand(n))
        elif kind == 3: v = np.where(rng.rand(n) < 0.1, np.nan, rng.randint(0, 10, n).astype(np.float64))
        else: v = np.full(n, -np.inf)
        v = v.astype(np.float32)
        assert torch.from_numpy(v).topk(k)[1].tolist() == M.topk_libstdcxx(v.tolist(), k), (n, k, kind)
        idx, tie = M.topk_gpu(v.tolist(), k)
        if not tie:
            assert idx == torch.from_numpy(v).topk(k)[1].tolist()


def test_multinomial_and_uniform_restatements_match_aten():
    rng = np.random.RandomState(0)
    for trial in range(60):
        n = int(rng.randint(2, 200)); k = int(rng.randint(2, 20))
        wt = rng.rand(n).astype(np.float32)
        if trial % 3 == 0:
            wt[rng.rand(n) < 0.5] = 0
            wt[0] = 1
        torch.manual_seed(trial)
        mine = M.biased_indices(M.Words(), wt, k, True)
        assert torch.multinomial(torch.from_numpy(wt), k, True).tolist() == mine
        torch.manual_seed(trial)
        words = M.Words().take(n)
        u = (words & np.uint32(0xFFFFFF)).astype(np.float32) * np.float32(2.0 ** -24)
        assert np.array_equal(torch.empty(n).uniform_().numpy(), u)


def test_logf_table_describes_this_hosts_torch_log():
    """torch.log(float32) on CPU == correctly rounded log + mkl_logf_table.inc, on all 2^24 values uniform_ can produce.
    (If this host's MKL took another code path the table would still pin the GPU to the fixtures, but not to a reference run
    on this host — so it is checked.)"""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tools'))
    from make_logf_table import read_inc
    t = read_inc()
    k = np.arange(0, 1 << 24, dtype=np.int64)
    u = k.astype(np.float32) * np.float32(2.0 ** -24)
    with np.errstate(divide='ignore'):
        bits = np.log(u.astype(np.float64)).astype(np.float32).view(np.int32).copy()
    bits[(t >> 1).astype(np.int64)] += np.where(t & 1, 1, -1).astype(np.int32)
    assert np.array_equal(torch.log(torch.from_numpy(u)).numpy().view(np.int32), bits)
    sample = np.array([0, 1, 51707, 146089, 16763221, (1 << 24) - 1], dtype=np.int64)
    assert np.array_equal(M.mkl_logf(sample).view(np.int32), bits[sample])


def test_device_topk_replay_compiled_for_the_host_matches_aten(tmp_path):
    """pyg_lib_b200/csrc/topk_replay.h — the code one lane of k_w_sample runs on a tie — built with g++ and compared with
    torch.topk on tie-heavy inputs (both branches: std::partial_sort and std::nth_element + std::sort, depth-limit fallbacks)."""
    import ctypes
    import subprocess
    so = str(tmp_path / 'libreplay.so')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', os.path.join(HERE, 'cpp', 'topk_replay_check.cpp'), '-o', so])
    lib = ctypes.CDLL(so)
    rng = np.random.RandomState(5)
    for trial in range(1500):
        n = int(rng.choice([2, 3, 4, 5, 8, 17, 33, 100, 300, 1000, 5000, 70000] if trial % 50 == 0 else [2, 3, 4, 5, 8, 17, 33, 100, 300, 1000]))
        k = int(rng.randint(1, min(n, 130)))
        if trial % 4 == 0:
            n = max(n, 64 * k + int(rng.randint(0, 50)))
        kind = trial % 6
        if kind == 0: v = rng.rand(n)
        elif kind == 1: v = rng.randint(0, 4, n).astype(np.float64)
        elif kind == 2: v = np.where(rng.rand(n) < 0.7, -np.inf, rng.rand(n))
        elif kind == 3: v = np.where(rng.rand(n) < 0.1, np.nan, rng.randint(0, 10, n).astype(np.float64))
        elif kind == 4: v = np.full(n, -np.inf)
        else: v = np.sort(rng.randint(0, 50, n).astype(np.float64))[::-1 if trial % 12 < 6 else 1]   # sorted runs: bad pivots
        K = np.ascontiguousarray(v, dtype=np.float32)
        ref = torch.from_numpy(K.copy()).topk(k)[1].tolist()
        I = np.arange(n, dtype=np.uint32)
        lib.topk_replay(K.ctypes.data_as(ctypes.c_void_p), I.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int(k))
        assert I[:k].tolist() == ref, (n, k, kind)
