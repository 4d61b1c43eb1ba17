"""CPU: the C-ABI library builds, loads, and exports every symbol include/pyg_b200.h declares
(no compute calls without a GPU); the dispatcher layer registers the reference's schemas verbatim."""
import ctypes as C
import os.path as osp
import re

import pytest
import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


def declared_symbols():
    src = open(osp.join(ROOT, 'include', 'pyg_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pygb200_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def built():
    from pyg_lib_b200.build import build
    build(verbose=False)
    return osp.join(ROOT, 'pyg_lib_b200')


def test_header_symbols_exported(built):
    lib = C.CDLL(osp.join(built, 'libpyg_b200.so'))
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/pyg_b200.h but not exported'
    lib.pygb200_cuda_version.restype = C.c_int
    assert lib.pygb200_cuda_version() >= 12000
    lib.pygb200_last_error.restype = C.c_char_p
    # argument validation happens before any CUDA call
    assert lib.pygb200_profile_read(b'no_such_kernel', None, None, None) == -2
    assert b'unknown kernel name' in lib.pygb200_last_error()


def test_sampler_create_fails_loudly_without_gpu(built):
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    lib = C.CDLL(osp.join(built, 'libpyg_b200.so'))
    lib.pygb200_last_error.restype = C.c_char_p
    h = C.c_void_p()
    assert lib.pygb200_sampler_create(C.byref(h)) != 0   # no silent CPU path
    assert lib.pygb200_last_error()


def test_schemas_match_reference(built):
    import pyg_lib_b200  # noqa: F401
    s = str(torch.ops.pyg.neighbor_sample.default._schema)
    assert s == ("pyg::neighbor_sample(Tensor rowptr, Tensor col, Tensor seed, int[] num_neighbors, Tensor? node_time=None, "
                 "Tensor? edge_time=None, Tensor? seed_time=None, Tensor? edge_weight=None, bool csc=False, bool replace=False, "
                 "bool directed=True, bool disjoint=False, str temporal_strategy=\"uniform\", bool return_edge_id=True) -> "
                 "(Tensor, Tensor, Tensor, Tensor?, int[], int[])")
    h = str(torch.ops.pyg.hetero_neighbor_sample.default._schema)
    assert h.startswith('pyg::hetero_neighbor_sample(str[] node_types, (str, str, str)[] edge_types, Dict(str, Tensor) rowptr_dict')
    assert h.endswith('-> (Dict(str, Tensor), Dict(str, Tensor), Dict(str, Tensor), Dict(str, Tensor)?, Dict(str, int[]), Dict(str, int[]))')
    assert str(torch.ops.pyg.segment_matmul.default._schema) == 'pyg::segment_matmul(Tensor input, Tensor ptr, Tensor other) -> Tensor'
    assert str(torch.ops.pyg.subgraph.default._schema) == ('pyg::subgraph(Tensor rowptr, Tensor col, Tensor nodes, bool return_edge_id) '
                                                            '-> (Tensor, Tensor, Tensor?)')   # sampler/subgraph.cpp:28-32
    assert str(torch.ops.pyg.dist_neighbor_sample.default._schema) == (   # sampler/neighbor.cpp:148-153
        'pyg::dist_neighbor_sample(Tensor rowptr, Tensor col, Tensor seed, int num_neighbors, Tensor? node_time=None, '
        'Tensor? edge_time=None, Tensor? seed_time=None, Tensor? edge_weight=None, bool csc=False, bool replace=False, '
        'bool directed=True, bool disjoint=False, str temporal_strategy="uniform") -> (Tensor, Tensor, int[])')
    assert str(torch.ops.pyg.relabel_neighborhood.default._schema) == (   # sampler/dist_relabel.cpp:71-76
        'pyg::relabel_neighborhood(Tensor seed, Tensor sampled_nodes_with_duplicates, int[] num_sampled_neighbors_per_node, '
        'int num_nodes, Tensor? batch=None, bool csc=False, bool disjoint=False) -> (Tensor, Tensor)')
    assert str(torch.ops.pyg.hetero_relabel_neighborhood.default._schema) == (   # sampler/dist_relabel.cpp:77-83
        'pyg::hetero_relabel_neighborhood(str[] node_types, (str, str, str)[] edge_types, Dict(str, Tensor) seed_dict, '
        'Dict(str, Tensor) sampled_nodes_with_duplicates_dict, Dict(str, int[][]) num_sampled_neighbors_per_node_dict, '
        'Dict(str, int) num_nodes_dict, Dict(str, Tensor)? batch_dict=None, bool csc=False, bool disjoint=False) -> '
        '(Dict(str, Tensor), Dict(str, Tensor))')
    assert str(torch.ops.pyg.merge_sampler_outputs.default._schema) == (   # sampler/dist_merge_outputs.cpp:51-55
        'pyg::merge_sampler_outputs(Tensor[] node_ids, Tensor[] edge_ids, int[][] cumsum_neighbors_per_node, int[] partition_ids, '
        'int[] partition_orders, int num_partitions, int num_neighbors, Tensor? batch, bool disjoint) -> (Tensor, Tensor, Tensor?, int[])')
    assert str(torch.ops.pyg.grouped_matmul.default._schema) == 'pyg::grouped_matmul(Tensor[] input, Tensor[] other) -> Tensor[]'
    assert torch.ops.pyg.cuda_version() >= 12000


def test_no_cpu_fallback(built):
    import pyg_lib
    with pytest.raises((RuntimeError, NotImplementedError)):
        pyg_lib.ops.segment_matmul(torch.randn(4, 8), torch.tensor([0, 4]), torch.randn(1, 8, 8))
    with pytest.raises((RuntimeError, NotImplementedError)):
        pyg_lib.sampler.neighbor_sample(torch.arange(3), torch.arange(2), torch.arange(1), [1])
    with pytest.raises(RuntimeError, match='CUDA tensors'):
        pyg_lib.sampler.hetero_neighbor_sample({('a', 'to', 'a'): torch.arange(3)}, {('a', 'to', 'a'): torch.arange(2)},
                                               {'a': torch.arange(1)}, {('a', 'to', 'a'): [1]})


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (a product path through the oracle voids parity)."""
    import glob
    for f in glob.glob(osp.join(ROOT, 'pyg_lib_b200', '**', '*'), recursive=True) + glob.glob(osp.join(ROOT, 'pyg_lib', '*.py')):
        if osp.isfile(f) and f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
            txt = open(f, errors='ignore').read()
            assert 'import oracle' not in txt and 'from oracle' not in txt and 'liboracle' not in txt, f


def test_sampler_bounds_host_only(built):
    """pygb200_sampler_bounds is pure host arithmetic (no CUDA call): the static frontier/edge recurrence of
    neighbor_kernel.cpp:430-475,718-812 — every frontier node emits <= k edges, every edge <= one new node of the
    relation's dst type — checked against the oracle's actual counts on a hetero case."""
    import sys
    sys.path.insert(0, osp.join(ROOT, 'tests'))
    from graphs import HETERO_CASES, build_hetero
    from oracle import oracle as O
    lib = C.CDLL(osp.join(built, 'libpyg_b200.so'))
    lib.pygb200_last_error.restype = C.c_char_p

    class REL(C.Structure):
        _fields_ = [('rowptr', C.c_void_p), ('col', C.c_void_p), ('num_src_nodes', C.c_int64), ('num_edges', C.c_int64),
                    ('src_type', C.c_int32), ('dst_type', C.c_int32)]

    # homogeneous closed form
    rel = REL(None, None, 100, 1000, 0, 0)
    ns, nn = C.c_int64(7), (C.c_int64 * 3)(5, 4, 3)
    ncap, ecap = C.c_int64(), C.c_int64()
    assert lib.pygb200_sampler_bounds(1, 1, 3, C.byref(rel), C.byref(ns), nn, C.byref(ncap), C.byref(ecap)) == 0
    assert ecap.value == 7 * 5 + 7 * 5 * 4 + 7 * 5 * 4 * 3 and ncap.value == 7 + ecap.value
    nn_all = (C.c_int64 * 3)(5, -1, 3)
    assert lib.pygb200_sampler_bounds(1, 1, 3, C.byref(rel), C.byref(ns), nn_all, C.byref(ncap), C.byref(ecap)) != 0
    assert b'-1' in lib.pygb200_last_error()
    bad = REL(None, None, 100, 1000, 0, 3)
    assert lib.pygb200_sampler_bounds(1, 1, 3, C.byref(bad), C.byref(ns), nn, C.byref(ncap), C.byref(ecap)) != 0

    # hetero: bounds dominate what the oracle (= the reference's algorithm) actually samples
    name = next(iter(HETERO_CASES))
    case = HETERO_CASES[name]
    node_types, edge_types, rowptr_d, col_d, seed_d, nn_d = build_hetero(case)
    types = list(seed_d.keys()) + [t for t in node_types if t not in seed_d]
    tix = {t: i for i, t in enumerate(types)}
    T, R, L = len(types), len(edge_types), len(next(iter(nn_d.values())))
    rels = (REL * R)()
    fan = (C.c_int64 * (R * L))()
    for r, et in enumerate(edge_types):
        k = '__'.join(et)
        src, dst = (et[2], et[0]) if case.get('csc', False) else (et[0], et[2])
        rels[r] = REL(None, None, rowptr_d[k].numel() - 1, col_d[k].numel(), tix[src], tix[dst])
        for h in range(L):
            fan[r * L + h] = nn_d[k][h]
    n_seeds = (C.c_int64 * T)(*[seed_d[t].numel() if t in seed_d else 0 for t in types])
    ncaps, ecaps = (C.c_int64 * T)(), (C.c_int64 * R)()
    if any(v < 0 for v in fan):
        pytest.skip('case uses full neighbourhoods')
    assert lib.pygb200_sampler_bounds(T, R, L, rels, n_seeds, fan, ncaps, ecaps) == 0, lib.pygb200_last_error()
    torch.manual_seed(case['rng_seed'])
    exp = O.hetero_neighbor_sample(node_types, edge_types, rowptr_d, col_d, seed_d, nn_d, csc=case.get('csc', False),
                                   replace=case.get('replace', False), disjoint=case.get('disjoint', False))
    for r, et in enumerate(edge_types):
        assert exp[0]['__'.join(et)].numel() <= ecaps[r]
    for t in node_types:
        assert exp[2][t].shape[0] <= ncaps[tix[t]]


def test_cpu_tensors_raise_clearly():
    """PyG sends CPU tensors to pyg-lib from loader workers; this build has no CPU kernels, so every
    CUDA-key op must fail with one clear message (not a generic dispatcher error, never a silent fallback)."""
    import pytest
    import torch
    import pyg_lib_b200 as P
    rowptr, col, seed = torch.tensor([0, 1, 2]), torch.tensor([1, 0]), torch.tensor([0])
    calls = [
        lambda: P.sampler.neighbor_sample(rowptr, col, seed, [1]),
        lambda: P.sampler.subgraph(rowptr, col, seed),
        lambda: torch.ops.pyg.dist_neighbor_sample(rowptr, col, seed, 1),
        lambda: P.ops.segment_matmul(torch.randn(4, 4), torch.tensor([0, 4]), torch.randn(1, 4, 4)),
        lambda: P.ops.segment_matmul(torch.randn(4, 4, requires_grad=True), torch.tensor([0, 4]), torch.randn(1, 4, 4)),
        lambda: P.ops.segment_matmul(torch.randn(4, 4), torch.tensor([0, 4]), torch.randn(1, 4, 4), bias=torch.randn(1, 4)),
        lambda: P.ops.grouped_matmul([torch.randn(4, 4)], [torch.randn(4, 4)]),
    ]
    for f in calls:
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            f()
    with pytest.raises(RuntimeError, match='no CPU fallback|CUDA tensors'):
        P.sampler.hetero_neighbor_sample({('a', 'to', 'a'): rowptr}, {('a', 'to', 'a'): col}, {'a': seed}, {('a', 'to', 'a'): [1]})
