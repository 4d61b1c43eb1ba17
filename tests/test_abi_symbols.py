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
