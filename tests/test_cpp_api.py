"""The C++ callers' entry points of libpyg.so (pyg::ops::segment_matmul / grouped_matmul, pyg::sampler::neighbor_sample /
hetero_neighbor_sample / dist_neighbor_sample — api.h; reference: ops/matmul.cpp:12-60, sampler/neighbor.cpp:11-127):
a small C++ program is compiled against the library and run.  CPU: the argument checks and the dispatcher's
"no CPU fallback" answer.  GPU: real calls, incl. autograd through the same entry point."""
import os
import os.path as osp
import subprocess
import sysconfig

import pytest
import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
EXE = osp.join(ROOT, 'pyg_lib_b200', '_build', 'api_check')


def _build():
    src = osp.join(ROOT, 'tests', 'cpp', 'api_check.cpp')
    lib = osp.join(ROOT, 'pyg_lib_b200', 'libpyg.so')
    if osp.exists(EXE) and osp.getmtime(EXE) > max(osp.getmtime(src), osp.getmtime(lib)):
        return
    os.makedirs(osp.dirname(EXE), exist_ok=True)
    tdir = osp.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ['g++', '-O1', '-std=c++17', '-w', f'-D_GLIBCXX_USE_CXX11_ABI={abi}', src, '-o', EXE,
           '-I' + osp.join(ROOT, 'pyg_lib_b200', 'csrc', 'torch'), '-I' + osp.join(tdir, 'include'),
           '-I' + osp.join(tdir, 'include', 'torch', 'csrc', 'api', 'include'), '-I' + sysconfig.get_paths()['include'],
           '-L' + osp.join(ROOT, 'pyg_lib_b200'), '-lpyg', '-lpyg_b200', '-L' + osp.join(tdir, 'lib'), '-ltorch', '-ltorch_cpu', '-lc10',
           '-Wl,-rpath,' + osp.join(ROOT, 'pyg_lib_b200'), '-Wl,-rpath,' + osp.join(tdir, 'lib'), '-Wl,--no-as-needed', '-ltorch_cuda', '-lc10_cuda']
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def _run(args):
    _build()
    r = subprocess.run([EXE] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'CPP_API_OK' in r.stdout, (r.stdout[-1000:], r.stderr[-2000:])


def test_cpp_api_argument_checks_cpu():
    _run([])


@pytest.mark.gpu
def test_cpp_api_calls_gpu():
    _run(['--gpu'])
