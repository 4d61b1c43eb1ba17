"""Multi-process tests of the frontier-sharded sampler.
CPU (gloo, world_size 2): the variable-length in-place all-gather the C-ABI callback relies on.
GPU: world_size 2 through the full sharded path, bit-exact vs the oracle — NCCL over NVLink when the
box has >= 2 GPUs, else two ranks sharing GPU 0 over gloo (exercises the same kernels and callback)."""
import os
import os.path as osp
import socket
import subprocess
import sys

import pytest
import torch

HERE = osp.dirname(osp.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(backend, mode, nproc=2, timeout=600):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), osp.join(HERE, 'dist_worker.py'), backend, mode]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout,
                         env=dict(os.environ, OMP_NUM_THREADS='1'))
    assert out.returncode == 0 and 'DIST_OK' in out.stdout, out.stdout[-4000:]


def test_allgather_segments_gloo_cpu():
    _run('gloo', 'segments')


def test_allgather_segments_gloo_cpu_world3():
    _run('gloo', 'segments', nproc=3)


@pytest.mark.gpu
def test_sharded_sampler_two_ranks():
    backend = 'nccl' if torch.cuda.device_count() >= 2 else 'gloo'
    _run(backend, 'sample')
