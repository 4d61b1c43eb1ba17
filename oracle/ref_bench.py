"""CPU timing of the reference's own implementation (oracle/_ref/libpyg_ref.so, falling back to the C
oracle port) on the bench workloads.  TEST/BENCH INFRASTRUCTURE ONLY — run as a subprocess by bench.py
(`cpu_baseline` leg and `--impl reference`), never imported by the product.

    python oracle/ref_bench.py sampler  --workers P --calls C [--graph products|small]
    python oracle/ref_bench.py matmul   --calls C

Prints one JSON object.  The sampler graph is regenerated here from the same seeds bench.py uses
(tests/graphs.py: lognormal_csr), so nothing has to be shipped between processes.
"""
import argparse
import json
import multiprocessing as mp
import os
import os.path as osp
import sys
import time

import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, osp.join(ROOT, 'tests'))

_REF = osp.join(ROOT, 'oracle', '_ref', 'libpyg_ref.so')
KIND = 'port'
if osp.exists(_REF):
    try:
        torch.ops.load_library(_REF)
        KIND = 'reference'
    except Exception:  # noqa
        KIND = 'port'
if KIND == 'port':
    from oracle import oracle as O


def sample_call(rowptr, col, seed, nn, weight=None):
    if KIND == 'reference':
        return torch.ops.pyg.neighbor_sample(rowptr, col, seed, nn, None, None, None, weight, False, False, True, False,
                                             'uniform', True)
    if weight is not None:
        from oracle import weighted as WO
        return WO.neighbor_sample(rowptr, col, seed, nn, weight)
    return O.neighbor_sample(rowptr, col, seed, nn)


def _worker(args):
    wid, rowptr, col, perm, batch, nn, calls, n_workers, weight = args
    torch.set_num_threads(1)
    torch.manual_seed(12345 + wid)
    n_batches = perm.numel() // batch
    edges = 0
    sample_call(rowptr, col, perm[:batch], nn, weight)  # warm-up
    t0 = time.perf_counter()
    for i in range(calls):
        b = (wid + i * n_workers) % n_batches
        out = sample_call(rowptr, col, perm[b * batch:(b + 1) * batch], nn, weight)
        edges += out[0].numel()
    return edges, time.perf_counter() - t0


def papers_shaped_csr(n: int, e: int):
    """Same recipe as tests/graphs.py lognormal_csr (log-normal degrees scaled to sum e, uniform targets), with the 12.9 GB
    of targets drawn in parallel chunks (one CPU generator per chunk): a timing workload of BASELINE configs[4]'s shape —
    not bit-identical to the device-generated graph of our arm, which no CPU generator can reproduce."""
    from concurrent.futures import ThreadPoolExecutor
    g = torch.Generator().manual_seed(1)
    w = torch.empty(n, dtype=torch.float64).log_normal_(3.0, 1.2, generator=g)
    deg = torch.floor(w * (e / w.sum())).to(torch.int64)
    del w
    rem = int(e - int(deg.sum()))
    if rem > 0:
        deg[:rem] += 1
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(deg, 0, out=rowptr[1:])
    del deg
    col = torch.empty(e, dtype=torch.int64)
    chunk = 1 << 25

    def fill(i):
        gi = torch.Generator().manual_seed(1000 + i)
        lo = i * chunk
        col[lo:min(lo + chunk, e)].random_(0, n, generator=gi)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(fill, range((e + chunk - 1) // chunk)))
    return rowptr, col


def bench_sampler(workers: int, calls: int, graph: str, batch: int, nn, weighted: bool = False):
    from graphs import lognormal_csr
    if graph == 'products':
        n, e = 2_449_029, 123_718_280
    elif graph == 'papers':
        n, e = 111_059_956, 1_615_685_872
    else:
        n, e = 200_000, 10_000_000
    if graph == 'papers':
        rowptr, col = papers_shaped_csr(n, e)
    else:
        rowptr, col = lognormal_csr(n, e, seed=1)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2))
    rowptr.share_memory_(); col.share_memory_(); perm.share_memory_()
    weight = None
    if weighted:   # biased sampling (edge_weight): uniform float32 weights
        weight = torch.rand(col.numel(), generator=torch.Generator().manual_seed(3))
        weight.share_memory_()
    jobs = [(w, rowptr, col, perm, batch, nn, calls, workers, weight) for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        res = [_worker(jobs[0])]
    else:
        with mp.get_context('fork').Pool(workers) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    edges = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return dict(kind=KIND, cores=workers, edges=edges, seconds=busy, wall_seconds=wall,
                edges_per_s=edges / busy, calls=calls * workers,
                sample=f'{calls * workers} {"biased " if weighted else ""}calls of {batch} seeds {nn} on the {graph}-shaped CSR '
                       f'({workers} single-threaded worker process(es), disjoint seed batches)')


def bench_matmul(calls: int, threads: int):
    from graphs import ragged_ptr
    torch.set_num_threads(threads)
    N, K, M, B = 1 << 20, 128, 128, 64
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(B, K, M, generator=g) / K ** 0.5).to(torch.bfloat16)
    ptr = ragged_ptr(N, B, 100)
    if KIND == 'reference':
        f = lambda: torch.ops.pyg.segment_matmul(x, ptr, w)  # noqa
    else:
        f = lambda: O.segment_matmul(x, ptr, w)  # noqa
    f()
    t0 = time.perf_counter()
    for _ in range(calls):
        f()
    dt = (time.perf_counter() - t0) / calls
    return dict(kind=KIND, cores=threads, seconds_per_call=dt, tflops=2.0 * N * K * M / dt / 1e12,
                sample=f'{calls} calls, 64 ragged segments, N=2^20, 128->128 bf16, {threads} threads')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['sampler', 'matmul'])
    ap.add_argument('--workers', type=int, default=1)
    ap.add_argument('--calls', type=int, default=100)
    ap.add_argument('--graph', default='products')
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--fanout', default='15,10')
    ap.add_argument('--weighted', action='store_true', help='biased sampling with uniform float32 edge weights')
    a = ap.parse_args()
    if a.what == 'sampler':
        r = bench_sampler(a.workers, a.calls, a.graph, a.batch, [int(v) for v in a.fanout.split(',')], a.weighted)
    else:
        r = bench_matmul(a.calls, a.workers if a.workers > 1 else (os.cpu_count() or 1))
    print('REFBENCH ' + json.dumps(r))
