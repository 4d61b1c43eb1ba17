"""Frontier-sharded sampling benchmark (BASELINE configs[4] shape): one batch of --seeds seeds, fan-out
[15,10], CSR replicated on every GPU, edge ids all-gathered over NCCL per hop.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools_dist_bench.py --graph papers
Prints one JSON line from rank 0: single-GPU op time vs sharded time, and checks that the sharded result
is identical to the single-GPU result on every rank."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graph', default='products')
    ap.add_argument('--seeds', type=int, default=65536)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--collective', action='store_true', help='also time the NCCL-broadcast transport with replicated dedup')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
    import pyg_lib_b200 as P
    from graphs import lognormal_csr
    n, e = (2_449_029, 123_718_280) if a.graph == 'products' else (111_059_956, 1_615_685_872)
    t0 = time.time()
    rowptr, col = lognormal_csr(n, e, seed=1, device=dev)   # same CUDA generator seed on every rank -> identical graph
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    seed = torch.randperm(n, generator=torch.Generator().manual_seed(2))[:a.seeds].to(dev)
    nn = [15, 10]

    def timed(fn):
        torch.manual_seed(12345)
        for _ in range(3):
            out = fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        edges = 0
        for _ in range(a.iters):
            out = fn()
            edges += out[0].numel()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / a.iters, edges / a.iters, out

    ms1, ed1, _ = timed(lambda: P.sampler.neighbor_sample(rowptr, col, seed, nn))
    msN, edN, _ = timed(lambda: P.sampler.dist_neighbor_sample(rowptr, col, seed, nn))
    msC, _, _ = timed(lambda: P.sampler.dist_neighbor_sample(rowptr, col, seed, nn, transport='collective')) if a.collective else (None, None, None)
    # identical results: same seed, one call each
    torch.manual_seed(7); o1 = P.sampler.neighbor_sample(rowptr, col, seed, nn)
    torch.manual_seed(7); oN = P.sampler.dist_neighbor_sample(rowptr, col, seed, nn)
    same = all(torch.equal(x, y) for x, y in zip(o1[:4], oN[:4])) and o1[4] == oN[4] and o1[5] == oN[5]
    chk = torch.tensor([int(oN[0].sum() % (1 << 40)), int(oN[1].sum() % (1 << 40)), int(oN[2].sum() % (1 << 40))], device=dev)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'graph': a.graph, 'n_gpus': world, 'seeds': a.seeds, 'fanout': nn, 'graph_gen_s': gen_s,
                          'single_gpu_ms': ms1, 'single_gpu_edges_per_s': ed1 / (ms1 * 1e-3), 'sharded_ms': msN, 'sharded_collective_ms': msC,
                          'sharded_edges_per_s': edN / (msN * 1e-3), 'edges_per_call': edN,
                          'sharded_equals_single_gpu': bool(same), 'all_ranks_identical': bool(torch.equal(lo, hi))}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
