"""hetero_neighbor_sample on a MAG240M-shaped graph (BASELINE configs[3]): 3 node types / 6 edge types,
fan-out [25,15] for every relation, 1024 paper seeds.  `--scale` multiplies all node/edge counts
(1.0 = 121.7M papers, 122.4M authors, 25.7k institutions; 1.30G cites, 386M writes, 44.6M affiliated + reverses)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import pyg_lib_b200 as P
from graphs import lognormal_csr

ap = argparse.ArgumentParser(); ap.add_argument('--scale', type=float, default=1.0); ap.add_argument('--iters', type=int, default=50)
a = ap.parse_args()
dev = 'cuda:0'
N = {'paper': int(121_751_666 * a.scale), 'author': int(122_383_112 * a.scale), 'institution': max(int(25_721 * a.scale), 64)}
E = {('paper', 'cites', 'paper'): 1_297_748_926, ('author', 'writes', 'paper'): 386_022_720,
     ('author', 'affiliated_with', 'institution'): 44_592_586, ('paper', 'rev_writes', 'author'): 386_022_720,
     ('institution', 'rev_affiliated_with', 'author'): 44_592_586, ('paper', 'rev_cites', 'paper'): 1_297_748_926}
rowptr_d, col_d, nn_d = {}, {}, {}
t0 = time.time()
for i, (k, e) in enumerate(E.items()):
    e = int(e * a.scale)
    rp, _ = lognormal_csr(N[k[0]], e, seed=10 + i, device=dev) if False else (None, None)
    g = torch.Generator(device=dev).manual_seed(10 + i)
    w = torch.empty(N[k[0]], device=dev, dtype=torch.float32).log_normal_(3.0, 1.2, generator=g)
    deg = torch.floor(w.double() * (e / float(w.double().sum()))).to(torch.int64); del w
    rem = e - int(deg.sum()); deg[:rem] += 1
    rowptr = torch.zeros(N[k[0]] + 1, dtype=torch.int64, device=dev); torch.cumsum(deg, 0, out=rowptr[1:]); del deg
    col = torch.randint(0, N[k[2]], (e,), generator=g, device=dev, dtype=torch.int64)
    rowptr_d[k], col_d[k], nn_d[k] = rowptr, col, [25, 15]
torch.cuda.synchronize()
gen_s = time.time() - t0
perm = torch.randperm(N['paper'], device=dev)
torch.manual_seed(12345)
def call(i):
    return P.sampler.hetero_neighbor_sample(rowptr_d, col_d, {'paper': perm[i * 1024:(i + 1) * 1024]}, nn_d)
for i in range(5): out = call(i)
torch.cuda.synchronize()
t0 = time.perf_counter(); edges = 0
for i in range(a.iters):
    out = call(5 + i); edges += sum(v.numel() for v in out[0].values())
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.iters
print(json.dumps({'scale': a.scale, 'graph_bytes_GB': sum(v.numel() for v in col_d.values()) * 8 / 1e9, 'graph_gen_s': gen_s,
                  'ms_per_call': dt * 1e3, 'edges_per_call': edges / a.iters, 'edges_per_s': edges / a.iters / dt,
                  'kernel_launches_per_call': None}))
