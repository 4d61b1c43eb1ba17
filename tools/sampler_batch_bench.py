import sys, os, json, torch, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import pyg_lib_b200 as P
from graphs import lognormal_csr
dev = 'cuda:0'
S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rowptr, col = lognormal_csr(2_449_029, 123_718_280, seed=1, device=dev)
seed = torch.randperm(2_449_029, generator=torch.Generator().manual_seed(2))[:S].to(dev)
torch.manual_seed(12345)
for _ in range(3): out = P.sampler.neighbor_sample(rowptr, col, seed, [15, 10])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters): out = P.sampler.neighbor_sample(rowptr, col, seed, [15, 10])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / iters
print(json.dumps({'seeds': S, 'ms': dt * 1e3, 'edges': out[0].numel(), 'edges_per_s': out[0].numel() / dt}))
