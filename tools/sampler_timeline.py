"""Device timeline of one sampler call (debug; pygb200_timeline_enable): per-kernel start/end stamps of block 0
and of the serial last-block sections, relative to the first stamp of the call.
    python tools/sampler_timeline.py [seeds] [calls]"""
import sys, os, ctypes, json, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import pyg_lib_b200 as P
from graphs import lognormal_csr
dev = 'cuda:0'
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = ctypes.CDLL(os.path.join(os.path.dirname(P.__file__), 'libpyg_b200.so'))
lib.pygb200_timeline_read.restype = ctypes.c_int64
rowptr, col = lognormal_csr(2_449_029, 123_718_280, seed=1, device=dev)
seed = torch.randperm(2_449_029, generator=torch.Generator().manual_seed(2))[:S].to(dev)
torch.manual_seed(12345)
for _ in range(50): out = P.sampler.neighbor_sample(rowptr, col, seed, [15, 10])
torch.cuda.synchronize()
lib.pygb200_timeline_enable(1)
t0 = time.perf_counter()
for _ in range(calls): out = P.sampler.neighbor_sample(rowptr, col, seed, [15, 10])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / calls
buf = (ctypes.c_uint64 * 8192)()
n = lib.pygb200_timeline_read(buf, 8192)
lib.pygb200_timeline_enable(0)
names = {1: 'seed', 2: 'count', 3: 'sample', 4: 'mark', 5: 'assign', 6: 'final', 7: 'export'}
ev = sorted(((buf[i] & ((1 << 48) - 1)), buf[i] >> 56, (buf[i] >> 48) & 0xff) for i in range(n))
# split into calls at every 'seed' start
starts = [i for i, (t, k, sm) in enumerate(ev) if k == 1]
print(json.dumps({'seeds': S, 'us_per_call_host': dt * 1e6, 'stamps': n, 'calls_seen': len(starts)}))
if len(starts) >= 3:
    a, b, c = starts[-3], starts[-2], starts[-1]
    base = ev[a][0]
    for t, k, sm in ev[a:c + 1]:
        nm = names.get(k & 7, '?') + ('.last' if k & 8 else '') + ('.end' if k & 16 else '.start')
        print(f'{(t - base) / 1e3:9.2f} us  {nm:18s} sm{sm}')
