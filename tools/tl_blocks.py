import sys, os, ctypes, json, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import pyg_lib_b200 as P
from graphs import lognormal_csr
import collections
dev = 'cuda:0'
lib = ctypes.CDLL(os.path.join(os.path.dirname(P.__file__), 'libpyg_b200.so'))
lib.pygb200_timeline_read.restype = ctypes.c_int64
rowptr, col = lognormal_csr(2_449_029, 123_718_280, seed=1, device=dev)
seed = torch.randperm(2_449_029, generator=torch.Generator().manual_seed(2))[:1024].to(dev)
torch.manual_seed(12345)
for _ in range(50): out = P.sampler.neighbor_sample(rowptr, col, seed, [15, 10])
torch.cuda.synchronize()
lib.pygb200_timeline_enable(1)
for _ in range(3): out = P.sampler.neighbor_sample(rowptr, col, seed, [15, 10])
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 8192)()
n = lib.pygb200_timeline_read(buf, 8192)
ev = sorted(((buf[i] & ((1 << 48) - 1)), buf[i] >> 56, (buf[i] >> 48) & 0xff) for i in range(n))
# last call: find last 'final.end' (id 6|16=22) before and take events after previous final.end
ends = [i for i, e in enumerate(ev) if e[1] == 22]
a = ends[-2] + 1; b = ends[-1]
evs = ev[a:b + 1]
base = evs[0][0]
# per-kernel: second k_sample (hop 2): block stamps id 3|8=11 (start) and 3|8|16=27 (end)
starts = [(t - base) / 1e3 for t, k, sm in evs if k == 11]
endsb = [(t - base) / 1e3 for t, k, sm in evs if k == 27]
sms = collections.Counter(sm for t, k, sm in evs if k == 11)
main = [((t - base) / 1e3, k) for t, k, sm in evs if k not in (11, 27)]
print('main stamps', main)
print('n block starts', len(starts), 'ends', len(endsb))
import statistics
def hist(x, name):
    x = sorted(x)
    print(name, 'min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f' % (x[0], x[len(x)//10], x[len(x)//2], x[9*len(x)//10], x[-1]))
# split hop1/hop2 by time gap: hop1 blocks < first mark start
marks = [t for t, k in main if k == 4]
if marks:
    h1s = [t for t in starts if t < marks[0]]; h2s = [t for t in starts if t > marks[0]]
    h1e = [t for t in endsb if t < marks[0]]; h2e = [t for t in endsb if t > marks[0]]
    if h1s: hist(h1s, 'hop1 starts'); hist(h1e, 'hop1 ends')
    if h2s: hist(h2s, 'hop2 starts'); hist(h2e, 'hop2 ends')
print('blocks per SM (all sample blocks): max', max(sms.values()), 'distinct SMs', len(sms))
