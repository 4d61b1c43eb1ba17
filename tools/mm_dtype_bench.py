import os, sys, json, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import pyg_lib_b200 as P
from graphs import ragged_ptr
dev='cuda:0'
N,K,M,B = 1<<20,128,128,64
g = torch.Generator().manual_seed(0)
res={}
for dt, prec in ((torch.float32, 'highest'), (torch.float32, 'high'), (torch.bfloat16, 'highest')):
    torch.set_float32_matmul_precision(prec)
    x = torch.randn(N,K,generator=g).to(dt).to(dev); w = (torch.randn(B,K,M,generator=g)/K**0.5).to(dt).to(dev)
    ptr = ragged_ptr(N,B,100).to(dev)
    x.requires_grad_(); w.requires_grad_()
    def fwd(): return P.ops.segment_matmul(x, ptr, w)
    for _ in range(3): y = fwd()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = fwd()
    e1.record(); torch.cuda.synchronize(); f_ms = e0.elapsed_time(e1)/10
    gy = torch.randn_like(y)
    for _ in range(2): y = fwd(); y.backward(gy)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): y = fwd(); y.backward(gy)
    e1.record(); torch.cuda.synchronize(); fb_ms = e0.elapsed_time(e1)/5
    # torch reference: per-segment matmul loop
    pl = ptr.tolist()
    def ref(): return torch.cat([x[pl[i]:pl[i+1]] @ w[i] for i in range(B)])
    for _ in range(2): r = ref()
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): r = ref()
    e1.record(); torch.cuda.synchronize(); ref_ms = e0.elapsed_time(e1)/5
    res[str(dt) + '/' + prec] = dict(fwd_ms=f_ms, fwd_tflops=2*N*K*M/f_ms/1e9, fwd_bwd_ms=fb_ms, torch_loop_fwd_ms=ref_ms)
print(json.dumps(res))
