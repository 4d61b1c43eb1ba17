import sys, os, json, torch, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import pyg_lib_b200 as P
from graphs import ragged_ptr
dev='cuda:0'
abi = C.CDLL(os.path.join(os.path.dirname(P.__file__), 'libpyg_b200.so'))
N,K,M,B = 1<<20,128,128,64
g = torch.Generator().manual_seed(0)
x = torch.randn(N,K,generator=g).to(torch.bfloat16).to(dev)
w = (torch.randn(B,K,M,generator=g)/K**0.5).to(torch.bfloat16).to(dev)
res = {}
for name, ptr in (('ragged', ragged_ptr(N,B,100).to(dev)), ('uniform', torch.arange(0,N+1,N//B).to(dev))):
    for _ in range(5): y = P.ops.segment_matmul(x, ptr, w)
    torch.cuda.synchronize()
    abi.pygb200_profile_enable(1)
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = int(sys.argv[1]) if len(sys.argv)>1 else 50
    e0.record()
    for _ in range(it): y = P.ops.segment_matmul(x, ptr, w)
    e1.record(); torch.cuda.synchronize()
    abi.pygb200_profile_enable(0)
    ms = e0.elapsed_time(e1)/it
    kms, kl, kw = C.c_double(), C.c_int64(), C.c_int64()
    abi.pygb200_profile_read(b'segment_matmul', C.byref(kms), C.byref(kl), C.byref(kw))
    byts = N*K*2+N*M*2+B*K*M*2+(B+1)*8
    res[name] = dict(op_ms=ms, kernel_ms=kms.value/max(kl.value,1), launches=kl.value, tflops=2*N*K*M/(kms.value/max(kl.value,1)*1e-3)/1e12 if kl.value else None, gbs=byts/(kms.value/max(kl.value,1)*1e-3)/1e9 if kl.value else None)
print(json.dumps(res))
