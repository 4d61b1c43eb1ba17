import torch, sys
import pyg_lib_b200 as P
torch.set_float32_matmul_precision('high')
dev='cuda:0'
for (N,K,M) in [(100,64,32),(128,64,32),(256,64,32),(128,128,128),(100,128,128)]:
    g=torch.Generator().manual_seed(0)
    x=torch.randn(N,K,generator=g); w=torch.randn(1,K,M,generator=g)
    out=P.ops.segment_matmul(x.to(dev), torch.tensor([0,N]).to(dev), w.to(dev)).cpu()
    ref=x@w[0]
    print(N,K,M,'out norm',float(out.norm()),'ref norm',float(ref.norm()),'err',float((out-ref).norm()/ref.norm()), 'nonzero frac', float((out!=0).float().mean()))
    if N<=128 and M==32:
        print(' out[0,:6]',out[0,:6].tolist()); print(' ref[0,:6]',ref[0,:6].tolist())
