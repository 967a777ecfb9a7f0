import os, sys, time
import numpy as np, torch
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
from gptpinn import precomp
g = torch.Generator().manual_seed(0)
layers=[2,128,128,128,128,1]
W=[(torch.randn(layers[i+1],layers[i],generator=g)*(1.0/np.sqrt(layers[i]))).cuda() for i in range(5)]
B=[torch.zeros(layers[i+1]).cuda() for i in range(5)]
net=list(zip(W,B))
for N in (20000, 200000):
    xt=torch.rand(N,2,generator=g).cuda()
    out=torch.zeros(N,3,device='cuda')
    for rep in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); precomp.mlp_channels(net,"tanh",xt,v=out[:,0],t=out[:,1],q=out[:,2]); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)
    fl=N*(2*128+3*3*128*128+3*128)*2
    print(f"AC precompute N={N}: {ms:.3f} ms  ({fl/ms/1e9:.1f} GFLOP/s algorithmic)")
