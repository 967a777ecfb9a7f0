import os, sys, time
import numpy as np, torch
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
import KG_data, KG_models, KG_precomp
from gptpinn import pinn
for nc in (20, 50, 88, 100, 130, 160):
    xt, f_hat, _ = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, nc, 40)
    IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, 512, 512)
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = (x.cuda() for x in (xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u))
    xcos = KG_precomp.xcos_term(xt[:, 0:1], xt[:, 1:2])
    net = KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, xcos).cuda()
    a = (net, "kg", (-1.5, 0.5, 0.5), xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u)
    pinn.train_fused(*a, 10, 5e-4, xcos=xcos); torch.cuda.synchronize()
    t0 = time.perf_counter(); pinn.train_fused(*a, 2000, 5e-4, xcos=xcos); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tiles = (xt.shape[0] + 31)//32 + 16 + 32
    print(f"NR={xt.shape[0]:6d} tiles={tiles:4d}  {1e6*dt/2000:7.1f} us/epoch")
