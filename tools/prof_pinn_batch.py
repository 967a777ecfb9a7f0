"""Short batched run of the fused trainer (8 trainings per launch) for ncu."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
import KG_data, KG_models, KG_precomp
from gptpinn import pinn
xt, f_hat, _ = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, 100, 40)
IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, 512, 512)
xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = (x.cuda() for x in (xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u))
xcos = KG_precomp.xcos_term(xt[:, 0:1], xt[:, 1:2])
mus = [(-1.5 + 0.05 * k, 0.5, 0.5) for k in range(8)]
nets = [KG_models.NN(np.array([2, 40, 40, 1]), *mu, xcos).cuda() for mu in mus]
E = int(os.environ.get("EPOCHS", 40))
pinn.train_fused_batch(nets, "kg", mus, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, E, 5e-4, xcos=xcos)
torch.cuda.synchronize()
print("done")
