"""Small invocations of every kernel, meant to be run under compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from conftest import load_golden, lr_of  # noqa: E402
from gptpinn import pinn, precomp, sweep  # noqa: E402

g = load_golden("kg_small")
KG = ["out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat", "IC_u1", "IC_u2", "BC_u"]
d = {k: torch.tensor(g[k], device="cuda") for k in KG}
n = 7
c0 = torch.full((n,), 1.0 / n, device="cuda")
ref = None
for cfg in (None, dict(mode=1), dict(mode=2, K=1), dict(mode=2, K=2, M=5), dict(mode=2, K=8, M=2), dict(M=1, SL=32, mode=2, K=4)):
    r = sweep.kg_sweep(g["grid"], c0, d["out"][:, :n], d["out_xx"][:, :n], d["out_tt"][:, :n], d["out_IC"][:, :n], d["out_IC_t"][:, :n],
                       d["out_BC"][:, :n], d["xcos"], d["f_hat"], d["IC_u1"], d["IC_u2"], d["BC_u"], 5, lr_of(g, 0), cfg=cfg)
    f = r["f"].cpu().numpy()
    ref = f if ref is None else ref
    print("kg", cfg, r["info"]["priv"], r["info"]["K"], float(np.max(np.abs(f - ref) / np.abs(ref))))
fh = torch.randn_like(d["f_hat"]) * 0.1                     # non-zero f_hat: the general path
r = sweep.kg_sweep(g["grid"], c0, d["out"][:, :n], d["out_xx"][:, :n], d["out_tt"][:, :n], d["out_IC"][:, :n], d["out_IC_t"][:, :n],
                   d["out_BC"][:, :n], d["xcos"], fh, d["IC_u1"], d["IC_u2"], d["BC_u"], 5, lr_of(g, 0))
print("kg f_hat", float(r["f"].sum()))
L, idx = sweep.argmax_scan(r["f"], r["m"])
print("argmax", float(L), idx)

import KG_data, KG_models, KG_precomp  # noqa: E402
xt, f_hat, _ = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, 12, 4)
IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, 20, 20)
xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = (x.cuda() for x in (xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u))
xcos = KG_precomp.xcos_term(xt[:, 0:1], xt[:, 1:2])
nets = [KG_models.NN(np.array([2, 40, 40, 1]), -1.5 + 0.1 * k, 0.5, 0.5, xcos).cuda() for k in range(3)]
res = pinn.train_fused_batch(nets, "kg", [(-1.5 + 0.1 * k, 0.5, 0.5) for k in range(3)], xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u,
                             4, 5e-4, xcos=xcos, trace_every=1)
print("pinn batch", [float(r["loss"]) for r in res])
out = torch.zeros((xt.shape[0], 3), device="cuda")
precomp.mlp_channels(nets[0], "cos", xt, v=out[:, 0], q=out[:, 1], r=out[:, 2])
print("precompute", float(out.abs().sum()))
torch.cuda.synchronize()
print("done")
