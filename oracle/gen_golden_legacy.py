"""TEST INFRASTRUCTURE.  Fixture for the legacy (`original/`) Klein-Gordon call surface.

Runs the UNMODIFIED legacy reference (/root/reference/Klein-Gordon/original: KG_GPT_train.gpt_train, KG_GPT_PINN.GPT,
KG_GPT_precomp helpers) on CPU over the activation matrices of tests/golden/kg_small.npz, parameter by parameter as
original/KG_main.py:186-207 does, and records after every call the returned (largest_loss, largest_case) and the
coefficients left in GPT.linears[1].weight; plus the testing mode for a few parameters.
    python oracle/gen_golden_legacy.py   ->  tests/golden/legacy_kg.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GPTPINN_REFERENCE", "/root/reference") + "/Klein-Gordon/original"
sys.path.insert(0, REF)
import KG_GPT_PINN, KG_GPT_optimizer, KG_GPT_precomp, KG_GPT_train   # noqa: E401,E402  (legacy reference modules)

for mod in (KG_GPT_PINN, KG_GPT_optimizer, KG_GPT_precomp):
    mod.device = torch.device("cpu")
g = np.load(os.path.join(ROOT, "tests", "golden", "kg_small.npz"))
t = {k: torch.from_numpy(np.ascontiguousarray(g[k])) for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat",
                                                                "IC_u1", "IC_u2", "BC_u", "xt_resid", "IC_xt", "BC_xt")}
rng = np.random.default_rng(3)
t["f_hat"] = torch.from_numpy((0.05 * rng.standard_normal(g["f_hat"].shape)).astype(np.float32))   # non-zero: enters the loss only
grid, E, lr = g["grid"], 120, 0.025
out = {"grid": grid, "epochs": E, "lr": lr, "f_hat": t["f_hat"].numpy()}
for n in (2, 7):
    layers = np.array([2, n, 1])
    c_initial = torch.full((1, n), 1 / n)
    largest_loss, largest_case = 0, 0
    Ls, cases, cs = [], [], []
    sl = {k: t[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
    for mu in grid:
        a, b, gm = mu
        T = KG_GPT_precomp.Ptt_aPxx_bP(a, b, sl["out_tt"], sl["out_xx"], sl["out"])
        G2 = KG_GPT_precomp.gamma2_P(gm, sl["out"])
        net = KG_GPT_PINN.GPT(layers, a, b, gm, None, c_initial.clone(), t["IC_u1"], t["IC_u2"], t["BC_u"], t["f_hat"], t["xcos"],
                              sl["out"], sl["out_IC"], sl["out_BC"], sl["out_IC_t"], T)
        largest_loss, largest_case = KG_GPT_train.gpt_train(net, a, b, gm, t["xt_resid"], t["IC_xt"], t["IC_u1"], t["IC_u2"], t["BC_xt"], t["BC_u"],
                                                            t["xcos"], T, G2, sl["out"], sl["out_IC"], sl["out_BC"], sl["out_IC_t"], E, lr,
                                                            largest_loss, largest_case)
        Ls.append(float(largest_loss))
        cases.append(-1 if isinstance(largest_case, int) else int(np.nonzero((grid == np.asarray(largest_case)).all(1))[0][0]))
        cs.append(net.linears[1].weight.data.view(-1).numpy().copy())
    out[f"n{n}_L"], out[f"n{n}_case"], out[f"n{n}_c"] = np.array(Ls), np.array(cases), np.array(cs)
    tc = []
    for mu in grid[:5]:
        a, b, gm = mu
        T = KG_GPT_precomp.Ptt_aPxx_bP(a, b, sl["out_tt"], sl["out_xx"], sl["out"])
        G2 = KG_GPT_precomp.gamma2_P(gm, sl["out"])
        net = KG_GPT_PINN.GPT(layers, a, b, gm, None, c_initial.clone(), t["IC_u1"], t["IC_u2"], t["BC_u"], t["f_hat"], t["xcos"],
                              sl["out"], sl["out_IC"], sl["out_BC"], sl["out_IC_t"], T)
        KG_GPT_train.gpt_train(net, a, b, gm, t["xt_resid"], t["IC_xt"], t["IC_u1"], t["IC_u2"], t["BC_xt"], t["BC_u"], t["xcos"], T, G2,
                               sl["out"], sl["out_IC"], sl["out_BC"], sl["out_IC_t"], E, lr, testing=True)
        tc.append(net.linears[1].weight.data.view(-1).numpy().copy())
        if n == 7 and len(tc) == 1:
            out["n7_loss_after_test"] = float(net.loss())
            GD = KG_GPT_optimizer.grad_descent(a, b, gm, t["xt_resid"], t["IC_xt"], t["BC_xt"], t["IC_u1"], t["IC_u2"], t["BC_u"], t["xcos"], T, G2,
                                               sl["out"], sl["out_IC"], sl["out_BC"], sl["out_IC_t"], E, lr)
            out["n7_one_update"] = GD.update(c_initial.view(-1).clone()).numpy().copy()
    out[f"n{n}_test_c"] = np.array(tc)
    print(n, "exits:", sum(1 for i in range(1, len(Ls)) if Ls[i] == Ls[i - 1]), "of", len(Ls), "final", Ls[-1], cases[-1])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "legacy_kg.npz"), **out)
print("wrote tests/golden/legacy_kg.npz")
