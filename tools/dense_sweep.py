"""BASELINE config 5: dense-collocation Klein-Gordon (N_R = 10^6 residual rows), the GPT-PINN test sweep of KG_test.py
(200 test parameters drawn from the default 10x10x10 grid, n = 15).  Times a short run and reports param*epochs/s."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpt-pinn_b200"))
from gptpinn import sweep  # noqa: E402


def main():
    NR = int(os.environ.get("DS_NR", 1_000_000))
    epochs = int(os.environ.get("DS_EPOCHS", 40))
    NI, NB, n = 512, 1024, 15
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda r: (torch.randn(r, 15, device="cuda", generator=g) * 0.3)
    out, out_xx, out_tt = mk(NR), mk(NR), mk(NR)
    out_IC, out_IC_t, out_BC = mk(NI), mk(NI), mk(NB)
    xcos = torch.randn(NR, 1, device="cuda", generator=g) * 0.3
    f_hat = torch.zeros(NR, 1, device="cuda")
    IC_u1 = torch.randn(NI, 1, device="cuda", generator=g)
    IC_u2 = torch.zeros(NI, 1, device="cuda")
    BC_u = torch.randn(NB, 1, device="cuda", generator=g)
    a = np.linspace(-2, -1, 10); b = np.linspace(0, 1, 10); gm = np.linspace(0, 1, 10)
    kg_grid = np.array(np.meshgrid(a, b, gm)).T.reshape(-1, 3)
    test = kg_grid[np.random.default_rng(1234).choice(1000, 200, replace=False)]
    c0 = torch.full((n,), 1.0 / n, device="cuda")
    W_n = 8 * n * NR + 4 * n * (NB + 2 * NI)
    cfgs = [None] + [dict(zip(("M", "SL", "W", "mode", "K"), (int(x) for x in c.split(":")))) for c in os.environ.get("DS_CFGS", "").split(",") if c]
    for cfg in cfgs:
        def run():
            return sweep.kg_sweep(test, c0, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, xcos, f_hat, IC_u1, IC_u2, BC_u,
                                  epochs, 0.001, want_c=False, cfg=cfg)
        r = run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        pe = len(test) * epochs / (ms * 1e-3)
        stream_mb = r["info"]["tiles_per_pass"] * (3 * 16 * (32 if r["info"]["priv"] else 64) + 2 * (32 if r["info"]["priv"] else 64)) * 4 / 1e6
        print(json.dumps(dict(cfg=cfg, info=r["info"], ms=ms, param_epochs_per_s=pe, tflops_alg=pe * W_n / 1e12, stream_MB=stream_mb,
                              est_5000_epochs_s=5000 * len(test) / pe)), flush=True)


if __name__ == "__main__":
    main()
