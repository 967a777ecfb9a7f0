"""Quick device-side throughput probe of the KG sweep for several (M, SL, W) mappings."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpt-pinn_b200"))
from gptpinn import _lib, sweep  # noqa: E402
import ctypes as C  # noqa: E402


def main():
    n = int(os.environ.get("QB_N", 15))
    epochs = int(os.environ.get("QB_EPOCHS", 40))
    na, nb, ng = (int(x) for x in os.environ.get("QB_GRID", "50,50,40").split(","))
    NR, NI, NB = 10000, 512, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda r: (torch.randn(r, 15, device="cuda", generator=g) * 0.3)
    out, out_xx, out_tt = mk(NR), mk(NR), mk(NR)
    out_IC, out_IC_t, out_BC = mk(NI), mk(NI), mk(NB)
    xcos = torch.randn(NR, 1, device="cuda", generator=g) * 0.3
    f_hat = torch.zeros(NR, 1, device="cuda")
    IC_u1 = torch.randn(NI, 1, device="cuda", generator=g)
    IC_u2 = torch.zeros(NI, 1, device="cuda")
    BC_u = torch.randn(NB, 1, device="cuda", generator=g)
    a = np.linspace(-2, -1, na); b = np.linspace(0, 1, nb); gm = np.linspace(0, 1, ng)
    grid = np.array(np.meshgrid(a, b, gm)).T.reshape(-1, 3)
    if os.environ.get("QB_SHARD"):          # "r/W": rank r's block of the (alpha, beta)-sorted grid, as offline.sharded_sweep cuts it
        r_, w_ = (int(x) for x in os.environ["QB_SHARD"].split("/"))
        order = np.lexsort((grid[:, 1], grid[:, 0]))
        base, rem = divmod(len(grid), w_)
        lo = r_ * base + min(r_, rem)
        grid = grid[order[lo:lo + base + (1 if r_ < rem else 0)]]
    c0 = torch.full((n,), 1.0 / n, device="cuda")
    tf = C.c_double()
    _lib.check(_lib.lib().gptpinn_ffma_peak(C.byref(tf), 3, None), "peak")
    print("FFMA peak measured TFLOP/s:", tf.value, flush=True)
    W_n = 8 * n * NR + 4 * n * (NB + 2 * NI)
    def parse(c):
        v = [int(x) for x in c.split(":")] + [0] * 5
        return dict(M=v[0], SL=v[1], W=v[2], mode=v[3], K=v[4])
    cfgs = [None] + [parse(c) for c in os.environ.get("QB_CFGS", "5:8:8:1,5:8:8:2,5:0:8:2,5:4:8:2,4:8:8:2,2:0:16:2").split(",")]
    for cfg in cfgs:
        def run():
            return sweep.kg_sweep(grid, c0, out[:, :n], out_xx[:, :n], out_tt[:, :n], out_IC[:, :n], out_IC_t[:, :n],
                                  out_BC[:, :n], xcos, f_hat, IC_u1, IC_u2, BC_u, epochs, 0.001, want_c=False, cfg=cfg)
        r = run(); r = run(); torch.cuda.synchronize()          # two warm-ups: module load + clock ramp
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        pe = len(grid) * epochs / (ms * 1e-3)
        print(json.dumps(dict(cfg=cfg, info=r["info"], ms=ms, param_epochs_per_s=pe, tflops_alg=pe * W_n / 1e12,
                              frac_nominal=pe * W_n / 74.5e12, frac_measured=pe * W_n / (tf.value * 1e12))), flush=True)


if __name__ == "__main__":
    main()
