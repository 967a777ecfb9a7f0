"""SA-PINN trainer timings at the reference sizes (AC_main.py: [2,128,128,128,128,1], 20 000 + 512 + 2 x 100 points): ms per
Adam iteration (graph replay), ms per L-BFGS iteration with the one-launch direction kernel and with tensor expressions, and
the direction kernel alone for a full 50-pair history."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpt-pinn_b200"))
from gptpinn.sa_pinn import SAPinn  # noqa: E402


def main():
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    NR, NI, NB = 20000, 512, 100
    xt_f = torch.stack((2 * torch.rand(NR, generator=g) - 1, torch.rand(NR, generator=g)), 1).to(dev)
    x0 = torch.linspace(-1, 1, NI)
    xt0 = torch.stack((x0, torch.zeros(NI)), 1).to(dev)
    u0 = (x0 ** 2 * torch.cos(np.pi * x0)).to(dev)
    tb = torch.rand(NB, generator=g)
    xt_lb, xt_ub = torch.stack((-torch.ones(NB), tb), 1).to(dev), torch.stack((torch.ones(NB), tb), 1).to(dev)
    m = SAPinn([2, 128, 128, 128, 128, 1], xt_f, xt0, u0, xt_lb, xt_ub, 1e-4, 5.0, seed=1)
    out = {}
    n_adam = int(os.environ.get("SA_ADAM", 600))
    m.adam_fit(50)
    if n_adam > 0:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.adam_fit(n_adam)
        torch.cuda.synchronize(); out["adam_ms_per_iter"] = 1e3 * (time.perf_counter() - t0) / n_adam
    th = m.theta.clone()
    n_lb = int(os.environ.get("SA_LBFGS", 400))
    for name, fused in (("lbfgs_fused_ms_per_iter", True), ("lbfgs_tensor_expr_ms_per_iter", False)):
        if n_lb <= 0 or (not fused and os.environ.get("SA_SKIP_UNFUSED")):
            continue
        m.theta.copy_(th)
        m.lbfgs_fit(20, 0.8, use_fused=fused)
        m.theta.copy_(th)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fh = m.lbfgs_fit(n_lb, 0.8, use_fused=fused)
        torch.cuda.synchronize(); out[name] = 1e3 * (time.perf_counter() - t0) / max(len(fh) - 1, 1)
        out[name.replace("ms_per_iter", "iters")] = len(fh) - 1
        out[name.replace("ms_per_iter", "final_loss")] = fh[-1]
    # the direction kernel alone, full history
    P, hist = m.P, 50
    ld = (P + 3) & ~3
    S = (torch.randn(hist, ld, device=dev) * 1e-2)[:, :P]
    Y = (torch.randn(hist, ld, device=dev) * 1e-2)[:, :P]
    ro = torch.rand(hist, device=dev)
    gg, d = torch.randn(P, device=dev), torch.empty(P, device=dev)
    for _ in range(3):
        m.ops.lbfgs_direction(S, Y, ro, hist, 7, 0.5, gg, d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.ops.lbfgs_direction(S, Y, ro, hist, 7, 0.5, gg, d)
    e1.record(); torch.cuda.synchronize()
    out["direction_kernel_us_k50"] = 1e3 * e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20):
        m.loss_and_grad()
    e1.record(); torch.cuda.synchronize()
    out["loss_and_grad_ms"] = e0.elapsed_time(e1) / 20
    print(json.dumps(out))
    if os.environ.get("SA_PROFILE"):            # kernel-level split of one loss / gradient evaluation
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                m.loss_and_grad()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))


if __name__ == "__main__":
    main()
