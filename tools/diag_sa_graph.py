"""Diagnostic: Adam iteration of the SA-PINN trainer with the fused map kernels vs the library-GEMM composition, as a CUDA
graph and eagerly, alternating in one process (device time by events, wall time beside it)."""
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
    m.adam_fit(20)
    n = int(os.environ.get("DIAG_ITERS", 150))
    for name, fused, graph in (("fused_graph", True, True), ("composed_graph", False, True), ("fused_graph_again", True, True),
                               ("fused_eager", True, False), ("composed_eager", False, False), ("fused_graph_third", True, True)):
        if fused:
            os.environ.pop("GPTPINN_SA_NO_FUSED_MAPS", None)
        else:
            os.environ["GPTPINN_SA_NO_FUSED_MAPS"] = "1"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
        m.adam_fit(n, use_graph=graph)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({name: dict(wall_ms_per_iter=1e3 * (time.perf_counter() - t0) / n, device_ms_per_iter=e0.elapsed_time(e1) / n)}), flush=True)


if __name__ == "__main__":
    main()
