"""Short runs of the few-parameter sweeps for ncu: PR_CASE = b9 | ac9 | kg15, PR_MODE = kernel mapping (0 auto, 2 teams, 3 rows)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
from gptpinn import sweep
from bench import kg_grid

case = os.environ.get("PR_CASE", "b9")
mode = int(os.environ.get("PR_MODE", "3"))
E = int(os.environ.get("PR_EPOCHS", "100"))
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda r, n: torch.randn(r, n, device="cuda", generator=g) * 0.3
vec = lambda r: torch.randn(r, 1, device="cuda", generator=g) * 0.3
zer = lambda r: torch.zeros(r, 1, device="cuda")
cfg = dict(mode=mode) if mode else None
if case == "b9":
    n, NR, NI, NB = 9, 10000, 200, 400
    m = [mk(NR, n) for _ in range(4)] + [mk(NI, n), mk(NB, n)]
    grid = np.linspace(0.005, 1, 129)[:128]
    c0 = torch.full((128, n), 1.0 / n, device="cuda")
    run = lambda: sweep.b_sweep(grid, c0, *m, zer(NR), vec(NI), zer(NB), E, 0.001, want_c=False, cfg=cfg)
elif case == "ac9":
    n, NR, NI, NB = 9, 20000, 512, 100
    m = [mk(NR, n) for _ in range(3)] + [mk(NI, n)] + [mk(NB, n) for _ in range(6)]
    lm, ep = np.linspace(1e-3, 1e-4, 11), np.linspace(1, 5, 11)
    grid = np.array(np.meshgrid(lm, ep)).T.reshape(-1, 2)[:120]
    c0 = torch.full((n,), 1.0 / n, device="cuda")
    run = lambda: sweep.ac_sweep(grid, c0, *m, zer(NR), vec(NI), E, 0.0001, want_c=False, cfg=cfg)
else:
    n, NR, NI, NB = 15, 10000, 512, 1024
    m = [mk(NR, n), mk(NR, n), mk(NR, n), mk(NI, n), mk(NI, n), mk(NB, n)]
    grid = kg_grid(10, 10, 10)
    c0 = torch.full((n,), 1.0 / n, device="cuda")
    run = lambda: sweep.kg_sweep(grid, c0, *m, vec(NR), zer(NR), vec(NI), zer(NI), vec(NB), E, 0.001, want_c=False, cfg=cfg)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = run(); e1.record(); torch.cuda.synchronize()
    print(case, "mode", mode, E, "epochs:", e0.elapsed_time(e1), "ms", r["info"], flush=True)
