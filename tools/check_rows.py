"""Few-parameter sweeps: every kernel mapping against the team kernel (parity) and per-epoch times.
CR_CASES = comma list of b1,b9,ac1,ac9,kg1,kg8,kg15 ; CR_EPOCHS"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
from gptpinn import sweep
from bench import kg_grid

E = int(os.environ.get("CR_EPOCHS", "500"))
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda r, n: torch.randn(r, n, device="cuda", generator=g) * 0.3
vec = lambda r: torch.randn(r, 1, device="cuda", generator=g) * 0.3
zer = lambda r: torch.zeros(r, 1, device="cuda")


def problem(case):
    pde, n = case.rstrip("0123456789"), int(case.lstrip("abckg"))
    if pde == "b":
        NR, NI, NB = 10000, 200, 400
        m = [mk(NR, n) for _ in range(4)] + [mk(NI, n), mk(NB, n)]
        grid = np.linspace(0.005, 1, 129)[:int(os.environ.get("CR_B_NP", "128"))]
        c0 = torch.full((len(grid), n), 1.0 / n, device="cuda")
        fh, iu, bu = zer(NR), vec(NI), zer(NB)
        W = 12 * n * NR + 4 * n * (NB + NI)
        return len(grid), W, lambda cfg: sweep.b_sweep(grid, c0, *m, fh, iu, bu, E, 0.001, want_c=True, cfg=cfg)
    if pde == "ac":
        NR, NI, NB = 20000, 512, 100
        m = [mk(NR, n) for _ in range(3)] + [mk(NI, n)] + [mk(NB, n) for _ in range(6)]
        lm, ep = np.linspace(1e-3, 1e-4, 11), np.linspace(1, 5, 11)
        grid = np.array(np.meshgrid(lm, ep)).T.reshape(-1, 2)[:120]
        c0 = torch.full((n,), 1.0 / n, device="cuda")
        fh, iu = zer(NR), vec(NI)
        W = 8 * n * NR + 4 * n * NI + 12 * n * NB
        return len(grid), W, lambda cfg: sweep.ac_sweep(grid, c0, *m, fh, iu, E, 0.0001, want_c=True, cfg=cfg)
    NR, NI, NB = 10000, 512, 1024
    m = [mk(NR, n), mk(NR, n), mk(NR, n), mk(NI, n), mk(NI, n), mk(NB, n)]
    grid = kg_grid(10, 10, 10)[:int(os.environ.get("CR_KG_NP", "1000"))]
    c0 = torch.full((n,), 1.0 / n, device="cuda")
    xc, fh, u1, u2, bu = vec(NR), zer(NR), vec(NI), zer(NI), vec(NB)
    W = 8 * n * NR + 4 * n * (NB + 2 * NI)
    return len(grid), W, lambda cfg: sweep.kg_sweep(grid, c0, *m, xc, fh, u1, u2, bu, E, 0.001, want_c=True, cfg=cfg)


for case in os.environ.get("CR_CASES", "b9,ac9,kg15").split(","):
    Np, W, run = problem(case)
    ref = run(dict(mode=2))
    cfgs = [None, dict(mode=2), dict(mode=3)] + [dict(mode=4, K=k) for k in (16, 8, 4, 2, 1)]
    for cfg in cfgs:
        try:
            r = run(cfg); torch.cuda.synchronize()
        except Exception as e:
            print(f"{case:5s} cfg={cfg}: {str(e)[:120]}")
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = run(cfg); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ef = float(((r["f"] - ref["f"]).abs() / ref["f"].abs()).max())
        em = float(((r["m"] - ref["m"]).abs() / ref["m"].abs()).max())
        ec = float(((r["c"] - ref["c"]).norm(dim=1) / ref["c"].norm(dim=1)).max())
        i = r["info"]
        if ef > 1e-3:
            bad = torch.nonzero((r["f"] - ref["f"]).abs() / ref["f"].abs() > 1e-3).flatten().tolist()
            print("      bad parameters:", bad[:40], "of", Np)
        print(f"{case:5s} cfg={str(cfg):24s} {1e3 * ms / E:8.2f} us/epoch  frac {W * Np * E / (ms * 1e-3) / 72.5e12:5.3f}  err f {ef:.1e} m {em:.1e} c {ec:.1e}  "
              f"priv {i['priv']} K {i['K']} ctas {i['ctas']} SL {i['SL']} smem {i['smem_bytes']}", flush=True)
