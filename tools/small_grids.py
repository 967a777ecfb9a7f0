"""Per-stage sweep times at the reference's default grid sizes (KG 1000, Burgers 128, Allen-Cahn 120)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpt-pinn_b200"), os.path.join(ROOT, "gpt-pinn_b200", "dropin")):
    sys.path.insert(0, p)
from gptpinn import sweep
from bench import kg_grid

g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda r, n: torch.randn(r, n, device="cuda", generator=g) * 0.3
vec = lambda r: torch.randn(r, 1, device="cuda", generator=g) * 0.3
zer = lambda r: torch.zeros(r, 1, device="cuda")


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


for n in (1, 8, 15):
    NR, NI, NB = 10000, 512, 1024
    m = [mk(NR, 15), mk(NR, 15), mk(NR, 15), mk(NI, 15), mk(NI, 15), mk(NB, 15)]
    grid = kg_grid(10, 10, 10)
    c0 = torch.full((n,), 1.0 / n, device="cuda")
    xc, fh, u1, u2, bu = vec(NR), zer(NR), vec(NI), zer(NI), vec(NB)
    for cfg in (None, dict(mode=1), dict(mode=2), dict(mode=3)):
        dt, r = timeit(lambda: sweep.kg_sweep(grid, c0, *[a[:, :n] for a in m], xc, fh, u1, u2, bu, 2000, 0.001, want_c=False, cfg=cfg))
        print(f"KG  n={n:2d} 1000 params x 2000 epochs cfg={cfg}: {1e3*dt:8.1f} ms  {1000*2000/dt/1e6:6.2f} M pe/s  {r['info']}")
for n in (1, 9):
    NR, NI, NB = 10000, 200, 400
    m = [mk(NR, 9), mk(NR, 9), mk(NR, 9), mk(NR, 9), mk(NI, 9), mk(NB, 9)]
    grid = np.linspace(0.005, 1, 129)[:128]
    c0 = torch.full((128, n), 1.0 / n, device="cuda")
    for cfg in (None, dict(mode=1), dict(mode=2), dict(mode=3)):
        dt, r = timeit(lambda: sweep.b_sweep(grid, c0, *[a[:, :n] for a in m], zer(NR), vec(NI), zer(NB), 2000, 0.001, want_c=False, cfg=cfg))
        print(f"B   n={n:2d}  128 params x 2000 epochs cfg={cfg}: {1e3*dt:8.1f} ms  {128*2000/dt/1e6:6.2f} M pe/s  {r['info']}")
for n in (1, 9):
    NR, NI, NB = 20000, 512, 100
    m = [mk(NR, 9), mk(NR, 9), mk(NR, 9), mk(NI, 9)] + [mk(NB, 9) for _ in range(6)]
    lm, ep = np.linspace(1e-3, 1e-4, 11), np.linspace(1, 5, 11)
    grid = np.array(np.meshgrid(lm, ep)).T.reshape(-1, 2)[:120]
    c0 = torch.full((n,), 1.0 / n, device="cuda")
    for cfg in (None, dict(mode=1), dict(mode=2), dict(mode=3)):
        dt, r = timeit(lambda: sweep.ac_sweep(grid, c0, *[a[:, :n] for a in m], zer(NR), vec(NI), 2000, 0.0001, want_c=False, cfg=cfg))
        print(f"AC  n={n:2d}  120 params x 2000 epochs cfg={cfg}: {1e3*dt:8.1f} ms  {120*2000/dt/1e6:6.2f} M pe/s  {r['info']}")
