"""GPU parity: the CUDA Klein-Gordon sweep (through the C ABI) vs the reference's golden vectors
and vs the CPU oracle on seeded inputs.  Tolerance 1e-5 relative (north_star), identical selection."""
import numpy as np
import pytest
import torch

from conftest import load_golden, lr_of, rel_scalar, rel_vec, selection_ok

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev(g, names):
    return {k: torch.from_numpy(np.ascontiguousarray(g[k])).cuda() for k in names}


KG_NAMES = ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat", "IC_u1", "IC_u2", "BC_u")


def _run(d, n, grid, epochs, lr, rec_every=0, cfg=None, c0=None):
    from gptpinn import sweep
    if c0 is None:
        c0 = torch.full((n,), 1.0 / n, device="cuda")
    # column-slice VIEWS with the parent's row stride, exactly what KG_precomp.inputs returns
    v = {k: d[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
    r = sweep.kg_sweep(grid, c0, v["out"], v["out_xx"], v["out_tt"], v["out_IC"], v["out_IC_t"], v["out_BC"],
                       d["xcos"], d["f_hat"], d["IC_u1"], d["IC_u2"], d["BC_u"], epochs, lr,
                       rec_every=rec_every, cfg=cfg)
    torch.cuda.synchronize()
    return r


@pytest.mark.parametrize("case", ["kg_small", "kg_full", "kg_nonmono"])
def test_kg_matches_reference_golden(case):
    from gptpinn import sweep
    g = load_golden(case)
    d = _dev(g, KG_NAMES)
    for i, n in enumerate(g["ns"]):
        n = int(n)
        r = _run(d, n, g["grid"], int(g["epochs"]), lr_of(g, i), int(g["rec_every"]))
        f, m, c, tr = (r[k].cpu().numpy() for k in ("f", "m", "c", "trace"))
        assert rel_scalar(f, g[f"n{n}_f"]) < TOL, (case, n)
        assert rel_scalar(m, g[f"n{n}_m"]) < TOL, (case, n)
        assert rel_vec(c, g[f"n{n}_c"]) < TOL, (case, n)
        assert rel_scalar(tr, g[f"n{n}_trace"]) < TOL, (case, n)
        L, idx = sweep.argmax_scan(r["f"], r["m"])
        iref = int(np.nonzero((g["grid"] == g[f"n{n}_case"]).all(1))[0][0])
        assert selection_ok(g[f"n{n}_f"], iref, idx), (case, n)      # kg_small n = 1: two parameters tie to 1 ulp
        assert abs(float(L) - float(g[f"n{n}_L"])) <= TOL * float(g[f"n{n}_L"])


@pytest.mark.parametrize("cfg", [dict(M=1, SL=1, mode=1), dict(M=1, SL=32, mode=1), dict(M=2, SL=4, mode=1),
                                 dict(M=4, SL=2, mode=1), dict(M=5, SL=1, mode=1), dict(M=5, SL=8, W=3, mode=1),
                                 dict(M=2, SL=16, W=16, mode=1),
                                 dict(M=1, SL=1, mode=2), dict(M=1, SL=32, mode=2), dict(M=2, SL=4, mode=2),
                                 dict(M=4, SL=2, mode=2), dict(M=5, SL=1, mode=2), dict(M=5, SL=8, W=3, mode=2),
                                 dict(M=5, mode=2), dict(M=4, mode=2), dict(M=2, mode=2), dict(mode=1), dict(mode=2),
                                 # teams of K warps on one work item (tiles of a pass split K ways, gradients summed per epoch)
                                 dict(M=5, SL=8, mode=2, K=2), dict(M=5, SL=1, mode=2, K=8), dict(M=2, SL=4, mode=2, K=4),
                                 dict(M=4, SL=2, mode=2, K=2), dict(M=1, SL=32, mode=2, K=8), dict(mode=2, K=4),
                                 dict(M=5, SL=2, W=6, mode=2, K=2)])
def test_kg_every_mapping_agrees(cfg):
    """All (M, SL, W) mappings compute the same thing (different summation trees only)."""
    g = load_golden("kg_small")
    d = _dev(g, KG_NAMES)
    for n in (2, 15):
        r = _run(d, n, g["grid"], int(g["epochs"]), lr_of(g, 0), cfg=cfg)
        if "M" in cfg:
            assert r["info"]["M"] == cfg["M"]
        assert r["info"]["priv"] == cfg["mode"] - 1
        if "K" in cfg:
            assert r["info"]["K"] == cfg["K"]
        assert rel_scalar(r["f"].cpu().numpy(), g[f"n{n}_f"]) < TOL
        assert rel_vec(r["c"].cpu().numpy(), g[f"n{n}_c"]) < TOL


def test_kg_all_n_against_oracle():
    """n = 1..15 on seeded inputs, CUDA vs the C oracle (f32 build) and the f64 adjudicator."""
    from oracle import oracle
    g = load_golden("kg_small")
    d = _dev(g, KG_NAMES)
    grid = g["grid"][::3]
    for n in range(1, 16):
        pr = oracle.kg_problem(n, g["out"], g["out_xx"], g["out_tt"], g["out_IC"], g["out_IC_t"], g["out_BC"],
                               g["xcos"], g["f_hat"], g["IC_u1"], g["IC_u2"], g["BC_u"])
        o = oracle.sweep("kg", pr, grid, np.full(n, 1.0 / n, np.float32), 150, 0.025, 0, "f64")
        r = _run(d, n, grid, 150, 0.025)
        assert rel_scalar(r["f"].cpu().numpy(), o["f"]) < TOL, n
        assert rel_scalar(r["m"].cpu().numpy(), o["m"]) < TOL, n
        assert rel_vec(r["c"].cpu().numpy(), o["c"]) < TOL, n


def test_kg_nonzero_fhat_icu2_and_per_param_c0():
    """The update ignores f_hat / IC_u2 but the loss does not; c0 may differ per parameter."""
    from oracle import oracle
    g = load_golden("kg_small")
    rng = np.random.default_rng(5)
    h = {k: np.array(g[k]) for k in KG_NAMES}
    h["f_hat"] = (0.3 * rng.standard_normal(h["f_hat"].shape)).astype(np.float32)
    h["IC_u2"] = (0.2 * rng.standard_normal(h["IC_u2"].shape)).astype(np.float32)
    d = {k: torch.from_numpy(v).cuda() for k, v in h.items()}
    n = 7
    grid = g["grid"][:17]
    c0 = (rng.standard_normal((len(grid), n)) * 0.2).astype(np.float32)
    pr = oracle.kg_problem(n, h["out"], h["out_xx"], h["out_tt"], h["out_IC"], h["out_IC_t"], h["out_BC"],
                           h["xcos"], h["f_hat"], h["IC_u1"], h["IC_u2"], h["BC_u"])
    o = oracle.sweep("kg", pr, grid, c0, 120, 0.025, 40, "f64")
    r = _run(d, n, grid, 120, 0.025, rec_every=40, c0=torch.from_numpy(c0).cuda())
    assert rel_scalar(r["f"].cpu().numpy(), o["f"]) < TOL
    assert rel_vec(r["c"].cpu().numpy(), o["c"]) < TOL
    assert rel_scalar(r["trace"].cpu().numpy(), o["trace"]) < TOL


def test_kg_edge_cases():
    from gptpinn import sweep
    g = load_golden("kg_small")
    d = _dev(g, KG_NAMES)
    # empty grid
    r = _run(d, 3, np.zeros((0, 3)), 10, 0.025)
    assert r["f"].numel() == 0
    assert sweep.argmax_scan(r["f"], r["m"])[1] == -1
    # zero epochs: f = loss(c0), m = +inf (no prefix)
    r = _run(d, 3, g["grid"][:5], 0, 0.025)
    assert torch.isinf(r["m"]).all()
    assert torch.allclose(r["c"], torch.full_like(r["c"], 1.0 / 3))
    # one parameter, duplicated parameters give identical answers (determinism)
    grid = np.repeat(g["grid"][7:8], 9, axis=0)
    r = _run(d, 15, grid, 60, 0.025)
    assert (r["f"] == r["f"][0]).all() and (r["c"] == r["c"][0]).all()
    # bit-reproducible run to run
    r2 = _run(d, 15, grid, 60, 0.025)
    assert torch.equal(r["f"], r2["f"]) and torch.equal(r["c"], r2["c"])


def test_argmax_scan_matches_oracle_on_random_and_ties():
    from gptpinn import sweep
    from oracle import oracle
    rng = np.random.default_rng(0)
    for Np in (1, 5, 1023, 1024, 1025, 5000, 100000):
        f = rng.random(Np).astype(np.float32)
        m = (f * rng.choice([0.5, 1.0, 1.0, 1.0], Np)).astype(np.float32)
        f[rng.integers(0, Np, max(1, Np // 50))] = np.float32(0.75)       # exact ties
        if Np > 10:
            f[3] = np.nan
            m[5] = np.nan
        L, idx = sweep.argmax_scan(torch.from_numpy(f).cuda(), torch.from_numpy(m).cuda())
        Lo, io = oracle.argmax_scan(f, m)
        assert idx == io and float(L) == Lo, Np
    # small adversarial cases from a discrete value set: ties at the maximum, non-positive losses, NaN in f and m,
    # winners whose minimum dipped below an earlier final loss (the order-dependent fallback), nothing accepted
    vals = np.array([-1.0, 0.0, 0.25, 0.5, 0.75, 1.0, np.nan], np.float32)
    for trial in range(300):
        Np = int(rng.integers(1, 48))
        f = vals[rng.integers(0, len(vals), Np)]
        m = vals[rng.integers(0, len(vals), Np)]
        L, idx = sweep.argmax_scan(torch.from_numpy(f).cuda(), torch.from_numpy(m).cuda())
        Lo, io = oracle.argmax_scan(f, m)
        assert idx == io and float(L) == Lo, (trial, f, m)
    big = np.full(5000, 0.5, np.float32); mb = big.copy()
    big[4000] = 0.9; mb[4000] = 0.1; big[10] = 0.6                          # the arg-max is NOT accepted (m_w < 0.6): fallback
    L, idx = sweep.argmax_scan(torch.from_numpy(big).cuda(), torch.from_numpy(mb).cuda())
    assert (float(L), idx) == oracle.argmax_scan(big, mb)
    neg = -np.ones(3000, np.float32)
    assert sweep.argmax_scan(torch.from_numpy(neg).cuda(), torch.from_numpy(neg).cuda())[1] == -1


def test_kg_dense_rows_config5_properties():
    """N_R = 10^6 rows (BASELINE config 5 size), few parameters and epochs: CUDA vs the f64 oracle, and the
    row-permutation invariance of the update (a size-independent property of the sums over rows)."""
    from oracle import oracle
    rng = np.random.default_rng(1)
    NR, NI, NB, n = 1_000_000, 512, 1024, 15
    base = (rng.standard_normal((4096, 15)) * 0.3).astype(np.float32)
    h = dict(out=np.tile(base, (NR // 4096 + 1, 1))[:NR] * (1 + 0.01 * rng.standard_normal((NR, 1)).astype(np.float32)))
    h["out_xx"] = np.roll(h["out"], 1, axis=1) * 0.7
    h["out_tt"] = np.roll(h["out"], 2, axis=1) * 0.5
    h["out_IC"] = (rng.standard_normal((NI, 15)) * 0.3).astype(np.float32)
    h["out_IC_t"] = (rng.standard_normal((NI, 15)) * 0.3).astype(np.float32)
    h["out_BC"] = (rng.standard_normal((NB, 15)) * 0.3).astype(np.float32)
    h["xcos"] = (rng.standard_normal((NR, 1)) * 0.3).astype(np.float32)
    h["f_hat"] = np.zeros((NR, 1), np.float32)
    h["IC_u1"] = rng.standard_normal((NI, 1)).astype(np.float32)
    h["IC_u2"] = np.zeros((NI, 1), np.float32)
    h["BC_u"] = rng.standard_normal((NB, 1)).astype(np.float32)
    h = {k: np.ascontiguousarray(v.astype(np.float32)) for k, v in h.items()}
    d = {k: torch.from_numpy(v).cuda() for k, v in h.items()}
    grid = np.array([[-1.5, 0.5, 0.2], [-1.5, 0.5, 0.9], [-1.1, 0.1, 0.5], [-1.9, 0.8, 0.0]])
    r = _run(d, n, grid, 6, 0.01, rec_every=2)
    pr = oracle.kg_problem(n, *(h[k] for k in KG_NAMES))
    o = oracle.sweep("kg", pr, grid, np.full(n, 1.0 / n, np.float32), 6, 0.01, 2, "f64")
    assert rel_scalar(r["f"].cpu().numpy(), o["f"]) < TOL
    assert rel_vec(r["c"].cpu().numpy(), o["c"]) < TOL
    assert rel_scalar(r["trace"].cpu().numpy(), o["trace"]) < TOL
    perm = torch.from_numpy(rng.permutation(NR)).cuda()
    d2 = dict(d)
    for k in ("out", "out_xx", "out_tt", "xcos", "f_hat"):
        d2[k] = d[k][perm].contiguous()
    r2 = _run(d2, n, grid, 6, 0.01)
    assert rel_scalar(r2["f"].cpu().numpy(), r["f"].cpu().numpy()) < TOL
    assert rel_vec(r2["c"].cpu().numpy(), r["c"].cpu().numpy()) < TOL


def test_kg_config4_grid_properties():
    """The full 10^5-parameter grid of BASELINE config 4 (few epochs): results do not depend on the order of the
    grid rows (bitwise: the same sibling sets land in the same work items), on the team size, or on how the grid
    is cut into per-rank shards; a random sample of parameters is checked against the f64 oracle."""
    from gptpinn import offline
    from oracle import oracle
    g = load_golden("kg_full")
    d = _dev(g, KG_NAMES)
    n, epochs, lr = 15, 3, 0.025
    a = np.linspace(-2, -1, 50); b = np.linspace(0, 1, 50); gm = np.linspace(0, 1, 40)
    grid = np.array(np.meshgrid(a, b, gm)).T.reshape(-1, 3)
    r = _run(d, n, grid, epochs, lr)
    f, m, c = (r[k].cpu().numpy() for k in ("f", "m", "c"))
    assert np.isfinite(f).all() and (m >= 0).all()
    # (1) permutation of the grid rows
    rng = np.random.default_rng(7)
    perm = rng.permutation(len(grid))
    rp = _run(d, n, grid[perm], epochs, lr)
    assert np.array_equal(rp["f"].cpu().numpy(), f[perm])
    assert np.array_equal(rp["c"].cpu().numpy(), c[perm])
    # (2) independent warps vs teams of 4 warps vs the shared ring: different summation trees only
    for cfg in (dict(mode=2, K=1), dict(mode=2, K=4, M=5), dict(mode=1)):
        rk = _run(d, n, grid, epochs, lr, cfg=cfg)
        assert rel_scalar(rk["f"].cpu().numpy(), f) < TOL, cfg
        assert rel_vec(rk["c"].cpu().numpy(), c) < TOL, cfg
    # (3) the 8-way shard of rank 3, as offline.sharded_sweep cuts it
    order = np.lexsort((grid[:, 1], grid[:, 0]))
    lo, hi = offline.shard_bounds(len(grid), 3, 8)
    rows = order[lo:hi]
    rs = _run(d, n, grid[rows], epochs, lr)
    assert rs["info"]["K"] > 1                                  # few items per warp: the planner forms teams
    assert rel_scalar(rs["f"].cpu().numpy(), f[rows]) < TOL
    assert rel_vec(rs["c"].cpu().numpy(), c[rows]) < TOL
    # (4) sample vs the oracle
    pick = rng.choice(len(grid), 48, replace=False)
    pr = oracle.kg_problem(n, *(g[k] for k in KG_NAMES))
    o = oracle.sweep("kg", pr, grid[pick], np.full(n, 1.0 / n, np.float32), epochs, lr, 0, "f64")
    assert rel_scalar(f[pick], o["f"]) < TOL
    assert rel_scalar(m[pick], o["m"]) < TOL
    assert rel_vec(c[pick], o["c"]) < TOL


def test_kg_diverging_parameters_nan_semantics():
    """A learning rate at which part of the grid diverges (parity case 2): 16 of 60 parameters end in NaN, 4 in
    losses of 1e9..1e23, the rest converge.  The CUDA path must produce NaN for the same parameters, agree on the finite
    ones, and select what the reference's literal loop selects (a NaN loss never breaks and never wins)."""
    from gptpinn import sweep
    from oracle import oracle
    g = load_golden("kg_small")
    d = _dev(g, KG_NAMES)
    n, epochs, lr = 7, 60, 0.14
    pr = oracle.kg_problem(n, *(g[k] for k in KG_NAMES))
    c0 = np.full(n, 1.0 / n, np.float32)
    o = oracle.sweep("kg", pr, g["grid"], c0, epochs, lr, 0, "f64")
    Lo, io = oracle.kg_offline_literal(pr, g["grid"], c0, epochs, lr, "f64")
    assert np.isnan(o["f"]).sum() >= 8 and np.isfinite(o["f"]).sum() >= 8          # the case is what it claims to be
    for cfg in (None, dict(mode=1), dict(mode=2, K=1)):
        r = _run(d, n, g["grid"], epochs, lr, cfg=cfg)
        f, m = r["f"].cpu().numpy(), r["m"].cpu().numpy()
        assert np.array_equal(np.isnan(f), np.isnan(o["f"])), cfg
        fin = np.isfinite(o["f"])
        assert np.max(np.abs(f[fin] - o["f"][fin]) / np.abs(o["f"][fin])) < 1e-4, cfg   # diverging-but-finite losses amplify rounding
        small = fin & (o["f"] < 10)
        assert rel_scalar(f[small], o["f"][small]) < TOL, cfg
        L, idx = sweep.argmax_scan(r["f"], r["m"])
        assert idx == io and abs(float(L) - Lo) <= 1e-4 * abs(Lo), cfg
        assert (float(L), idx) == oracle.argmax_scan(f, m)
