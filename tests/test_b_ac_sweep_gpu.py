"""GPU parity: Burgers and Allen-Cahn sweeps (C ABI) vs the reference's golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, lr_of, rel_scalar, rel_vec, selection_ok

pytestmark = pytest.mark.gpu
TOL = 1e-5

B_NAMES = ("out", "out_x", "out_t", "out_xx", "out_IC", "out_BC", "f_hat", "IC_u", "BC_u")
AC_NAMES = ("out", "out_t", "out_xx", "out_IC", "out_BC_ub", "out_BC_lb", "out_BC_diff", "out_BC_ub_x",
            "out_BC_lb_x", "out_BC_diff_x", "f_hat", "IC_u")


def _dev(g, names):
    return {k: torch.from_numpy(np.ascontiguousarray(g[k])).cuda() for k in names}


def _b_run(d, n, grid, c0, epochs, lr, rec_every=0, cfg=None):
    from gptpinn import sweep
    v = {k: d[k][:, :n] for k in ("out", "out_x", "out_t", "out_xx", "out_IC", "out_BC")}
    r = sweep.b_sweep(grid, c0, v["out"], v["out_x"], v["out_t"], v["out_xx"], v["out_IC"], v["out_BC"],
                      d["f_hat"], d["IC_u"], d["BC_u"], epochs, lr, rec_every=rec_every, cfg=cfg)
    torch.cuda.synchronize()
    return r


def _ac_run(d, n, grid, c0, epochs, lr, rec_every=0, cfg=None):
    from gptpinn import sweep
    v = {k: d[k][:, :n] for k in AC_NAMES[:10]}
    r = sweep.ac_sweep(grid, c0, v["out"], v["out_t"], v["out_xx"], v["out_IC"], v["out_BC_ub"], v["out_BC_lb"],
                       v["out_BC_diff"], v["out_BC_ub_x"], v["out_BC_lb_x"], v["out_BC_diff_x"], d["f_hat"],
                       d["IC_u"], epochs, lr, rec_every=rec_every, cfg=cfg)
    torch.cuda.synchronize()
    return r


@pytest.mark.parametrize("cfg", [None, dict(mode=1), dict(mode=2), dict(M=2, SL=1, mode=1), dict(M=1, SL=4, mode=2),
                                 dict(mode=2, K=2), dict(M=2, SL=1, mode=2, K=4)])
def test_burgers_matches_reference_golden(cfg):
    from gptpinn import sweep
    g = load_golden("b_small")
    d = _dev(g, B_NAMES)
    for i, n in enumerate(g["ns"]):
        n = int(n)
        c0 = torch.from_numpy(g[f"n{n}_c0"]).cuda()
        r = _b_run(d, n, g["grid"], c0, int(g["epochs"]), lr_of(g, i), int(g["rec_every"]), cfg)
        assert rel_scalar(r["f"].cpu().numpy(), g[f"n{n}_f"]) < TOL, n
        assert rel_scalar(r["m"].cpu().numpy(), g[f"n{n}_m"]) < TOL, n
        assert rel_vec(r["c"].cpu().numpy(), g[f"n{n}_c"]) < TOL, n
        assert rel_scalar(r["trace"].cpu().numpy(), g[f"n{n}_trace"]) < TOL, n
        L, idx = sweep.argmax_scan(r["f"], r["m"])
        assert g["grid"][idx] == float(g[f"n{n}_case"])
        assert abs(float(L) - float(g[f"n{n}_L"])) <= TOL * float(L)


@pytest.mark.parametrize("cfg", [None, dict(mode=1), dict(mode=2), dict(M=2, SL=1, mode=1), dict(mode=2, K=4),
                                 dict(M=1, SL=2, mode=2, K=8)])
def test_allen_cahn_matches_reference_golden(cfg):
    from gptpinn import sweep
    from oracle import oracle
    g = load_golden("ac_small")
    d = _dev(g, AC_NAMES)
    for i, n in enumerate(g["ns"]):
        n = int(n)
        c0 = torch.full((n,), 1.0 / n, device="cuda")
        r = _ac_run(d, n, g["grid"], c0, int(g["epochs"]), lr_of(g, i), int(g["rec_every"]), cfg)
        assert rel_scalar(r["f"].cpu().numpy(), g[f"n{n}_f"]) < TOL, n
        assert rel_scalar(r["m"].cpu().numpy(), g[f"n{n}_m"]) < TOL, n
        assert rel_vec(r["c"].cpu().numpy(), g[f"n{n}_c"]) < TOL, n
        L, idx = sweep.argmax_scan(r["f"], r["m"])
        _, iref = oracle.argmax_scan(g[f"n{n}_f"], g[f"n{n}_m"])
        assert selection_ok(g[f"n{n}_f"], iref, idx), n           # n = 1 ties to 2 ulp across lambda


def test_burgers_nonzero_targets_against_oracle():
    """BC_u and f_hat enter the loss but not the update."""
    from oracle import oracle
    g = load_golden("b_small")
    rng = np.random.default_rng(3)
    h = {k: np.array(g[k]) for k in B_NAMES}
    h["f_hat"] = (0.2 * rng.standard_normal(h["f_hat"].shape)).astype(np.float32)
    h["BC_u"] = (0.3 * rng.standard_normal(h["BC_u"].shape)).astype(np.float32)
    d = {k: torch.from_numpy(v).cuda() for k, v in h.items()}
    n = 5
    pr = oracle.b_problem(n, h["out"], h["out_x"], h["out_t"], h["out_xx"], h["out_IC"], h["out_BC"], h["f_hat"],
                          h["IC_u"], h["BC_u"])
    o = oracle.sweep("b", pr, g["grid"], g["n5_c0"], 100, 0.02, 25, "f64")
    r = _b_run(d, n, g["grid"], torch.from_numpy(g["n5_c0"]).cuda(), 100, 0.02, 25)
    assert rel_scalar(r["f"].cpu().numpy(), o["f"]) < TOL
    assert rel_vec(r["c"].cpu().numpy(), o["c"]) < TOL
    assert rel_scalar(r["trace"].cpu().numpy(), o["trace"]) < TOL


def test_all_n_burgers_ac_against_oracle():
    from oracle import oracle
    gb, ga = load_golden("b_small"), load_golden("ac_small")
    db, da = _dev(gb, B_NAMES), _dev(ga, AC_NAMES)
    for n in range(1, 10):
        pr = oracle.b_problem(n, gb["out"], gb["out_x"], gb["out_t"], gb["out_xx"], gb["out_IC"], gb["out_BC"],
                              gb["f_hat"], gb["IC_u"], gb["BC_u"])
        c0 = np.full((len(gb["grid"]), n), 1.0 / n, np.float32)
        o = oracle.sweep("b", pr, gb["grid"], c0, 80, 0.02, 0, "f64")
        r = _b_run(db, n, gb["grid"], torch.from_numpy(c0).cuda(), 80, 0.02)
        assert rel_scalar(r["f"].cpu().numpy(), o["f"]) < TOL, ("b", n)
        assert rel_vec(r["c"].cpu().numpy(), o["c"]) < TOL, ("b", n)
        pr = oracle.ac_problem(n, ga["out"], ga["out_t"], ga["out_xx"], ga["out_IC"], ga["out_BC_ub"],
                               ga["out_BC_lb"], ga["out_BC_diff"], ga["out_BC_ub_x"], ga["out_BC_lb_x"],
                               ga["out_BC_diff_x"], ga["f_hat"], ga["IC_u"])
        o = oracle.sweep("ac", pr, ga["grid"], np.full(n, 1.0 / n, np.float32), 80, 0.0025, 0, "f64")
        r = _ac_run(da, n, ga["grid"], torch.full((n,), 1.0 / n, device="cuda"), 80, 0.0025)
        assert rel_scalar(r["f"].cpu().numpy(), o["f"]) < TOL, ("ac", n)
        assert rel_vec(r["c"].cpu().numpy(), o["c"]) < TOL, ("ac", n)


def test_burgers_and_allen_cahn_at_reference_default_sizes():
    """The reference's default problem sizes (Burgers N_R = 10 000 / N_IC = 200 / N_BC = 400, Allen-Cahn N_R = 20 000 /
    N_IC = 512 / N_BC = 100 per side), n = 9, 2000 epochs: fixtures `b_full`, `ac_full` from the unmodified reference."""
    from gptpinn import sweep
    b = load_golden("b_full")
    d = _dev(b, B_NAMES)
    n = int(b["ns"][0])
    r = _b_run(d, n, b["grid"], torch.from_numpy(b[f"n{n}_c0"]).cuda(), int(b["epochs"]), lr_of(b, 0), int(b["rec_every"]))
    assert rel_scalar(r["f"].cpu().numpy(), b[f"n{n}_f"]) < TOL
    assert rel_scalar(r["m"].cpu().numpy(), b[f"n{n}_m"]) < TOL
    assert rel_vec(r["c"].cpu().numpy(), b[f"n{n}_c"]) < TOL
    assert rel_scalar(r["trace"].cpu().numpy(), b[f"n{n}_trace"]) < TOL
    L, idx = sweep.argmax_scan(r["f"], r["m"])
    assert b["grid"][idx] == float(b[f"n{n}_case"]) and abs(float(L) - float(b[f"n{n}_L"])) <= TOL * float(L)
    a = load_golden("ac_full")
    d = _dev(a, AC_NAMES)
    n = int(a["ns"][0])
    r = _ac_run(d, n, a["grid"], torch.full((n,), 1.0 / n, device="cuda"), int(a["epochs"]), lr_of(a, 0), int(a["rec_every"]))
    assert rel_scalar(r["f"].cpu().numpy(), a[f"n{n}_f"]) < TOL
    assert rel_scalar(r["m"].cpu().numpy(), a[f"n{n}_m"]) < TOL
    assert rel_vec(r["c"].cpu().numpy(), a[f"n{n}_c"]) < TOL
    L, idx = sweep.argmax_scan(r["f"], r["m"])
    assert np.array_equal(a["grid"][idx], a[f"n{n}_case"]) and abs(float(L) - float(a[f"n{n}_L"])) <= TOL * float(L)
