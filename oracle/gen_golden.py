#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (skoohy/GPT-PINN) on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference, which does
not exist on the GPU box); its outputs are committed under tests/golden/*.npz and are what
pins both the oracle restatement (oracle/) and the CUDA path.

How the reference is driven:
  * the reference folder is put on sys.path and its modules imported as-is;
  * the module attribute `device` (hard-coded torch.device("cuda")) is re-bound to cpu;
  * synthetic full-PINN "neurons" are random MLPs of the reference architecture pushed
    through the reference's own `inputs()` (KG_precomp.py:34-52, B_precomp.py:22-55,
    AC_precompute.py:26-58), so the activation matrices carry the reference's mixed
    second-derivative semantics;
  * per-parameter trajectories come from the reference classes `grad_descent.update`
    (KG_optimizer.py:34-68 ...) and `GPT.loss` (KG_models.py:45-50 ...), driven by the
    same loop as `offline_generation` (KG_train.py:27-37) without the early exit;
  * `(largest_loss, largest_case)` come from the reference `offline_generation` itself,
    and the test sweeps from `gpt_test` / `gpt_test_loss`.

Usage:  python oracle/gen_golden.py [kg_small|kg_full|b_small|ac_small|kg_precomp|...|all]
"""
import importlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get("GPTPINN_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CPU = torch.device("cpu")


def _import_ref(folder, names):
    path = os.path.join(REF, folder)
    if path not in sys.path:
        sys.path.insert(0, path)
    mods = {}
    for n in names:
        m = importlib.import_module(n)
        if hasattr(m, "device"):
            m.device = CPU
        mods[n] = m
    return mods


def _randomize(pinn, seed, scale=0.6, pert=0.05):
    """Seeded synthetic weights: Xavier-normal x scale + N(0, pert^2), zero bias."""
    g = torch.Generator().manual_seed(seed)
    for lin in pinn.linears:
        fan_out, fan_in = lin.weight.shape
        std = (2.0 / (fan_in + fan_out)) ** 0.5
        w = torch.randn(lin.weight.shape, generator=g) * std * scale
        w = w + torch.randn(lin.weight.shape, generator=g) * pert
        lin.weight.data = w.float()
        lin.bias.data = torch.zeros_like(lin.bias.data)


def _weights(pinn):
    out = {}
    for li, lin in enumerate(pinn.linears):
        out[f"W{li}"] = lin.weight.data.numpy().copy()
        out[f"b{li}"] = lin.bias.data.numpy().copy()
    return out


# --------------------------------------------------------------------------------------
# Klein-Gordon
# --------------------------------------------------------------------------------------
def kg_problem(Nc, IC_pts, BC_pts, N_test, nmax, seed0=1000):
    m = _import_ref("Klein-Gordon", ["KG_data", "KG_models", "KG_precomp", "KG_optimizer",
                                     "KG_train", "KG_test"])
    xt_resid, f_hat, xt_test = m["KG_data"].residual_data(-1.0, 1.0, 0.0, 5.0, Nc, N_test)
    IC_xt, IC_u1, IC_u2, BC_xt, BC_u = m["KG_data"].ICBC_data(-1.0, 1.0, 0.0, 5.0, BC_pts, IC_pts)
    xcos = m["KG_precomp"].xcos_term(xt_resid[:, 0].unsqueeze(1), xt_resid[:, 1].unsqueeze(1))
    xt_resid = xt_resid.requires_grad_()
    IC_xt = IC_xt.requires_grad_()
    N, NI, NB, NT = xt_resid.shape[0], IC_xt.shape[0], BC_xt.shape[0], xt_test.shape[0]
    bufs = dict(out=torch.ones(N, nmax), out_xx=torch.ones(N, nmax), out_tt=torch.ones(N, nmax),
                out_IC=torch.ones(NI, nmax), out_IC_t=torch.ones(NI, nmax),
                out_BC=torch.ones(NB, nmax), out_test=torch.zeros(NT, nmax))
    layers = np.array([2, 40, 40, 1])
    weights = []
    for i in range(nmax):
        pinn = m["KG_models"].NN(layers, -1.5, 0.5, 0.5, xcos)
        _randomize(pinn, seed0 + i)
        weights.append(_weights(pinn))
        m["KG_precomp"].inputs(pinn, xt_resid, bufs["out"], bufs["out_xx"], bufs["out_tt"],
                               bufs["out_IC_t"], bufs["out_IC"], bufs["out_BC"], IC_xt, BC_xt,
                               i, bufs["out_test"], xt_test)
    bufs = {k: v.detach() for k, v in bufs.items()}
    return m, dict(xt_resid=xt_resid.detach(), xt_test=xt_test, IC_xt=IC_xt.detach(), BC_xt=BC_xt,
                   IC_u1=IC_u1, IC_u2=IC_u2, BC_u=BC_u, f_hat=f_hat, xcos=xcos.detach(),
                   weights=weights, **bufs)


def kg_trajectories(m, p, grid, n, epochs, lr, rec_every):
    """Per-mu: c_final, final loss f_p, min-prefix loss m_p, loss trace (KG_train.py:27-37 loop)."""
    v = {k: p[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
    N, NI, NB = p["out"].shape[0], p["out_IC"].shape[0], p["out_BC"].shape[0]
    c0 = torch.full((n,), 1.0 / n)
    nrec = epochs // rec_every + 1
    C = np.zeros((len(grid), n), np.float32)
    F = np.zeros(len(grid), np.float32)
    M = np.zeros(len(grid), np.float32)
    T = np.zeros((len(grid), nrec), np.float32)
    for gi, (a, b, g) in enumerate(grid):
        Tt = m["KG_precomp"].Ptt_aPxx_bP(a, b, v["out_tt"], v["out_xx"], v["out"])
        G2 = m["KG_precomp"].gamma2_P(g, v["out"])
        GD = m["KG_optimizer"].grad_descent(a, b, g, N, NI, NB, p["IC_u1"], p["IC_u2"], p["BC_u"],
                                            p["xcos"], Tt, G2, v["out"], v["out_IC"],
                                            v["out_IC_t"], v["out_BC"], epochs, lr)
        NNg = m["KG_models"].GPT(a, b, g, v["out"], v["out_IC"], v["out_IC_t"], v["out_BC"],
                                 p["xcos"], p["f_hat"], Tt, p["IC_u1"], p["IC_u2"], p["BC_u"])
        c = c0
        cr = c.unsqueeze(1)
        loss = NNg.loss(cr)
        mp = float("inf")
        T[gi, 0] = loss.item()
        for j in range(1, epochs + 1):
            mp = min(mp, loss.item())          # min over l_0 .. l_{E-1}
            c = GD.update(c, cr)
            cr = c.unsqueeze(1)
            loss = NNg.loss(cr)
            if j % rec_every == 0:
                T[gi, j // rec_every] = loss.item()
        C[gi] = c.numpy()
        F[gi] = loss.item()
        M[gi] = mp
    return C, F, M, T


def kg_case(tag, Nc, IC_pts, BC_pts, N_test, ns, grid, epochs, lr, rec_every, test_mu, test_epochs,
            nmax=15):
    m, p = kg_problem(Nc, IC_pts, BC_pts, N_test, nmax)
    N, NI, NB = p["out"].shape[0], p["out_IC"].shape[0], p["out_BC"].shape[0]
    save = {k: p[k].numpy() for k in ("xt_resid", "xt_test", "IC_xt", "BC_xt", "IC_u1", "IC_u2",
                                      "BC_u", "f_hat", "xcos", "out", "out_xx", "out_tt", "out_IC",
                                      "out_IC_t", "out_BC", "out_test")}
    for i, w in enumerate(p["weights"]):
        for k, a in w.items():
            save[f"pinn{i}_{k}"] = a
    lrs = lr if isinstance(lr, dict) else {n: lr for n in ns}
    save.update(grid=grid, ns=np.array(ns), epochs=epochs, lr=np.array([lrs[n] for n in ns]),
                rec_every=rec_every)
    for n in ns:
        lr = lrs[n]
        C, F, M, T = kg_trajectories(m, p, grid, n, epochs, lr, rec_every)
        v = {k: p[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
        L, case = m["KG_train"].offline_generation(
            grid, torch.full((n,), 1.0 / n), N, NI, NB, p["IC_u1"], p["IC_u2"], p["BC_u"], p["xcos"],
            v["out"], v["out_xx"], v["out_tt"], v["out_IC"], v["out_IC_t"], v["out_BC"], p["f_hat"],
            epochs, lr)
        save[f"n{n}_c"] = C
        save[f"n{n}_f"] = F
        save[f"n{n}_m"] = M
        save[f"n{n}_trace"] = T
        save[f"n{n}_L"] = np.float32(float(L))
        save[f"n{n}_case"] = np.array(case, np.float64)
        print(tag, "n", n, "L", float(L), "case", case, flush=True)
    if test_mu is not None:
        n = nmax
        c0 = torch.full((n,), 1.0 / n)
        times, soln = m["KG_test"].gpt_test(
            test_mu, p["out"], p["out_xx"], p["out_tt"], p["out_IC"], p["out_IC_t"], p["out_BC"],
            p["f_hat"], p["xcos"], N, NI, NB, p["IC_u1"], p["IC_u2"], p["BC_u"], c0, test_epochs, lr,
            p["out_test"])
        losses = m["KG_test"].gpt_test_loss(
            test_mu, p["out"], p["out_xx"], p["out_tt"], p["out_IC"], p["out_IC_t"], p["out_BC"],
            p["f_hat"], p["xcos"], N, NI, NB, p["IC_u1"], p["IC_u2"], p["BC_u"], c0, test_epochs, lr)
        save.update(test_mu=test_mu, test_epochs=test_epochs, test_soln=soln, test_losses=losses)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **save)
    print("wrote", tag)


def kg_grid(na, nb, ng):
    a = np.linspace(-2, -1, na)
    b = np.linspace(0, 1, nb)
    g = np.linspace(0, 1, ng)
    return np.array(np.meshgrid(a, b, g)).T.reshape(-1, 3)   # KG_main.py:46-49 ordering


def gen_kg_small():
    grid = kg_grid(4, 3, 5)
    rng = np.random.default_rng(1234)
    test_mu = grid[rng.choice(len(grid), 6, replace=False)]
    kg_case("kg_small", 24, 64, 64, 10, [1, 2, 7, 15], grid, 300, 0.025, 50, test_mu, 1000)


def gen_kg_nonmono():
    """Large lr: non-monotone / diverging trajectories exercise the exact N1 arg-max scan."""
    grid = kg_grid(3, 3, 4)
    kg_case("kg_nonmono", 16, 32, 32, 8, [3, 6], grid, 120, {3: 0.28, 6: 0.13}, 20, None, 0, nmax=6)


def gen_kg_full():
    grid = kg_grid(10, 10, 10)
    rng = np.random.default_rng(7)
    sel = np.sort(rng.choice(len(grid), 10, replace=False))
    kg_case("kg_full", 100, 512, 512, 40, [15], grid[sel], 2000, 0.025, 500, None, 0)


# --------------------------------------------------------------------------------------
# generic trajectory driver (the loop of *_train.offline_generation without the early exit)
# --------------------------------------------------------------------------------------
def trajectories(make_pair, grid, c0_of, n, epochs, rec_every):
    nrec = epochs // rec_every + 1
    C = np.zeros((len(grid), n), np.float32)
    F = np.zeros(len(grid), np.float32)
    M = np.zeros(len(grid), np.float32)
    T = np.zeros((len(grid), nrec), np.float32)
    for gi, mu in enumerate(grid):
        GD, NNg = make_pair(mu)
        c = c0_of(gi)
        cr = c.unsqueeze(1)
        loss = NNg.loss(cr)
        mp = float("inf")
        T[gi, 0] = loss.item()
        for j in range(1, epochs + 1):
            mp = min(mp, loss.item())
            c = GD.update(c, cr)
            cr = c.unsqueeze(1)
            loss = NNg.loss(cr)
            if j % rec_every == 0:
                T[gi, j // rec_every] = loss.item()
        C[gi] = c.numpy()
        F[gi] = loss.item()
        M[gi] = mp
    return C, F, M, T


# --------------------------------------------------------------------------------------
# Burgers
# --------------------------------------------------------------------------------------
def gen_b_small():
    _gen_burgers("b_small", 24, 10, 40, 40, [1, 2, 5, 9], np.delete(np.linspace(0.005, 1, 33), [16]), 300, 50, True)


def gen_b_full():
    """Reference default sizes (B_main.py:34-37: N_R = 10 000, N_IC = 200, N_BC = 400), n = 9, 2000 epochs, 6 parameters."""
    _gen_burgers("b_full", 100, 10, 200, 200, [9], np.linspace(0.005, 1, 129)[[3, 30, 57, 77, 101, 126]], 2000, 500, False)


def _gen_burgers(tag, Nc, N_test, BC_pts, IC_pts, ns, grid, epochs, rec, small):
    m = _import_ref("Burgers", ["B_data", "B_models", "B_precomp", "B_optimizer", "B_train", "B_test"])
    nmax = 9
    xt_resid, f_hat, xt_test = m["B_data"].residual_data(-1.0, 1.0, 0.0, 1.0, Nc, N_test)
    IC_xt, IC_u, BC_xt, BC_u = m["B_data"].ICBC_data(-1.0, 1.0, 0.0, 1.0, BC_pts, IC_pts)
    xt_resid = xt_resid.requires_grad_()
    N, NI, NB, NT = xt_resid.shape[0], IC_xt.shape[0], BC_xt.shape[0], xt_test.shape[0]
    z = lambda r: torch.zeros(r, nmax)
    out, out_t, out_x, out_xx, out_IC, out_BC, out_test = z(N), z(N), z(N), z(N), z(NI), z(NB), z(NT)
    num_mag = int(N * 0.2)
    idx_list = torch.zeros((nmax, num_mag), dtype=torch.long)
    layers = np.array([2, 20, 20, 20, 20, 1])
    neurons = np.linspace(0.005, 1, 129)[[64, 0, 128, 20, 100, 40, 80, 10, 118]]
    save = {}
    for i in range(nmax):
        pinn = m["B_models"].NN(layers, neurons[i])
        _randomize(pinn, 2000 + i, scale=1.0, pert=0.05)
        if small:
            for k, a in _weights(pinn).items():
                save[f"pinn{i}_{k}"] = a
        m["B_precomp"].inputs(pinn, xt_resid, out, out_t, out_x, out_xx, out_IC, out_BC, IC_xt, BC_xt, i,
                              out_test, xt_test, N, num_mag, idx_list)
    bufs = dict(out=out, out_t=out_t, out_x=out_x, out_xx=out_xx, out_IC=out_IC, out_BC=out_BC, out_test=out_test)
    bufs = {k: v.detach() for k, v in bufs.items()}
    save.update({k: v.numpy() for k, v in bufs.items()})
    save.update(xt_resid=xt_resid.detach().numpy(), xt_test=xt_test.numpy(), IC_xt=IC_xt.numpy(), BC_xt=BC_xt.numpy(),
                IC_u=IC_u.numpy(), BC_u=BC_u.numpy(), f_hat=f_hat.numpy(), idx_list=idx_list.numpy(), neurons=neurons)
    lr = 0.02
    save.update(grid=grid, ns=np.array(ns), epochs=epochs, lr=lr, rec_every=rec)
    for n in ns:
        v = {k: bufs[k][:, :n] for k in ("out", "out_t", "out_x", "out_xx", "out_IC", "out_BC")}
        # initial c per parameter exactly as B_train.offline_generation builds it (B_train.py:27-47)
        c0s = []
        for nu in grid:
            if n == 1:
                c0s.append(torch.ones(1))
                continue
            dist = np.abs(neurons[:n] - nu)
            d = np.argsort(dist)
            c0 = torch.zeros(n)
            c0[d[0]] = dist[d[1]] / (dist[d[0]] + dist[d[1]])
            c0[d[1]] = dist[d[0]] / (dist[d[0]] + dist[d[1]])
            c0s.append(c0)

        def make_pair(nu):
            T = m["B_precomp"].Pt_nu_P_xx(nu, v["out_t"], v["out_xx"])
            GD = m["B_optimizer"].grad_descent(nu, N, NI, NB, IC_u, v["out"], v["out_IC"], v["out_BC"], v["out_x"], T, lr)
            NNg = m["B_models"].GPT(nu, v["out"], v["out_IC"], v["out_BC"], v["out_x"], T, IC_u, BC_u, f_hat)
            return GD, NNg
        C, F, M, T = trajectories(make_pair, grid, lambda gi: c0s[gi], n, epochs, rec)
        L, case = m["B_train"].offline_generation(grid, N, NI, NB, IC_u, BC_u, v["out"], v["out_x"], v["out_t"],
                                                  v["out_xx"], v["out_IC"], v["out_BC"], f_hat, epochs, lr, neurons, n - 1)
        save[f"n{n}_c0"] = np.stack([c.numpy() for c in c0s])
        save[f"n{n}_c"], save[f"n{n}_f"], save[f"n{n}_m"], save[f"n{n}_trace"] = C, F, M, T
        save[f"n{n}_L"] = np.float32(float(L))
        save[f"n{n}_case"] = np.float64(case)
        print(tag, "n", n, "L", float(L), "case", case, flush=True)
    if not small:
        for k in ("out_test", "xt_test", "xt_resid", "idx_list"):
            save.pop(k, None)
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **save)
        print("wrote", tag)
        return
    test_nu = grid[[3, 11, 29]]
    times, soln = m["B_test"].gpt_test(test_nu, N, NI, NB, IC_u, BC_u, bufs["out"], bufs["out_x"], bufs["out_t"],
                                       bufs["out_xx"], bufs["out_IC"], bufs["out_BC"], f_hat, 1000, lr, neurons,
                                       bufs["out_test"])
    losses = m["B_test"].gpt_test_loss(test_nu, N, NI, NB, IC_u, BC_u, bufs["out"], bufs["out_x"], bufs["out_t"],
                                       bufs["out_xx"], bufs["out_IC"], bufs["out_BC"], f_hat, 1000, lr, neurons)
    save.update(test_mu=test_nu, test_epochs=1000, test_soln=soln, test_losses=losses)
    np.savez_compressed(os.path.join(OUT, "b_small.npz"), **save)
    print("wrote b_small")


# --------------------------------------------------------------------------------------
# Allen-Cahn (GPT / precompute path only: AC_main.py itself needs TensorFlow + pyDOE)
# --------------------------------------------------------------------------------------
def gen_ac_small():
    lm = np.linspace(1e-3, 1e-4, 5)
    ep = np.linspace(1, 5, 5)
    _gen_ac("ac_small", 800, 64, 20, [1, 3, 9], np.array(np.meshgrid(lm, ep)).T.reshape(-1, 2), 300, 50, True)   # AC_main.py:115-117 ordering


def gen_ac_full():
    """Reference default sizes (AC_main.py:49-51: N_R = 20 000, N_IC = 512, N_BC = 100 per side), n = 9, 2000 epochs."""
    grid = np.array([[1e-3, 1.0], [7.75e-4, 2.0], [5.5e-4, 4.0], [3.25e-4, 5.0], [1e-4, 3.0], [1e-4, 5.0]])
    _gen_ac("ac_full", 20000, 512, 100, [9], grid, 2000, 500, False)


def _gen_ac(tag, N, NI, NB, ns, grid, epochs, rec, small):
    import scipy.io
    m = _import_ref("Allen-Cahn", ["AC_models", "AC_precompute", "AC_optimizer", "AC_train", "AC_test"])
    rng = np.random.default_rng(0)
    nmax = 9
    data = scipy.io.loadmat(os.path.join(REF, "Allen-Cahn", "AC.mat"))
    x = data["x"].flatten()[:, None]
    t = data["tt"].flatten()[:, None]
    uu = data["uu"]
    idx_x = np.sort(rng.choice(x.shape[0], NI, replace=False))
    IC_xt = torch.tensor(np.hstack((x[idx_x], 0 * x[idx_x])), dtype=torch.float32)
    IC_u = torch.tensor(uu[idx_x, 0:1], dtype=torch.float32)
    tb = t[rng.choice(t.shape[0], NB, replace=False)]
    BC_xt_ub = torch.tensor(np.hstack((0 * tb + 1.0, tb)), dtype=torch.float32).requires_grad_()
    BC_xt_lb = torch.tensor(np.hstack((0 * tb - 1.0, tb)), dtype=torch.float32).requires_grad_()
    xt = torch.tensor(np.hstack((rng.uniform(-1, 1, (N, 1)), rng.uniform(0, 1, (N, 1)))), dtype=torch.float32).requires_grad_()
    f_hat = torch.zeros(N, 1)
    xg, tg = torch.meshgrid((torch.linspace(-1, 1, 12), torch.linspace(0, 1, 12)), indexing="ij")
    xt_test = torch.hstack((xg.transpose(1, 0).flatten().unsqueeze(1), tg.transpose(1, 0).flatten().unsqueeze(1)))
    NT = xt_test.shape[0]
    z = lambda r: torch.zeros(r, nmax)
    out, out_t, out_xx, out_IC, out_test = z(N), z(N), z(N), z(NI), z(NT)
    ub, lb, df, ubx, lbx, dfx = z(NB), z(NB), z(NB), z(NB), z(NB), z(NB)
    layers = [2, 128, 128, 128, 128, 1]
    save = {}
    for i in range(nmax):
        g = torch.Generator().manual_seed(3000 + i)
        w, b = [], []
        for l in range(len(layers) - 1):
            fi, fo = layers[l], layers[l + 1]
            std = (2.0 / (fi + fo)) ** 0.5
            w.append((torch.randn(fo, fi, generator=g) * std * 1.2).float())
            b.append((torch.randn(fo, generator=g) * 0.1).float())
        pinn = m["AC_models"].P(w, b, layers)
        if small and i < 2:
            for l in range(len(layers) - 1):
                save[f"pinn{i}_W{l}"] = pinn.linears[l].weight.data.numpy().copy()
                save[f"pinn{i}_b{l}"] = pinn.linears[l].bias.data.numpy().copy()
        m["AC_precompute"].inputs(pinn, xt, out, out_xx, out_t, out_IC, ub, lb, IC_xt, BC_xt_ub, BC_xt_lb, i,
                                  out_test, xt_test, f_hat, N, df, ubx, lbx, dfx)
    bufs = dict(out=out, out_t=out_t, out_xx=out_xx, out_IC=out_IC, out_test=out_test, out_BC_ub=ub, out_BC_lb=lb,
                out_BC_diff=df, out_BC_ub_x=ubx, out_BC_lb_x=lbx, out_BC_diff_x=dfx)
    bufs = {k: v.detach() for k, v in bufs.items()}
    save.update({k: v.numpy() for k, v in bufs.items()})
    save.update(xt_resid=xt.detach().numpy(), xt_test=xt_test.numpy(), IC_xt=IC_xt.numpy(),
                BC_xt_ub=BC_xt_ub.detach().numpy(), BC_xt_lb=BC_xt_lb.detach().numpy(), IC_u=IC_u.numpy(),
                f_hat=f_hat.numpy())
    lr = 0.0025
    save.update(grid=grid, ns=np.array(ns), epochs=epochs, lr=lr, rec_every=rec)
    names = ("out", "out_t", "out_xx", "out_IC", "out_BC_ub", "out_BC_lb", "out_BC_diff", "out_BC_ub_x",
             "out_BC_lb_x", "out_BC_diff_x")
    for n in ns:
        v = {k: bufs[k][:, :n] for k in names}
        c0 = torch.full((n,), 1.0 / n)

        def make_pair(mu):
            lam, eps = mu
            T = m["AC_precompute"].Pt_lPxx_eP(v["out_t"], v["out_xx"], v["out"], lam, eps)
            GD = m["AC_optimizer"].grad_descent(lam, eps, N, NI, NB, IC_u, T, v["out"], v["out_IC"], v["out_BC_ub"],
                                                v["out_BC_lb"], v["out_BC_diff"], v["out_BC_ub_x"], v["out_BC_lb_x"],
                                                v["out_BC_diff_x"], epochs, lr)
            NNg = m["AC_models"].GPT(lam, eps, v["out"], v["out_IC"], v["out_BC_ub"], v["out_BC_lb"],
                                     v["out_BC_ub_x"], v["out_BC_lb_x"], f_hat, T, IC_u)
            return GD, NNg
        C, F, M, T = trajectories(make_pair, grid, lambda gi: c0, n, epochs, rec)
        L, case, tc = m["AC_train"].offline_generation(grid, c0, N, NI, NB, IC_u, v["out"], v["out_t"], v["out_xx"],
                                                       v["out_IC"], v["out_BC_ub"], v["out_BC_lb"], v["out_BC_diff"],
                                                       v["out_BC_ub_x"], v["out_BC_lb_x"], v["out_BC_diff_x"], f_hat,
                                                       epochs, lr)
        save[f"n{n}_c"], save[f"n{n}_f"], save[f"n{n}_m"], save[f"n{n}_trace"] = C, F, M, T
        save[f"n{n}_L"] = np.float32(float(L))
        save[f"n{n}_case"] = np.array(case, np.float64)
        save[f"n{n}_trained_c"] = tc.numpy()
        print(tag, "n", n, "L", float(L), "case", case, flush=True)
    if not small:
        for k in ("out_test", "xt_test", "xt_resid", "IC_xt", "BC_xt_ub", "BC_xt_lb"):
            save.pop(k, None)
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **save)
        print("wrote", tag)
        return
    test_mu = grid[[2, 13, 21]]
    c0 = torch.full((nmax,), 1.0 / nmax)
    args = (bufs["out"], bufs["out_t"], bufs["out_xx"], bufs["out_IC"], f_hat, N, NI, NB, IC_u, c0, 400, lr,
            bufs["out_test"], bufs["out_BC_ub"], bufs["out_BC_lb"], bufs["out_BC_diff"], bufs["out_BC_ub_x"],
            bufs["out_BC_lb_x"], bufs["out_BC_diff_x"])
    times, soln = m["AC_test"].gpt_test(test_mu, *args)
    losses = m["AC_test"].gpt_test_loss(test_mu, *args)
    save.update(test_mu=test_mu, test_epochs=400, test_soln=soln, test_losses=losses)
    np.savez_compressed(os.path.join(OUT, "ac_small.npz"), **save)
    print("wrote ac_small")


def gen_kg_greedy():
    """The greedy outer loop of KG_main.py:100-138 with the full-PINN training replaced by fixed synthetic
    weights (training is chaotic and not reproducible across implementations; everything else -- inputs(),
    the grid bookkeeping, offline_generation and the neuron hand-over -- is the reference's own code)."""
    nmax = 6
    m, p = kg_problem(24, 64, 64, 10, nmax)           # same point sets / weights as kg_small's first six neurons
    N, NI, NB = p["out"].shape[0], p["out_IC"].shape[0], p["out_BC"].shape[0]
    kg_train = kg_grid(4, 3, 5)
    neurons = np.zeros((nmax, 3))
    neurons[0] = (np.median(np.linspace(-2, -1, 4)), np.median(np.linspace(0, 1, 3)), np.median(np.linspace(0, 1, 5)))
    loss_list = np.zeros(nmax)
    sizes = []
    for i, neuron in enumerate(neurons):
        kg_train = np.delete(kg_train, np.where(np.all(kg_train == neuron, axis=1))[0], axis=0)      # KG_main.py:102
        sizes.append(len(kg_train))
        n = i + 1
        v = {k: p[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
        L, case = m["KG_train"].offline_generation(
            kg_train, torch.full((n,), 1.0 / n), N, NI, NB, p["IC_u1"], p["IC_u2"], p["BC_u"], p["xcos"],
            v["out"], v["out_xx"], v["out_tt"], v["out_IC"], v["out_IC_t"], v["out_BC"], p["f_hat"], 200, 0.025)
        loss_list[i] = L
        if i + 1 < nmax:
            neurons[i + 1] = case
        print("kg_greedy stage", i, "L", float(L), "case", case, flush=True)
    np.savez_compressed(os.path.join(OUT, "kg_greedy.npz"), neurons=neurons, loss_list=loss_list, sizes=np.array(sizes),
                        epochs=200, lr=0.025, nmax=nmax)
    print("wrote kg_greedy")


# --------------------------------------------------------------------------------------
# single-step classes: grad_descent.update / GPT.loss(+ partial losses) of all three PDEs, as user code drives them
# --------------------------------------------------------------------------------------
def gen_classes():
    """One reference `update` and every reference loss method on the matrices of the *_small fixtures (loaded back
    from tests/golden), for a few parameters and neuron counts, from seeded random c."""
    save = {}
    rng = np.random.default_rng(99)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    # ---- Klein-Gordon (KG_optimizer.py:4-68, KG_models.py:7-50)
    m = _import_ref("Klein-Gordon", ["KG_models", "KG_precomp", "KG_optimizer"])
    g = np.load(os.path.join(OUT, "kg_small.npz"))
    N, NI, NB = g["out"].shape[0], g["out_IC"].shape[0], g["out_BC"].shape[0]
    for n in (2, 7, 15):
        v = {k: t(g[k])[:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
        mus = g["grid"][[5, 33, 59]]
        cin = (rng.standard_normal((len(mus), n)) * 0.4).astype(np.float32)
        cout, L = np.zeros_like(cin), np.zeros((len(mus), 5), np.float32)
        for k, (a, b, gm) in enumerate(mus):
            T = m["KG_precomp"].Ptt_aPxx_bP(a, b, v["out_tt"], v["out_xx"], v["out"])
            G2 = m["KG_precomp"].gamma2_P(gm, v["out"])
            GD = m["KG_optimizer"].grad_descent(a, b, gm, N, NI, NB, t(g["IC_u1"]), t(g["IC_u2"]), t(g["BC_u"]), t(g["xcos"]), T, G2,
                                                v["out"], v["out_IC"], v["out_IC_t"], v["out_BC"], 1, 0.025)
            NNg = m["KG_models"].GPT(a, b, gm, v["out"], v["out_IC"], v["out_IC_t"], v["out_BC"], t(g["xcos"]), t(g["f_hat"]), T,
                                     t(g["IC_u1"]), t(g["IC_u2"]), t(g["BC_u"]))
            c = t(cin[k])
            cout[k] = GD.update(c, c.unsqueeze(1)).numpy()
            cr = c.unsqueeze(1)
            L[k] = [NNg.loss(cr).item(), NNg.lossR(cr).item(), NNg.lossIC1(cr).item(), NNg.lossIC2(cr).item(), NNg.lossBC(cr).item()]
        save.update({f"kg_n{n}_mu": mus, f"kg_n{n}_cin": cin, f"kg_n{n}_cout": cout, f"kg_n{n}_loss": L})
    # ---- Burgers (B_optimizer.py:4-58, B_models.py:11-46)
    m = _import_ref("Burgers", ["B_models", "B_precomp", "B_optimizer"])
    g = np.load(os.path.join(OUT, "b_small.npz"))
    N, NI, NB = g["out"].shape[0], g["out_IC"].shape[0], g["out_BC"].shape[0]
    for n in (2, 9):
        v = {k: t(g[k])[:, :n] for k in ("out", "out_t", "out_x", "out_xx", "out_IC", "out_BC")}
        mus = g["grid"][[1, 17, 30]]
        cin = (rng.standard_normal((len(mus), n)) * 0.4).astype(np.float32)
        cout, L = np.zeros_like(cin), np.zeros((len(mus), 4), np.float32)
        for k, nu in enumerate(mus):
            T = m["B_precomp"].Pt_nu_P_xx(nu, v["out_t"], v["out_xx"])
            GD = m["B_optimizer"].grad_descent(nu, N, NI, NB, t(g["IC_u"]), v["out"], v["out_IC"], v["out_BC"], v["out_x"], T, 0.02)
            NNg = m["B_models"].GPT(nu, v["out"], v["out_IC"], v["out_BC"], v["out_x"], T, t(g["IC_u"]), t(g["BC_u"]), t(g["f_hat"]))
            c = t(cin[k])
            cout[k] = GD.update(c, c.unsqueeze(1)).numpy()
            cr = c.unsqueeze(1)
            L[k] = [NNg.loss(cr).item(), NNg.lossR(cr).item(), NNg.lossIC(cr).item(), NNg.lossBC(cr).item()]
        save.update({f"b_n{n}_mu": mus, f"b_n{n}_cin": cin, f"b_n{n}_cout": cout, f"b_n{n}_loss": L})
    # ---- Allen-Cahn (AC_optimizer.py:4-66, AC_models.py:6-54)
    m = _import_ref("Allen-Cahn", ["AC_models", "AC_precompute", "AC_optimizer"])
    g = np.load(os.path.join(OUT, "ac_small.npz"))
    N, NI, NB = g["out"].shape[0], g["out_IC"].shape[0], g["out_BC_ub"].shape[0]
    names = ("out", "out_t", "out_xx", "out_IC", "out_BC_ub", "out_BC_lb", "out_BC_diff", "out_BC_ub_x", "out_BC_lb_x", "out_BC_diff_x")
    for n in (3, 9):
        v = {k: t(g[k])[:, :n] for k in names}
        mus = g["grid"][[0, 12, 24]]
        cin = (rng.standard_normal((len(mus), n)) * 0.4).astype(np.float32)
        cout, L = np.zeros_like(cin), np.zeros((len(mus), 5), np.float32)
        for k, (lam, eps) in enumerate(mus):
            T = m["AC_precompute"].Pt_lPxx_eP(v["out_t"], v["out_xx"], v["out"], lam, eps)
            GD = m["AC_optimizer"].grad_descent(lam, eps, N, NI, NB, t(g["IC_u"]), T, v["out"], v["out_IC"], v["out_BC_ub"], v["out_BC_lb"],
                                                v["out_BC_diff"], v["out_BC_ub_x"], v["out_BC_lb_x"], v["out_BC_diff_x"], 1, 0.0025)
            NNg = m["AC_models"].GPT(lam, eps, v["out"], v["out_IC"], v["out_BC_ub"], v["out_BC_lb"], v["out_BC_ub_x"], v["out_BC_lb_x"],
                                     t(g["f_hat"]), T, t(g["IC_u"]))
            c = t(cin[k])
            cout[k] = GD.update(c, c.unsqueeze(1)).numpy()
            cr = c.unsqueeze(1)
            L[k] = [NNg.loss(cr).item(), NNg.lossR(cr).item(), NNg.lossIC(cr).item(), NNg.lossBC(cr).item(), NNg.lossBC_x(cr).item()]
        save.update({f"ac_n{n}_mu": mus, f"ac_n{n}_cin": cin, f"ac_n{n}_cout": cout, f"ac_n{n}_loss": L})
    np.savez_compressed(os.path.join(OUT, "classes.npz"), **save)
    print("wrote classes")


# --------------------------------------------------------------------------------------
# full-PINN training step: the reference's NN.loss + torch.optim.Adam
# --------------------------------------------------------------------------------------
def _flat(params):
    return np.concatenate([p.detach().numpy().reshape(-1) for p in params])


def _pinn_fixture(make, loss_of, lr, steps, tol=None):
    """loss, all gradients, weights after one Adam step, and a `steps`-long loss trace (loss BEFORE each step, as the
    reference's loop evaluates it); plus the same loss / gradient in float64 (autograd in double on the same
    weights) so the tests can state how far fp32 autograd itself sits from the exact value."""
    net = make()
    params = [q for lin in net.linears for q in (lin.weight, lin.bias)]
    w0 = _flat(params)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    loss = loss_of(net, torch.float32)
    opt.zero_grad()
    loss.backward()
    g32 = _flat([q.grad for q in params])
    opt.step()
    w1 = _flat(params)
    trace = [loss.item()]
    for _ in range(steps - 1):
        loss = loss_of(net, torch.float32)
        if tol is not None and loss < tol:
            break
        trace.append(loss.item())
        opt.zero_grad()
        loss.backward()
        opt.step()
    wN = _flat(params)
    final = loss_of(net, torch.float32).item()
    net64 = make().double()
    loss64 = loss_of(net64, torch.float64)
    loss64.backward()
    g64 = _flat([q.grad for lin in net64.linears for q in (lin.weight, lin.bias)])
    return dict(w0=w0, loss=np.float32(trace[0]), grad=g32, w1=w1, trace=np.array(trace, np.float32), wN=wN,
                final_loss=np.float32(final), loss64=np.float64(loss64.item()), grad64=g64)


def gen_pinn_step():
    save = {}
    # ---- Klein-Gordon: NN([2,40,40,1], cos), the constructor's own seed-1234 Xavier init (KG_models.py:62-82)
    m = _import_ref("Klein-Gordon", ["KG_data", "KG_models", "KG_precomp"])
    for tag, (Nc, ICp, BCp), steps in (("kg_small", (24, 64, 64), 200), ("kg_default", (100, 512, 512), 40)):
        xt, f_hat, _ = m["KG_data"].residual_data(-1.0, 1.0, 0.0, 5.0, Nc, 4)
        IC_xt, IC_u1, IC_u2, BC_xt, BC_u = m["KG_data"].ICBC_data(-1.0, 1.0, 0.0, 5.0, BCp, ICp)
        xcos = m["KG_precomp"].xcos_term(xt[:, 0].unsqueeze(1), xt[:, 1].unsqueeze(1))
        mu = (-1.3, 0.4, 0.7)

        def make():
            return m["KG_models"].NN(np.array([2, 40, 40, 1]), *mu, xcos)

        def loss_of(net, dt):
            net.xcos_x2cos2 = xcos.to(dt)
            a = xt.to(dt).clone().requires_grad_()
            b = IC_xt.to(dt).clone().requires_grad_()
            return net.loss(a, f_hat.to(dt), b, IC_u1.to(dt), IC_u2.to(dt), BC_xt.to(dt), BC_u.to(dt))
        r = _pinn_fixture(make, loss_of, 5e-4, steps)
        save.update({f"{tag}_{k}": v for k, v in r.items()})
        save[f"{tag}_sizes"] = np.array([Nc, ICp, BCp])
        save[f"{tag}_mu"] = np.array(mu)
        save[f"{tag}_lr"] = 5e-4
        print(tag, "loss", r["loss"], "loss64", r["loss64"], "|g32-g64|max/|g|max", np.abs(r["grad"] - r["grad64"]).max() / np.abs(r["grad64"]).max(), flush=True)
    # ---- Burgers: NN([2,20,20,20,20,1], tanh) (B_models.py:52-105), lr 5e-3, tolerance semantics of B_train.py:84
    m = _import_ref("Burgers", ["B_data", "B_models"])
    for tag, (Nc, ICp, BCp), steps in (("b_small", (24, 40, 40), 200), ("b_default", (100, 200, 200), 40)):
        xt, f_hat, _ = m["B_data"].residual_data(-1.0, 1.0, 0.0, 1.0, Nc, 4)
        IC_xt, IC_u, BC_xt, BC_u = m["B_data"].ICBC_data(-1.0, 1.0, 0.0, 1.0, BCp, ICp)
        nu = 0.05

        def make():
            return m["B_models"].NN(np.array([2, 20, 20, 20, 20, 1]), nu)

        def loss_of(net, dt):
            a = xt.to(dt).clone().requires_grad_()
            return net.loss(a, IC_xt.to(dt), IC_u.to(dt), BC_xt.to(dt), BC_u.to(dt), f_hat.to(dt))
        r = _pinn_fixture(make, loss_of, 5e-3, steps)
        save.update({f"{tag}_{k}": v for k, v in r.items()})
        save[f"{tag}_sizes"] = np.array([Nc, ICp, BCp])
        save[f"{tag}_mu"] = np.array([nu])
        save[f"{tag}_lr"] = 5e-3
        print(tag, "loss", r["loss"], "loss64", r["loss64"], "|g32-g64|max/|g|max", np.abs(r["grad"] - r["grad64"]).max() / np.abs(r["grad64"]).max(), flush=True)
    np.savez_compressed(os.path.join(OUT, "pinn_step.npz"), **save)
    print("wrote pinn_step")


GENS = {"kg_small": gen_kg_small, "kg_nonmono": gen_kg_nonmono, "kg_full": gen_kg_full, "b_full": gen_b_full, "ac_full": gen_ac_full,
        "b_small": gen_b_small, "ac_small": gen_ac_small, "kg_greedy": gen_kg_greedy, "classes": gen_classes, "pinn_step": gen_pinn_step}

if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("OMP_NUM_THREADS", "8")))
    which = sys.argv[1:] or ["all"]
    os.makedirs(OUT, exist_ok=True)
    for name, fn in GENS.items():
        if "all" in which or name in which:
            fn()
