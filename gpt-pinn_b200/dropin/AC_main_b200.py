#!/usr/bin/env python
"""Allen-Cahn GPT-PINN driver without TensorFlow / pyDOE: the flow of the reference's Allen-Cahn/AC_main.py on the B200 path.

The reference's AC_main.py imports tensorflow, keras and pyDOE at the top (AC_main.py:3-18) and trains its self-adaptive
PINNs with them (AC_main.py:204-339), so -- unlike KG_main.py / B_main.py -- it cannot run unchanged in an image without
those packages.  This driver keeps its data flow, hyper-parameters, greedy loop, test phase and result files
(AC_main.py:44-150, 340-468) and replaces only the TensorFlow part by gptpinn.sa_pinn.SAPinn (same loss, same two Adam
optimizers with ascent on the self-adaptive weights, same fixed-step L-BFGS); everything downstream goes through the
drop-in modules AC_models.P, AC_precompute.inputs, AC_train.offline_generation, AC_test.gpt_test / gpt_test_loss.

  python AC_main_b200.py [--mat AC.mat] [--out ./ac_data] [--neurons 9] [--adam 10000] [--lbfgs 10000] [--seed 0]

Not reproducible bit for bit against the reference (Keras' Glorot seed stream, tf.random.uniform and pyDOE's lhs are not
available here): the collocation points are a seeded Latin hypercube of the same shape, the initial weights a seeded
truncated-normal Glorot draw, the self-adaptive weights seeded uniforms -- statistical parity only.
"""
import argparse
import os
import sys
import time
from datetime import datetime

import numpy as np
import scipy.io
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from AC_models import P                      # noqa: E402
from AC_precompute import inputs             # noqa: E402
from AC_test import gpt_test, gpt_test_loss  # noqa: E402
from AC_train import offline_generation      # noqa: E402
from _common import gptpinn                  # noqa: E402,F401
from gptpinn.sa_pinn import SAPinn           # noqa: E402


def latin_hypercube(n, dims, rng):
    """pyDOE lhs(dims, samples=n) with its default criterion: one uniform draw per stratum, strata permuted per dimension."""
    cut = np.linspace(0.0, 1.0, n + 1)
    u = rng.random((n, dims))
    pts = cut[:n, None] + u * (cut[1:, None] - cut[:n, None])
    for j in range(dims):
        pts[:, j] = pts[rng.permutation(n), j]
    return pts


def find_mat(path):
    for cand in (path, os.environ.get("GPTPINN_AC_MAT"), os.path.join(os.getcwd(), "AC.mat"), os.path.join(HERE, "AC.mat")):
        if cand and os.path.exists(cand):
            return cand
    raise SystemExit("AC.mat not found (the reference ships it as Allen-Cahn/AC.mat): pass --mat or set GPTPINN_AC_MAT")


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--mat", default=None)
    ap.add_argument("--out", default="./ac_data")
    ap.add_argument("--neurons", type=int, default=9)
    ap.add_argument("--adam", type=int, default=10000)           # epochs_adam_sa  (AC_main.py:122)
    ap.add_argument("--lbfgs", type=int, default=10000)          # epochs_lbfgs_sa (AC_main.py:123)
    ap.add_argument("--epochs-gpt", type=int, default=2000)
    ap.add_argument("--epochs-gpt-test", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)

    device = torch.device("cuda")
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    rng = np.random.default_rng(args.seed)
    os.makedirs(args.out, exist_ok=True)
    sep = 60 * "*"
    print(f"Start: {datetime.now()}\n")

    # ---- domain and data (AC_main.py:47-106)
    N0, N_b, N_f = 512, 100, 20000
    data = scipy.io.loadmat(find_mat(args.mat))
    t = data["tt"].flatten()[:, None]
    x = data["x"].flatten()[:, None]
    exact_u = data["uu"]
    idx_x = np.random.choice(x.shape[0], N0, replace=False)
    X0 = np.concatenate((x[idx_x, :], 0 * x[idx_x, :]), 1)
    u0 = exact_u[idx_x, 0:1]
    tb = t[np.random.choice(t.shape[0], N_b, replace=False), :]
    X_lb = np.concatenate((0 * tb - 1.0, tb), 1)
    X_ub = np.concatenate((0 * tb + 1.0, tb), 1)
    X_f = -1.0 + 2.0 * latin_hypercube(N_f, 2, rng)
    X_f[:, 1] = np.abs(X_f[:, 1])
    T = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=device)
    sa_xt_f, sa_xt0, sa_u0, sa_lb, sa_ub = T(X_f), T(X0), T(u0), T(X_lb), T(X_ub)

    # GPT side: IC sorted by x (AC_main.py:84-88), BC and collocation sets as the SA-PINN's
    order = np.argsort(X0[:, 0], kind="stable")
    IC_u = T(u0[order])
    IC_xt = T(np.hstack((X0[order, 0:1], X0[:, 1:2])))
    BC_xt_ub, BC_xt_lb, xt_resid = T(X_ub), T(X_lb), T(X_f)
    f_hat = torch.zeros(N_f, 1, device=device)
    xg, tg = torch.meshgrid((torch.linspace(-1, 1, 60), torch.linspace(0, 1, 60)), indexing="ij")
    xt_test = torch.hstack((xg.transpose(1, 0).flatten().unsqueeze(1), tg.transpose(1, 0).flatten().unsqueeze(1))).to(device)

    lmbda, eps = np.linspace(1e-3, 1e-4, 11), np.linspace(1, 5, 11)
    ac_train = np.array(np.meshgrid(lmbda, eps)).T.reshape(-1, 2)
    ac_train_all = ac_train.copy()
    layers_sa = [2, 128, 128, 128, 128, 1]
    n_neurons, lr_gpt = args.neurons, 0.0025
    test_cases = int(np.ceil(0.2 * len(ac_train)))
    loss_list, neurons = np.zeros(n_neurons), np.zeros((n_neurons, 2))
    neurons[0] = (np.median(lmbda), np.median(eps))
    c_init = [torch.full((1, i + 1), 1 / (i + 1)).to(device) for i in range(n_neurons)]

    test_size, xt_size, IC_size, BC_size = xt_test.shape[0], xt_resid.shape[0], IC_xt.shape[0], BC_xt_ub.shape[0]
    z = lambda rows: torch.zeros((rows, n_neurons), device=device)
    out_full, out_t_full, out_xx_full, out_IC = z(xt_size), z(xt_size), z(xt_size), z(IC_size)
    out_BC_ub, out_BC_lb, out_BC_diff = z(BC_size), z(BC_size), z(BC_size)
    out_BC_ub_x, out_BC_lb_x, out_BC_diff_x = z(BC_size), z(BC_size), z(BC_size)
    out_test = z(test_size)
    generation_time, sa_time, sa_final = np.zeros(n_neurons), np.zeros(n_neurons), np.zeros((n_neurons, 4))
    exact_err = float("nan")

    # ---- greedy loop (AC_main.py:340-410)
    t_total = time.time()
    for i, neuron in enumerate(neurons):
        print(sep)
        ac_train = np.delete(ac_train, np.where(np.all(ac_train == neuron, axis=1))[0], axis=0)
        lam_i, eps_i = neuron
        t1 = time.time()
        sa = SAPinn(layers_sa, sa_xt_f, sa_xt0, sa_u0, sa_lb, sa_ub, float(lam_i), float(eps_i), seed=args.seed + i)
        sa.fit(args.adam, args.lbfgs, verbose=True)
        torch.cuda.synchronize()
        sa_time[i] = (time.time() - t1) / 60
        sa_final[i] = sa.loss_and_grad().tolist()
        print(f"PINN time: {sa_time[i]} minutes\n")
        if abs(lam_i - 1e-4) < 1e-12 and abs(eps_i - 5.0) < 1e-12:
            # AC.mat holds the spectral solution for (lambda, eps) = (1e-4, 5): the one neuron whose full-order solution is known
            Xg, Tg = np.meshgrid(x.flatten(), t.flatten(), indexing="ij")
            pred = sa.predict(T(np.stack((Xg.flatten(), Tg.flatten()), 1))).cpu().numpy().reshape(exact_u.shape)
            exact_err = float(np.linalg.norm(pred - exact_u) / np.linalg.norm(exact_u))
            print(f"SA-PINN relative L2 error against the exact solution of AC.mat (lambda=1e-4, eps=5): {exact_err}\n")
        SA_w, SA_b = sa.weights_for_P()
        PINN = P(SA_w, SA_b, layers_sa).to(device)

        (train_out, train_out_xx, train_out_t, train_out_IC, train_out_BC_ub, train_out_BC_lb, train_out_BC_diff, fhat, train_len,
         train_out_BC_ub_x, train_out_BC_lb_x, train_out_BC_diff_x) = inputs(
            PINN, xt_resid, out_full, out_xx_full, out_t_full, out_IC, out_BC_ub, out_BC_lb, IC_xt, BC_xt_ub, BC_xt_lb, i, out_test,
            xt_test, f_hat, xt_size, out_BC_diff, out_BC_ub_x, out_BC_lb_x, out_BC_diff_x)

        t1 = time.time()
        largest_loss, largest_case, trained_c = offline_generation(
            ac_train, c_init[i][0], train_len, IC_size, BC_size, IC_u, train_out, train_out_t, train_out_xx, train_out_IC, train_out_BC_ub,
            train_out_BC_lb, train_out_BC_diff, train_out_BC_ub_x, train_out_BC_lb_x, train_out_BC_diff_x, fhat, args.epochs_gpt, lr_gpt)
        generation_time[i] = (time.time() - t1) / 60
        print(f"Generation time: {generation_time[i]} minutes")
        loss_list[i] = largest_loss
        if i + 1 < n_neurons:
            neurons[i + 1] = largest_case
        print(f"\nLargest Loss (Using {i + 1} Neurons): {largest_loss}")
        print(f"Parameter Case: {largest_case}\n")
    total_time = (time.time() - t_total) / 3600
    print(sep)
    print(f"Total Training Time: {total_time} Hours\n")
    print(f"Activation Function Parameters: \n{neurons}\n")
    print(f"Loss list: {loss_list}\n")

    # ---- test phase (AC_main.py:420-436)
    ac_test = ac_train[np.random.choice(len(ac_train), test_cases, replace=False)]
    c_initial = c_init[n_neurons - 1][0]
    print("SGPT-PINN Testing Started")
    targs = (ac_test, train_out, train_out_t, train_out_xx, train_out_IC, fhat, train_len, IC_size, BC_size, IC_u, c_initial,
             args.epochs_gpt_test, lr_gpt, out_test, train_out_BC_ub, train_out_BC_lb, train_out_BC_diff, train_out_BC_ub_x,
             train_out_BC_lb_x, train_out_BC_diff_x)
    test_gpt_time, test_gpt_soln = gpt_test(*targs)
    test_gpt_losses = gpt_test_loss(*targs)
    print("SGPT-PINN Testing Ended\n")

    # ---- result files: the names and layouts of AC_main.py:438-466
    sv = lambda name, a: np.savetxt(os.path.join(args.out, name), a)
    sv("generation_time.dat", generation_time); sv("loss_list.dat", loss_list); sv("neurons.dat", neurons)
    sv("total_time.dat", np.array([total_time])); sv("xt_resid.dat", xt_resid.detach().cpu().numpy()); sv("ac_test.dat", ac_test)
    sv("xt_test.dat", xt_test.cpu().numpy()); sv("test_gpt_losses.dat", test_gpt_losses); sv("test_gpt_soln.dat", test_gpt_soln)
    sv("test_gpt_time.dat", test_gpt_time + total_time)
    sv("sa_pinn_rel_l2_vs_exact.dat", np.array([exact_err]))
    sv("sa_pinn_time.dat", sa_time); sv("sa_pinn_final_losses.dat", sa_final)      # extra: per-neuron SA-PINN minutes, (total, mse_u, mse_b, mse_f)
    params = {"device": device, "domain": {"xi": -1, "xf": 1, "ti": 0, "tf": 1},
              "data sizes": {"N_f": N_f, "N_test": xt_size, "BC_pts": BC_size, "N0": IC_size}, "layers_pinn": layers_sa,
              "lr_adam_sa": 0.005, "lr_lbfgs_sa": 0.8, "epochs_adam_sa": args.adam, "epochs_lbfgs_sa": args.lbfgs,
              "parameter size": len(ac_train_all), "number_of_neurons": n_neurons, "lr_gpt": lr_gpt, "epochs gpt train": args.epochs_gpt,
              "test cases": test_cases, "epochs gpt test": args.epochs_gpt_test, "seed": args.seed}
    np.save(os.path.join(args.out, "params.npy"), params)
    print(f"End: {datetime.now()}\n")
    return dict(neurons=neurons, loss_list=loss_list, total_time=total_time, sa_time=sa_time, sa_final=sa_final, exact_err=exact_err)


if __name__ == "__main__":
    main()
