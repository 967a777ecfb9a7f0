"""Drop-in for Burgers/B_test.py.

gpt_test / gpt_test_loss: reference B_test.py:13-63, 68-111 -- inverse-distance initial c over all
neurons, 5000 updates per test parameter, fused into one launch; losses every 500 epochs.
pinn_test / pinn_test_loss: reference B_test.py:116-181 (tolerance-stopped Adam per parameter).
"""
import time

import numpy as np
import torch

from B_models import NN
from B_train import initial_coefficients
from _common import offline, sweep

device = torch.device("cuda")


def _c0(nu_train, neurons, neuron_cnt, dev):
    return torch.from_numpy(initial_coefficients(nu_train, neurons, neuron_cnt, allow_one_hot=False)).to(dev)


def gpt_test(nu_train, xt_size, IC_size, BC_size, IC_u, BC_u, out,
             out_x, out_t, out_xx, out_IC, out_BC, f_hat, epochs_gpt,
             lr_gpt, neurons, out_test):
    t0 = time.time()
    grid = np.asarray(nu_train, dtype=np.float64).reshape(-1)
    c0 = _c0(grid, neurons, out.shape[1], out.device)

    def run(rows, c0_rows):
        return sweep.b_sweep(rows, c0_rows, out, out_x, out_t, out_xx, out_IC, out_BC, f_hat, IC_u, BC_u,
                             epochs_gpt, lr_gpt, want_c=True)

    res = offline.sharded_sweep(run, grid, c0=c0, want_c=True, gather_c=True)
    gpt_soln = torch.matmul(out_test, res["c"].t()).cpu().numpy().astype(np.float64)
    return offline.spread_hours(time.time() - t0, len(grid)), gpt_soln


def gpt_test_loss(nu_train, xt_size, IC_size, BC_size, IC_u, BC_u, out,
                  out_x, out_t, out_xx, out_IC, out_BC, f_hat, epochs_gpt,
                  lr_gpt, neurons):
    grid = np.asarray(nu_train, dtype=np.float64).reshape(-1)
    c0 = _c0(grid, neurons, out.shape[1], out.device)
    lo, hi = offline.shard_bounds(len(grid))
    r = sweep.b_sweep(grid[lo:hi], c0[lo:hi], out, out_x, out_t, out_xx, out_IC, out_BC, f_hat, IC_u, BC_u,
                      epochs_gpt, lr_gpt, rec_every=500, want_c=False)
    tr = offline.gather_rows(r["trace"], len(grid)).t().cpu().numpy().astype(np.float64)
    losses = np.zeros((11, len(grid)))
    losses[:min(11, tr.shape[0])] = tr[:11]
    return losses


def pinn_test(b_test, layers_pinn, xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat,
              epochs_pinn, lr_pinn, xt_test, tol):
    times = np.zeros(len(b_test))
    pinn_soln = np.zeros((xt_test.shape[0], len(b_test)))
    for i, nu in enumerate(b_test):
        t_start = time.time()
        PINN = NN(layers_pinn, nu).to(device)
        optimizer = torch.optim.Adam(PINN.parameters(), lr=lr_pinn)
        for j in range(1, epochs_pinn + 1):
            loss = PINN.loss(xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat)
            if loss < tol:
                break
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
        with torch.no_grad():
            soln = PINN(xt_test)
        dt = (time.time() - t_start) / 3600
        times[i] = dt if i == 0 else dt + times[i - 1]
        pinn_soln[:, i][:, None] = soln.detach().cpu().numpy()
    return times, pinn_soln


def pinn_test_loss(b_test, layers_pinn, xt_resid, IC_xt, IC_u, BC_xt, BC_u,
                   f_hat, epochs_pinn, lr_pinn, tol):
    losses = np.zeros((61, len(b_test)))
    for i, nu in enumerate(b_test):
        PINN = NN(layers_pinn, nu).to(device)
        optimizer = torch.optim.Adam(PINN.parameters(), lr=lr_pinn)
        losses[0, i] = PINN.loss(xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat).item()
        for j in range(1, epochs_pinn + 1):
            loss = PINN.loss(xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat)
            if (j % 1000 == 0) or (j == epochs_pinn):
                losses[int(j / 1000), i] = loss.item()
            if loss < tol:
                losses[np.where(losses[:, i] == 0)[0][0], i] = loss.item()
                break
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
    return losses
