"""Drop-in for Burgers/B_train.py.

offline_generation: reference B_train.py:12-71.  The per-parameter initial coefficients
(B_train.py:27-47: one-hot if nu is a neuron, else inverse-distance weights on the two nearest
neuron parameters) are built on the host for the whole grid, then one fused launch trains every
parameter and the exact scan picks (largest_loss, largest_case).
pinn_train: reference B_train.py:76-90 (Adam with a tolerance checked before each step).
"""
import numpy as np
import torch

from _common import offline, pinn, sweep

device = torch.device("cuda")


def initial_coefficients(nu_grid, neuron_params, neuron_cnt, allow_one_hot=True):
    """[Np, n] float32 initial c exactly as the reference loop builds it (float64 math, fp32 store)."""
    nu_grid = np.asarray(nu_grid, dtype=np.float64).reshape(-1)
    c0 = np.zeros((len(nu_grid), neuron_cnt), dtype=np.float32)
    if neuron_cnt == 1:
        c0[:] = 1.0
        return c0
    neuron_params = np.asarray(neuron_params, dtype=np.float64).reshape(-1)
    for p, nu in enumerate(nu_grid):
        if allow_one_hot and nu in neuron_params:
            c0[p, np.where(nu == neuron_params)] = 1
            continue
        dist = np.abs(neuron_params[:neuron_cnt] - nu)
        d = np.argsort(dist)
        first, second = d[0], d[1]
        a, b = dist[first], dist[second]
        bottom = a + b
        c0[p, first] = b / bottom
        c0[p, second] = a / bottom
    return c0


def offline_generation(nu_train, xt_size, IC_size, BC_size, IC_u, BC_u, out,
                       out_x, out_t, out_xx, out_IC, out_BC, f_hat, epochs_gpt,
                       lr_gpt, neurons, i):
    neuron_params = neurons[:i + 1]
    neuron_cnt = out.shape[1]
    grid = np.asarray(nu_train, dtype=np.float64).reshape(-1)
    c0 = torch.from_numpy(initial_coefficients(grid, neuron_params, neuron_cnt)).to(out.device)

    def run(rows, c0_rows, epochs=epochs_gpt, **kw):
        return sweep.b_sweep(rows, c0_rows, out, out_x, out_t, out_xx, out_IC, out_BC, f_hat, IC_u, BC_u,
                             epochs, lr_gpt, want_c=False, **kw)

    largest_loss, idx, _ = offline.greedy_select(run, grid, epochs_gpt, c0=c0)
    if idx < 0:
        return 0, 0
    return largest_loss, nu_train[idx]


def pinn_train(PINN, nu, xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat,
               epochs_pinn, lr_pinn, tol):
    r = pinn.train_fused(PINN, "burgers", (nu,), xt_resid, f_hat, IC_xt, IC_u, None, BC_xt, BC_u,
                         epochs_pinn, lr_pinn, tol=tol)
    print(f"PINN Final Loss: {r['loss'].item()}")
