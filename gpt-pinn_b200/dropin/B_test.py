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
from _common import offline, pinn, sweep

device = torch.device("cuda")
PINN_BATCH = 8          # test parameters per training launch (1..8)


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


def _chunks(n):
    """Test parameters trained together in one launch (gptpinn_pinn_train_batch): as few launches as PINN_BATCH allows,
    of near-equal size (25 parameters -> 7, 6, 6, 6)."""
    b = max(1, min(pinn.MAX_BATCH, PINN_BATCH))
    k = max(1, -(-n // b))
    cuts = [round(i * n / k) for i in range(k + 1)]
    return [(cuts[i], cuts[i + 1]) for i in range(k) if cuts[i + 1] > cuts[i]]


def _trained(mine, layers_pinn, xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn, tol):
    """The tolerance-stopped full PINNs of this rank's test parameters, trained once per (parameters, data, epochs, lr, tol):
    pinn_test and pinn_test_loss ask for the same trainings (see gptpinn.pinn: memo of finished test-set trainings).  Per
    parameter: the trained network, the iterations entered, and the loss at the top of every iteration."""
    chunks = _chunks(len(mine))
    tensors = (xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat)
    key = pinn.train_memo_key("burgers", layers_pinn, mine, epochs_pinn, lr_pinn, tol, chunks, tensors)
    recs = pinn.train_memo_get(key)
    if recs is not None:
        return recs
    recs = []
    for a, b in chunks:
        nets = [NN(layers_pinn, mine[k]).to(device) for k in range(a, b)]
        res = pinn.train_fused_batch(nets, "burgers", [(mine[k],) for k in range(a, b)], xt_resid, f_hat, IC_xt, IC_u, None,
                                     BC_xt, BC_u, epochs_pinn, lr_pinn, tol=tol, trace_every=1)
        for net, r in zip(nets, res):
            done = pinn.epochs_done(r)                                   # iterations entered
            recs.append(dict(net=net, done=done, trace=r["trace"][:max(done, 1)].cpu().numpy().astype(np.float64)))
    pinn.train_memo_put(key, tensors, recs)
    return recs


def pinn_test(b_test, layers_pinn, xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat,
              epochs_pinn, lr_pinn, xt_test, tol):
    """One tolerance-stopped full PINN per test parameter (B_test.py:116-149), fused training kernel, several
    parameters per launch; under torch.distributed each rank trains a contiguous block of the parameters."""
    b_test = np.asarray(b_test, dtype=np.float64).reshape(-1)
    t0 = time.time()
    lo, hi = offline.shard_bounds(len(b_test))
    recs = _trained(b_test[lo:hi], layers_pinn, xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn, tol)
    local = torch.zeros((hi - lo, xt_test.shape[0]), dtype=torch.float32, device=xt_test.device)
    with torch.no_grad():
        for k, rec in enumerate(recs):
            local[k] = rec["net"](xt_test)[:, 0]
    pinn_soln = offline.gather_rows(local, len(b_test)).t().cpu().numpy().astype(np.float64)
    return offline.spread_hours(time.time() - t0, len(b_test)), pinn_soln


def pinn_test_loss(b_test, layers_pinn, xt_resid, IC_xt, IC_u, BC_xt, BC_u,
                   f_hat, epochs_pinn, lr_pinn, tol):
    """B_test.py:154-181: losses[0] = initial loss; losses[j/1000] = the loss evaluated at the top of
    iteration j (after j-1 steps) for j % 1000 == 0 or j == epochs; on the tolerance break the stopping
    loss goes into the first still-zero slot."""
    b_test = np.asarray(b_test, dtype=np.float64).reshape(-1)
    lo, hi = offline.shard_bounds(len(b_test))
    recs = _trained(b_test[lo:hi], layers_pinn, xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn, tol)
    local = np.zeros((hi - lo, 61))
    for k, rec in enumerate(recs):
        done, tr = rec["done"], rec["trace"]                              # tr[j-1] = loss at the top of iteration j
        col = local[k]
        if done >= 1:
            col[0] = tr[0]
        stopped = tol is not None and done >= 1 and tr[done - 1] < tol
        for j in range(1, done + 1):
            if (j % 1000 == 0) or (j == epochs_pinn):
                col[int(j / 1000)] = tr[j - 1]
        if stopped:
            col[np.where(col == 0)[0][0]] = tr[done - 1]
    out = offline.gather_rows(torch.from_numpy(local).to(xt_resid.device), len(b_test))
    return out.t().cpu().numpy().astype(np.float64)
