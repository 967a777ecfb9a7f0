"""Drop-in for Klein-Gordon/KG_test.py.

gpt_test / gpt_test_loss (reference KG_test.py:10-44, 49-76): the per-test-parameter loop of 5000
updates is ONE fused launch over all test parameters; gpt_test_loss reads the loss trace the
kernel records every 500 epochs.  pinn_test / pinn_test_loss (KG_test.py:81-140) train one full
PINN per test parameter.
"""
import time

import numpy as np
import torch

from KG_models import NN
from _common import offline, pinn, sweep

device = torch.device("cuda")
PINN_BATCH = 8          # test parameters per training launch (1..8): 31 us per Adam epoch and training at the KG sizes, 66 us alone


def gpt_test(kg_test, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, fhat,
             xcos, xt_len, IC_size, BC_size, IC_u1, IC_u2, BC_u, c_initial,
             epochs_gpt, lr_gpt, U_test):
    t0 = time.time()
    grid = np.asarray(kg_test, dtype=np.float64).reshape(-1, 3)

    def run(rows, _c0):
        return sweep.kg_sweep(rows, c_initial, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, xcos, fhat,
                              IC_u1, IC_u2, BC_u, epochs_gpt, lr_gpt, want_c=True)

    res = offline.sharded_sweep(run, grid, want_c=True, gather_c=True, group_cols=(0, 1))
    soln = torch.matmul(U_test, res["c"].t())                 # [N_test, n_cases]
    gpt_soln = soln.detach().cpu().numpy().astype(np.float64)
    times = offline.spread_hours(time.time() - t0, len(grid))
    return times, gpt_soln


def gpt_test_loss(kg_test, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, fhat,
                  xcos, xt_len, IC_size, BC_size, IC_u1, IC_u2, BC_u,
                  c_initial, epochs_gpt, lr_gpt):
    grid = np.asarray(kg_test, dtype=np.float64).reshape(-1, 3)
    lo, hi = offline.shard_bounds(len(grid))
    r = sweep.kg_sweep(grid[lo:hi], c_initial, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, xcos, fhat,
                       IC_u1, IC_u2, BC_u, epochs_gpt, lr_gpt, rec_every=500, want_c=False)
    trace = offline.gather_rows(r["trace"], len(grid))        # [n_cases, epochs/500 + 1]
    losses = np.zeros((11, len(grid)))
    tr = trace.t().cpu().numpy().astype(np.float64)
    losses[:min(11, tr.shape[0])] = tr[:11]
    return losses


def _chunks(n):
    """Test parameters trained together in one launch (gptpinn_pinn_train_batch): as few launches as PINN_BATCH allows,
    of near-equal size (25 parameters -> 7, 6, 6, 6)."""
    b = max(1, min(pinn.MAX_BATCH, PINN_BATCH))
    k = max(1, -(-n // b))
    cuts = [round(i * n / k) for i in range(k + 1)]
    return [(cuts[i], cuts[i + 1]) for i in range(k) if cuts[i + 1] > cuts[i]]


def _trained(mine, layers_pinn, xcos_x2cos2, xt_resid, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn):
    """The full PINNs of this rank's test parameters, trained once per (parameters, data, epochs, lr): pinn_test and
    pinn_test_loss ask for the same trainings (see gptpinn.pinn: memo of finished test-set trainings).  Per parameter:
    the trained network, the loss before the step of epochs 1, 1001, 2001, ... and the loss after the last step."""
    chunks = _chunks(len(mine))
    tensors = (xcos_x2cos2, xt_resid, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, f_hat)
    key = pinn.train_memo_key("kg", layers_pinn, mine, epochs_pinn, lr_pinn, None, chunks, tensors)
    recs = pinn.train_memo_get(key)
    if recs is not None:
        return recs
    recs = []
    for a, b in chunks:
        nets = [NN(layers_pinn, *mine[k], xcos_x2cos2).to(device) for k in range(a, b)]
        mus = [tuple(mine[k]) for k in range(a, b)]
        # loss before the step of epochs 1, 1001, 2001, ... == loss after 0, 1000, 2000, ... steps
        res = pinn.train_fused_batch(nets, "kg", mus, xt_resid, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u,
                                     epochs_pinn, lr_pinn, xcos=xcos_x2cos2, trace_every=1000)
        fin = pinn.train_fused_batch(nets, "kg", mus, xt_resid, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u,
                                     1, 0.0, xcos=xcos_x2cos2)           # lr = 0: evaluates the loss, leaves the weights
        for net, r, f in zip(nets, res, fin):
            recs.append(dict(net=net, trace=r["trace"][:101].clone(), final=f["loss"].clone()))
    pinn.train_memo_put(key, tensors, recs)
    return recs


def pinn_test(kg_test, layers_pinn, xcos_x2cos2, xt_resid, IC_xt, IC_u1, IC_u2,
              BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn, xt_test):
    """One full PINN per test parameter (KG_test.py:81-111); the parameters are independent, so several are trained
    per launch, and under torch.distributed each rank trains a contiguous block of them (solutions all-gathered)."""
    kg_test = np.asarray(kg_test, dtype=np.float64).reshape(-1, 3)
    t0 = time.time()
    lo, hi = offline.shard_bounds(len(kg_test))
    recs = _trained(kg_test[lo:hi], layers_pinn, xcos_x2cos2, xt_resid, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn)
    local = torch.zeros((hi - lo, xt_test.shape[0]), dtype=torch.float32, device=xt_test.device)
    with torch.no_grad():
        for k, rec in enumerate(recs):
            local[k] = rec["net"](xt_test)[:, 0]
    soln = offline.gather_rows(local, len(kg_test))
    pinn_soln = soln.t().cpu().numpy().astype(np.float64)
    times = offline.spread_hours(time.time() - t0, len(kg_test))
    return times, pinn_soln


def pinn_test_loss(kg_test, layers_pinn, xcos_x2cos2, xt_resid, IC_xt, IC_u1,
                   IC_u2, BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn):
    """Loss after 0, 1000, 2000, ... steps and after the last step (KG_test.py:116-140), sharded like pinn_test."""
    kg_test = np.asarray(kg_test, dtype=np.float64).reshape(-1, 3)
    lo, hi = offline.shard_bounds(len(kg_test))
    recs = _trained(kg_test[lo:hi], layers_pinn, xcos_x2cos2, xt_resid, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, f_hat, epochs_pinn, lr_pinn)
    local = torch.zeros((hi - lo, 101), dtype=torch.float32, device=xt_resid.device)
    for k, rec in enumerate(recs):
        tr = rec["trace"]
        local[k, :min(101, tr.numel())] = tr[:101]
        local[k, min(100, int(epochs_pinn / 1000))] = rec["final"]
    return offline.gather_rows(local, len(kg_test)).t().cpu().numpy().astype(np.float64)
