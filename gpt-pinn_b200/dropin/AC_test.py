"""Drop-in for Allen-Cahn/AC_test.py: gpt_test (reference AC_test.py:12-46) and gpt_test_loss
(AC_test.py:50-76, which records the loss after EVERY epoch -> [epochs+1, n_cases])."""
import time

import numpy as np
import torch

from _common import offline, sweep

device = torch.device("cuda")


def gpt_test(ac_test, out, out_t, out_xx, out_IC, fhat, xt_len, IC_size, BC_size,
             IC_u, c_initial, epochs_gpt, lr_gpt, U_test, out_BC_ub, out_BC_lb,
             out_BC_diff, out_BC_ub_x, out_BC_lb_x, out_BC_diff_x):
    t0 = time.time()
    grid = np.asarray(ac_test, dtype=np.float64).reshape(-1, 2)

    def run(rows, _c0):
        return sweep.ac_sweep(rows, c_initial, out, out_t, out_xx, out_IC, out_BC_ub, out_BC_lb, out_BC_diff,
                              out_BC_ub_x, out_BC_lb_x, out_BC_diff_x, fhat, IC_u, epochs_gpt, lr_gpt, want_c=True)

    res = offline.sharded_sweep(run, grid, want_c=True, gather_c=True)
    gpt_soln = torch.matmul(U_test, res["c"].t()).detach().cpu().numpy().astype(np.float64)
    return offline.spread_hours(time.time() - t0, len(grid)), gpt_soln


def gpt_test_loss(ac_test, out, out_t, out_xx, out_IC, fhat, xt_len, IC_size, BC_size,
                  IC_u, c_initial, epochs_gpt, lr_gpt, U_test, out_BC_ub, out_BC_lb,
                  out_BC_diff, out_BC_ub_x, out_BC_lb_x, out_BC_diff_x):
    grid = np.asarray(ac_test, dtype=np.float64).reshape(-1, 2)
    lo, hi = offline.shard_bounds(len(grid))
    r = sweep.ac_sweep(grid[lo:hi], c_initial, out, out_t, out_xx, out_IC, out_BC_ub, out_BC_lb, out_BC_diff,
                       out_BC_ub_x, out_BC_lb_x, out_BC_diff_x, fhat, IC_u, epochs_gpt, lr_gpt,
                       rec_every=1, want_c=False)
    tr = offline.gather_rows(r["trace"], len(grid))           # [n_cases, epochs+1]
    return tr.t().cpu().numpy().astype(np.float64)
