"""Drop-in for Allen-Cahn/AC_train.py: offline_generation (reference AC_train.py:5-40) returns
(largest_loss, largest_case, trained_c) -- trained_c is the winner's coefficients after all epochs."""
import numpy as np

from _common import offline, sweep


def offline_generation(ac_train, c_initial, xt_size, IC_size, BC_size, IC_u,
                       out, out_t, out_xx, out_IC, out_BC_ub, out_BC_lb,
                       out_BC_diff, out_BC_ub_x, out_BC_lb_x, out_BC_diff_x,
                       f_hat, epochs_gpt, lr_gpt):
    grid = np.asarray(ac_train, dtype=np.float64).reshape(-1, 2)

    def run(rows, _c0, epochs=epochs_gpt, **kw):
        return sweep.ac_sweep(rows, c_initial, out, out_t, out_xx, out_IC, out_BC_ub, out_BC_lb, out_BC_diff,
                              out_BC_ub_x, out_BC_lb_x, out_BC_diff_x, f_hat, IC_u, epochs, lr_gpt,
                              want_c=True, **kw)

    largest_loss, idx, res = offline.greedy_select(run, grid, epochs_gpt, want_c=True)
    if idx < 0:
        # the reference leaves trained_c unbound here (UnboundLocalError); keep that failure loud
        raise UnboundLocalError("offline_generation: no parameter produced a loss above 0 (trained_c is unbound)")
    lmbda, eps = ac_train[idx]
    trained_c = offline.winner_row(res, idx, len(grid))
    return largest_loss, (lmbda, eps), trained_c
