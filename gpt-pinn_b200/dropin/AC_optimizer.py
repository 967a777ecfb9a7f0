"""Drop-in for Allen-Cahn/AC_optimizer.py: grad_descent(...).update(c, c_reshaped) -> c
(reference AC_optimizer.py:4-66).  One update = one single-parameter, single-epoch sweep launch;
the materialised T = P_t - lambda P_xx - eps P the constructor receives is handed over as is
(gptpinn_ac_problem.t_given: used bit for bit; eps still drives the cubic term)."""
import numpy as np

from _common import sweep


class grad_descent():
    def __init__(self, lmbda, eps, xt_size, IC_size, BC_size, IC_u,
                 Pt_lPxx_eP_term, out, out_IC, out_BC_ub, out_BC_lb,
                 out_BC_diff, out_BC_ub_x, out_BC_lb_x, out_BC_diff_x,
                 epochs_gpt, lr_gpt):
        self.lmbda, self.eps = lmbda, eps
        self.fac1, self.fac2, self.fac3 = 2 / xt_size, 2 / BC_size, 2 / IC_size
        self.out = out
        self.out_BC_ub, self.out_BC_lb, self.out_BC_diff = out_BC_ub, out_BC_lb, out_BC_diff
        self.out_BC_ub_x, self.out_BC_lb_x, self.out_BC_diff_x = out_BC_ub_x, out_BC_lb_x, out_BC_diff_x
        self.out_IC = out_IC
        self.Pt_lPxx_eP_term = Pt_lPxx_eP_term
        self.IC_u = IC_u
        self.lr_gpt = lr_gpt
        self._grid = np.array([[0.0, float(eps)]])

    def update(self, c, c_reshaped):
        T = self.Pt_lPxx_eP_term
        r = sweep.ac_sweep(self._grid, c.reshape(-1), self.out, T, T, self.out_IC, self.out_BC_ub,
                           self.out_BC_lb, self.out_BC_diff, self.out_BC_ub_x, self.out_BC_lb_x,
                           self.out_BC_diff_x, None, self.IC_u, 1, self.lr_gpt, want_c=True, t_given=True)
        return r["c"][0]
