"""Drop-in for Allen-Cahn/AC_optimizer.py: grad_descent(...).update(c, c_reshaped) -> c
(reference AC_optimizer.py:4-66).  One update = one single-parameter, single-epoch sweep launch;
T = P_t - lambda P_xx - eps P is passed as the "P_t" block with lambda = 0 and the eps P term
removed by handing a zero block for P in the T build (eps still drives the cubic term)."""
import numpy as np
import torch

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
        # T_given + eps*P is the "P_t" the kernel needs so that (P_t + 0*P_xx) + (-eps)*P == T_given
        self._Pt = torch.add(Pt_lPxx_eP_term, torch.mul(eps, out))
        self._grid = np.array([[0.0, float(eps)]])

    def update(self, c, c_reshaped):
        r = sweep.ac_sweep(self._grid, c.reshape(-1), self.out, self._Pt, self._Pt, self.out_IC, self.out_BC_ub,
                           self.out_BC_lb, self.out_BC_diff, self.out_BC_ub_x, self.out_BC_lb_x,
                           self.out_BC_diff_x, None, self.IC_u, 1, self.lr_gpt, want_c=True)
        return r["c"][0]
