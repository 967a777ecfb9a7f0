"""Drop-in for Burgers/B_optimizer.py: grad_descent(...).update(c, c_reshaped) -> c
(reference B_optimizer.py:4-58).  One update = one single-parameter, single-epoch launch of the
sweep kernel; T = P_t - nu P_xx is passed as the "P_t" block with nu = 0."""
import numpy as np

from _common import sweep


class grad_descent():
    def __init__(self, nu, xt_size, IC_size, BC_size, IC_u, out, out_IC,
                 out_BC, out_x, Pt_nu_P_xx_term, lr_gpt):
        self.nu = nu
        self.fac1, self.fac2, self.fac3 = 2 / xt_size, 2 / BC_size, 2 / IC_size
        self.out, self.out_BC, self.out_IC, self.out_x = out, out_BC, out_IC, out_x
        self.Pt_nu_P_xx_term = Pt_nu_P_xx_term
        self.IC_u = IC_u
        self.lr_gpt = lr_gpt
        self._grid = np.array([0.0])

    def update(self, c, c_reshaped):
        T = self.Pt_nu_P_xx_term
        r = sweep.b_sweep(self._grid, c.reshape(-1), self.out, self.out_x, T, T, self.out_IC, self.out_BC,
                          None, self.IC_u, None, 1, self.lr_gpt, want_c=True)
        return r["c"][0]
