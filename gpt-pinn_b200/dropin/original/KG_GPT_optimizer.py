"""Legacy adapter for Klein-Gordon/original/KG_GPT_optimizer.py: grad_descent(...).update(c) -> c[1, n]  (reference :5-77).

One update = one launch of the sweep kernel with one parameter and one epoch; the materialised T rides in as the "P_tt"
block with alpha = beta = 0 ((T + 0*T) + 0*P == T bit for bit).  The legacy update has no f_hat / IC_u2 term (N3).
"""
import numpy as np
import torch

from _legacy_common import sweep


class grad_descent(object):
    def __init__(self, alpha, beta, gamma, xt_resid, IC_xt, BC_xt, IC_u1, IC_u2, BC_u, xcos_x2cos2_term, Ptt_aPxx_bP_term,
                 gamm2_P_term, P_resid_values, P_IC_values, P_BC_values, Pi_t_term, epochs_gpt, lr_gpt):
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.N_R, self.N_IC, self.N_BC = xt_resid.shape[0], IC_xt.shape[0], BC_xt.shape[0]
        self.P_resid_values, self.P_BC_values, self.P_IC_values, self.Pi_t_term = P_resid_values, P_BC_values, P_IC_values, Pi_t_term
        self.xcos_x2cos2_term, self.Ptt_aPxx_bP_term, self.gamm2_P_term = xcos_x2cos2_term, Ptt_aPxx_bP_term, gamm2_P_term
        self.IC_u1, self.IC_u2, self.BC_u = IC_u1, IC_u2, BC_u
        self.lr_gpt = lr_gpt
        self._grid = np.array([[0.0, 0.0, float(gamma)]])

    def _sweep(self, c, epochs, **kw):
        T = self.Ptt_aPxx_bP_term
        return sweep.kg_sweep(self._grid, c.reshape(-1), self.P_resid_values, T, T, self.P_IC_values, self.Pi_t_term,
                              self.P_BC_values, self.xcos_x2cos2_term, None, self.IC_u1, None, self.BC_u, epochs, self.lr_gpt, **kw)

    def update(self, c):
        return self._sweep(c, 1, want_c=True)["c"]
