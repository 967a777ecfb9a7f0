"""Legacy adapter for Klein-Gordon/original/KG_GPT_PINN.py: class GPT, the reduced model that owns its coefficients as
`linears[1].weight` (reference :6-84).  loss() is one zero-epoch launch of the sweep kernel (T passed as the "P_tt" block);
forward / the partial losses are the reference's tensor expressions, kept for API compatibility."""
import numpy as np
import torch
import torch.nn as nn

from _legacy_common import sweep

device = torch.device("cuda")


class GPT(nn.Module):
    def __init__(self, layers, alpha, beta, gamma, P, initial_c, IC_u1, IC_u2, BC_u, f_hat, xcos_x2cos2_term,
                 activation_resid, activation_IC, activation_BC, Pi_t_term, Ptt_aPxx_bP_term):
        super().__init__()
        self.layers, self.alpha, self.beta, self.gamma, self.activation = layers, alpha, beta, gamma, P
        self.loss_function = nn.MSELoss(reduction='mean')
        self.linears = nn.ModuleList([nn.Linear(int(layers[i]), int(layers[i + 1]), bias=False) for i in range(len(layers) - 1)])
        self.IC_u1, self.IC_u2, self.BC_u, self.f_hat = IC_u1, IC_u2, BC_u, f_hat
        self.Pi_t_term, self.xcos_x2cos2_term, self.Ptt_aPxx_bP_term = Pi_t_term, xcos_x2cos2_term, Ptt_aPxx_bP_term
        self.activation_IC, self.activation_BC, self.activation_resid = activation_IC, activation_BC, activation_resid
        self.linears[0].weight.data = torch.ones(int(layers[1]), int(layers[0]))
        self.linears[1].weight.data = initial_c
        self._grid = np.array([[0.0, 0.0, float(gamma)]])

    def _c(self):
        return self.linears[1].weight.data.to(device).view(-1)

    def forward(self, datatype=None, test_data=None):
        if test_data is not None:
            a = torch.cat([p(test_data) for p in self.activation[:int(self.layers[1])]], 1)
            return a @ self._c()[:, None]
        a = {"residual": self.activation_resid, "initial": self.activation_IC, "boundary": self.activation_BC}[datatype]
        return a @ self._c()[:, None]

    def lossR(self):
        u = self.forward(datatype='residual')
        f = self.Ptt_aPxx_bP_term @ self._c()[:, None] + self.gamma * torch.square(u) + self.xcos_x2cos2_term
        return self.loss_function(f, self.f_hat)

    def lossIC1BC(self, datatype):
        return self.loss_function(self.forward(datatype=datatype), self.IC_u1 if datatype == 'initial' else self.BC_u)

    def lossIC2(self):
        return self.loss_function(self.Pi_t_term @ self._c()[:, None], self.IC_u2)

    def loss(self):
        T = self.Ptt_aPxx_bP_term
        r = sweep.kg_sweep(self._grid, self._c(), self.activation_resid, T, T, self.activation_IC, self.Pi_t_term, self.activation_BC,
                           self.xcos_x2cos2_term, self.f_hat, self.IC_u1, self.IC_u2, self.BC_u, 0, 0.0, want_c=False)
        return r["f"][0]
