"""Drop-in for Klein-Gordon/KG_models.py: GPT (reduced-model loss) and NN (the full PINN).

GPT: reference KG_models.py:7-50.  `loss(c)` is evaluated by the sweep kernel (zero epochs);
the four partial losses are three-line tensor expressions kept for API compatibility.
NN : reference KG_models.py:62-129 -- [2,40,40,1] cos network, Xavier-normal weights, zero
biases, torch.manual_seed(1234) in the constructor, residual built from the reference's mixed
second derivatives.
"""
import numpy as np
import torch
import torch.autograd as autograd
import torch.nn as nn

from _common import precomp, sweep

device = torch.device("cuda")


class GPT(nn.Module):
    def __init__(self, alpha, beta, gamma, out, out_IC, out_IC_t,
                 out_BC, xcos, fhat, Ptt_aPxx_bP_term, IC_u1, IC_u2, BC_u):
        super().__init__()
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.loss_function = nn.MSELoss(reduction='mean')
        self.IC_u1, self.IC_u2, self.BC_u, self.fhat = IC_u1, IC_u2, BC_u, fhat
        self.out_IC_t, self.xcos, self.Ptt_aPxx_bP_term = out_IC_t, xcos, Ptt_aPxx_bP_term
        self.out_IC, self.out_BC, self.out = out_IC, out_BC, out
        self._grid = np.array([[0.0, 0.0, float(gamma)]])

    def lossR(self, c):
        u = self.out @ c
        f = self.Ptt_aPxx_bP_term @ c + self.gamma * u * u + self.xcos
        return self.loss_function(f, self.fhat)

    def lossBC(self, c):
        return self.loss_function(self.out_BC @ c, self.BC_u)

    def lossIC1(self, c):
        return self.loss_function(self.out_IC @ c, self.IC_u1)

    def lossIC2(self, c):
        return self.loss_function(self.out_IC_t @ c, self.IC_u2)

    def loss(self, c):
        T = self.Ptt_aPxx_bP_term
        r = sweep.kg_sweep(self._grid, c.reshape(-1), self.out, T, T, self.out_IC, self.out_IC_t, self.out_BC,
                           self.xcos, self.fhat, self.IC_u1, self.IC_u2, self.BC_u, 0, 0.0, want_c=False)
        return r["f"][0]


class cos(nn.Module):
    def forward(self, x):
        return torch.cos(x)


class NN(nn.Module):
    def __init__(self, layers, alpha, beta, gamma, xcos_x2cos2):
        super().__init__()
        torch.manual_seed(1234)
        self.layers = layers
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.loss_function = nn.MSELoss(reduction="mean")
        self.linears = nn.ModuleList([nn.Linear(int(layers[i]), int(layers[i + 1])) for i in range(len(layers) - 1)])
        self.xcos_x2cos2 = xcos_x2cos2
        for lin in self.linears:
            nn.init.xavier_normal_(lin.weight.data)
            nn.init.zeros_(lin.bias.data)
        self.activation = cos()
        self.activation_name = "cos"

    def forward(self, x):
        if x.is_cuda and not (torch.is_grad_enabled() and (x.requires_grad or self.linears[0].weight.requires_grad)):
            u = torch.empty((x.shape[0], 1), dtype=torch.float32, device=x.device)
            precomp.mlp_channels(self, "cos", x, v=u)
            return u
        a = x
        for lin in self.linears[:-1]:
            a = torch.cos(lin(a))
        return self.linears[-1](a)

    def lossR(self, xt, f_hat):
        u = self.forward(xt)
        u_xt = autograd.grad(u, xt, torch.ones_like(u), create_graph=True)[0]
        u_xx_tt = autograd.grad(u_xt, xt, torch.ones_like(u_xt), create_graph=True)[0]
        u_xx = u_xx_tt[:, 0].unsqueeze(1)
        u_tt = u_xx_tt[:, 1].unsqueeze(1)
        f = u_tt + self.alpha * u_xx + self.beta * u + self.gamma * u * u + self.xcos_x2cos2
        return self.loss_function(f, f_hat)

    def lossIC1BC(self, ICBC_xt, ICBC_u):
        return self.loss_function(self.forward(ICBC_xt), ICBC_u)

    def lossIC2(self, IC_xt, IC_u2):
        u = self.forward(IC_xt)
        u_xt = autograd.grad(u, IC_xt, torch.ones_like(u), create_graph=True)[0]
        return self.loss_function(u_xt[:, 1].unsqueeze(1), IC_u2)

    def loss(self, xt_residual, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u):
        return (self.lossR(xt_residual, f_hat) + self.lossIC1BC(IC_xt, IC_u1)
                + self.lossIC2(IC_xt, IC_u2) + self.lossIC1BC(BC_xt, BC_u))
