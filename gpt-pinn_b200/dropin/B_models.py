"""Drop-in for Burgers/B_models.py: GPT (reference B_models.py:11-46) and NN (B_models.py:52-105,
[2,20,20,20,20,1] tanh PINN, Xavier-normal weights, zero biases, seed 1234 in the constructor)."""
import numpy as np
import torch
import torch.autograd as autograd
import torch.nn as nn

from _common import precomp, sweep

device = torch.device("cuda")


class GPT(nn.Module):
    def __init__(self, nu, out, out_IC, out_BC, out_x, Pt_nu_P_xx_term, IC_u,
                 BC_u, fhat):
        super().__init__()
        self.nu = nu
        self.loss_function = nn.MSELoss(reduction='mean')
        self.IC_u, self.BC_u, self.f_hat = IC_u, BC_u, fhat
        self.Pt_nu_P_xx_term = Pt_nu_P_xx_term
        self.out_IC, self.out_BC, self.out, self.out_x = out_IC, out_BC, out, out_x
        self._grid = np.array([0.0])

    def lossR(self, c):
        f = self.Pt_nu_P_xx_term @ c + (self.out @ c) * (self.out_x @ c)
        return self.loss_function(f, self.f_hat)

    def lossBC(self, c):
        return self.loss_function(self.out_BC @ c, self.BC_u)

    def lossIC(self, c):
        return self.loss_function(self.out_IC @ c, self.IC_u)

    def loss(self, c):
        T = self.Pt_nu_P_xx_term
        r = sweep.b_sweep(self._grid, c.reshape(-1), self.out, self.out_x, T, T, self.out_IC, self.out_BC,
                          self.f_hat, self.IC_u, self.BC_u, 0, 0.0, want_c=False)
        return r["f"][0]


class NN(nn.Module):
    def __init__(self, layers, nu):
        super().__init__()
        torch.manual_seed(1234)
        self.layers = layers
        self.nu = nu
        self.loss_function = nn.MSELoss(reduction="mean")
        self.linears = nn.ModuleList([nn.Linear(int(layers[i]), int(layers[i + 1])) for i in range(len(layers) - 1)])
        for lin in self.linears:
            nn.init.xavier_normal_(lin.weight.data)
            nn.init.zeros_(lin.bias.data)
        self.activation = nn.Tanh()
        self.activation_name = "tanh"

    def forward(self, x):
        if x.is_cuda and not (torch.is_grad_enabled() and (x.requires_grad or self.linears[0].weight.requires_grad)):
            u = torch.empty((x.shape[0], 1), dtype=torch.float32, device=x.device)
            precomp.mlp_channels(self, "tanh", x, v=u)
            return u
        a = x
        for lin in self.linears[:-1]:
            a = torch.tanh(lin(a))
        return self.linears[-1](a)

    def lossR(self, xt, f_hat):
        u = self.forward(xt)
        u_xt = autograd.grad(u, xt, torch.ones_like(u), create_graph=True)[0]
        u_xx_tt = autograd.grad(u_xt, xt, torch.ones_like(u_xt), create_graph=True)[0]
        u_xx = u_xx_tt[:, 0].unsqueeze(1)
        u_x = u_xt[:, 0].unsqueeze(1)
        u_t = u_xt[:, 1].unsqueeze(1)
        f = u_t + (u * u_x + (-self.nu) * u_xx)
        return self.loss_function(f, f_hat)

    def lossICBC(self, ICBC_xt, ICBC_u):
        return self.loss_function(self.forward(ICBC_xt), ICBC_u)

    def loss(self, xt_resid, IC_xt, IC_u, BC_xt, BC_u, f_hat):
        return self.lossR(xt_resid, f_hat) + self.lossICBC(IC_xt, IC_u) + self.lossICBC(BC_xt, BC_u)
