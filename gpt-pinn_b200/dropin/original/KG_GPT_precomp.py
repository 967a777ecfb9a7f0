"""Legacy adapter for Klein-Gordon/original/KG_GPT_precomp.py (reference lines 7-39).

autograd_calculations / Pi_t return the reference's columns -- the second-order autograd output of the legacy code is the
same mixed-derivative channel as the current code's (u_xx + u_xt, u_tt + u_xt) -- from the closed-form CUDA
precompute kernel; the three elementwise helpers are the reference's one-line tensor expressions.
"""
import torch
from torch import cos

from _legacy_common import precomp

device = torch.device("cuda")


def _layers(P):
    return [(lin.weight.data, lin.bias.data) for lin in P.linears]


def autograd_calculations(xt_resid, P):
    xt = xt_resid.detach().to(device)
    q = torch.empty(xt.shape[0], device=device)
    r = torch.empty(xt.shape[0], device=device)
    precomp.mlp_channels(_layers(P), "cos", xt, q=q, r=r)
    return q[:, None], r[:, None]


def Pi_t(IC_xt, P):
    xt = IC_xt.detach().to(device)
    t = torch.empty(xt.shape[0], device=device)
    precomp.mlp_channels(_layers(P), "cos", xt, t=t)
    return t[:, None]


def Ptt_aPxx_bP(alpha, beta, P_tt, P_xx, P_resid_values):
    return torch.add(torch.add(P_tt, torch.mul(alpha, P_xx)), torch.mul(beta, P_resid_values))


def gamma2_P(gamma, P_resid_values):
    return torch.mul(2 * gamma, P_resid_values)


def xcos_x2cos2(x_resid, t_resid):
    return x_resid * cos(t_resid) - (x_resid ** 2) * (cos(t_resid) ** 2)
