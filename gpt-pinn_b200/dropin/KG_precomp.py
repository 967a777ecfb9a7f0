"""Drop-in for Klein-Gordon/KG_precomp.py.

inputs(): reference KG_precomp.py:34-52 -- fills column i of the caller's seven [N, neurons]
buffers.  Here the values and derivative channels come from one closed-form CUDA kernel per
point set (gptpinn_mlp_channels) instead of four forwards and three autograd calls; the "P_xx" /
"P_tt" columns carry the reference's mixed-derivative semantics (u_xx+u_xt, u_tt+u_xt).
"""
import torch
import torch.autograd as autograd
from torch import cos

from _common import precomp

device = torch.device("cuda")


def second_derivatives(xt, out):
    """Reference KG_precomp.py:7-16 (autograd form, kept for callers that hold a graph)."""
    out_xt = autograd.grad(out, xt, torch.ones_like(out), create_graph=True)[0]
    out_xx_tt = autograd.grad(out_xt, xt, torch.ones_like(out_xt))[0]
    return out_xx_tt[:, 0].unsqueeze(1), out_xx_tt[:, 1].unsqueeze(1)


def initial_derivative(IC_xt, out_IC):
    """Reference KG_precomp.py:18-21."""
    out_t = autograd.grad(out_IC, IC_xt, torch.ones_like(out_IC))[0]
    return out_t[:, 1].unsqueeze(1)


def Ptt_aPxx_bP(alpha, beta, out_tt, out_xx, out):
    """Reference KG_precomp.py:23-26: (P_tt + alpha*P_xx) + beta*P."""
    return torch.add(torch.add(out_tt, torch.mul(alpha, out_xx)), torch.mul(beta, out))


def gamma2_P(gamma, out):
    """Reference KG_precomp.py:28-29."""
    return torch.mul(2 * gamma, out)


def xcos_term(x, t):
    """Reference KG_precomp.py:31-32."""
    return x * cos(t) - (x ** 2) * (cos(t) ** 2)


def inputs(PINN, xt, out, out_xx, out_tt, out_IC_t,
           out_IC, out_BC, IC_xt, BC_xt, i, out_test, xt_test):
    end = i + 1
    act = getattr(PINN, "activation_name", "cos")
    precomp.mlp_channels(PINN, act, xt, v=out[:, i], q=out_xx[:, i], r=out_tt[:, i])
    precomp.mlp_channels(PINN, act, IC_xt, v=out_IC[:, i], t=out_IC_t[:, i])
    precomp.mlp_channels(PINN, act, BC_xt, v=out_BC[:, i])
    precomp.mlp_channels(PINN, act, xt_test, v=out_test[:, i])
    return out[:, 0:end], out_xx[:, 0:end], out_tt[:, 0:end], out_IC_t[:, 0:end], \
        out_IC[:, 0:end], out_BC[:, 0:end]
