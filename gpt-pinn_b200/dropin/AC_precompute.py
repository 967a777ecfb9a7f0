"""Drop-in for Allen-Cahn/AC_precompute.py.

inputs(): reference AC_precompute.py:26-58 -- column i of out, out_t, out_xx ("P_xx" = u_xx+u_xt),
out_IC, the periodic-boundary blocks (value and x-derivative at x = +1 / -1 and their differences)
and out_test, via the closed-form CUDA kernel; returns the reference's 12-tuple.
"""
import torch
import torch.autograd as autograd

from AC_models import GPT_residual  # noqa: F401  (the reference module imports it too)
from _common import precomp

device = torch.device("cuda")


def derivatives(xt, out):
    """Reference AC_precompute.py:7-14."""
    out_xt = autograd.grad(out, xt, torch.ones_like(out), create_graph=True)[0]
    out_xx_tt = autograd.grad(out_xt, xt, torch.ones_like(out_xt))[0]
    return out_xt[:, 1].detach().unsqueeze(1), out_xx_tt[:, 0].unsqueeze(1)


def boundary_derivative(BC_xt, out_BC):
    """Reference AC_precompute.py:16-18."""
    out_x = autograd.grad(out_BC, BC_xt, torch.ones_like(out_BC))[0]
    return out_x[:, 0].unsqueeze(1)


def Pt_lPxx_eP(out_t, out_xx, out, lmbda, eps):
    """Reference AC_precompute.py:20-24: (P_t + (-lambda)*P_xx) + (-eps)*P."""
    return torch.add(torch.add(out_t, torch.mul(-lmbda, out_xx)), torch.mul(-eps, out))


def inputs(PINN, xt, out, out_xx, out_t, out_IC, out_BC_ub, out_BC_lb, IC_xt,
           BC_xt_ub, BC_xt_lb, i, out_test, xt_test, f_hat, xt_size, out_BC_diff,
           out_BC_ub_x, out_BC_lb_x, out_BC_diff_x):
    end = i + 1
    act = getattr(PINN, "activation_name", "tanh")
    precomp.mlp_channels(PINN, act, xt, v=out[:, i], t=out_t[:, i], q=out_xx[:, i])
    precomp.mlp_channels(PINN, act, IC_xt, v=out_IC[:, i])
    precomp.mlp_channels(PINN, act, BC_xt_ub, v=out_BC_ub[:, i], x=out_BC_ub_x[:, i])
    precomp.mlp_channels(PINN, act, BC_xt_lb, v=out_BC_lb[:, i], x=out_BC_lb_x[:, i])
    out_BC_diff[:, i] = torch.sub(out_BC_ub[:, i], out_BC_lb[:, i])
    out_BC_diff_x[:, i] = torch.sub(out_BC_ub_x[:, i], out_BC_lb_x[:, i])
    precomp.mlp_channels(PINN, act, xt_test, v=out_test[:, i])
    return out[:, 0:end], out_xx[:, 0:end], out_t[:, 0:end], out_IC[:, 0:end], \
        out_BC_ub[:, 0:end], out_BC_lb[:, 0:end], out_BC_diff[:, 0:end], \
        f_hat, xt_size, out_BC_ub_x[:, 0:end], out_BC_lb_x[:, 0:end], \
        out_BC_diff_x[:, 0:end]
