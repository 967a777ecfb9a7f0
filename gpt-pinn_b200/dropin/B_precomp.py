"""Drop-in for Burgers/B_precomp.py.

inputs(): reference B_precomp.py:22-55 -- fills column i of out, out_t, out_x, out_xx (with
"P_xx" = u_xx + u_xt), then zeroes, in that column only, the `num_mag` rows with the largest
|P_xx| (ascending sort, last `num_mag` indices, recorded in idx_list[i]); then the IC / BC / test
columns.  Values and derivatives come from the closed-form CUDA kernel instead of autograd.
"""
import torch
import torch.autograd as autograd

from _common import precomp

device = torch.device("cuda")


def derivatives(xt, out):
    """Reference B_precomp.py:6-16 (autograd form, kept for API compatibility)."""
    out_xt = autograd.grad(out, xt, torch.ones_like(out), create_graph=True)[0]
    out_xx_tt = autograd.grad(out_xt, xt, torch.ones_like(out_xt))[0]
    out_x = out_xt[:, 0].detach().unsqueeze(1)
    out_t = out_xt[:, 1].detach().unsqueeze(1)
    out_xx = out_xx_tt[:, 0].unsqueeze(1)
    return out_t, out_x, out_xx


def Pt_nu_P_xx(nu, out_t, out_xx):
    """Reference B_precomp.py:18-20: (-nu)*P_xx + P_t."""
    return torch.add(torch.mul(-nu, out_xx), out_t)


def inputs(PINN, xt, out, out_t, out_x, out_xx, out_IC, out_BC, IC_xt, BC_xt,
           i, out_test, xt_test, xt_size, num_mag, idx_list):
    end = i + 1
    act = getattr(PINN, "activation_name", "tanh")
    precomp.mlp_channels(PINN, act, xt, v=out[:, i], x=out_x[:, i], t=out_t[:, i], q=out_xx[:, i])
    val, index = torch.sort(torch.abs(out_xx[:, i]))
    largest = index[xt_size - num_mag:xt_size]
    idx_list[i] = largest.cpu()
    for buf in (out_t, out_x, out_xx, out):
        buf[largest, i] = 0.0
    precomp.mlp_channels(PINN, act, IC_xt, v=out_IC[:, i])
    precomp.mlp_channels(PINN, act, BC_xt, v=out_BC[:, i])
    precomp.mlp_channels(PINN, act, xt_test, v=out_test[:, i])
    return out[:, 0:end], out_x[:, 0:end], out_t[:, 0:end], out_xx[:, 0:end], \
        out_IC[:, 0:end], out_BC[:, 0:end]
