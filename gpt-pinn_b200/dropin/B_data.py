"""Drop-in for Burgers/B_data.py (reference B_data.py:12-64): IC u(x,0) = -sin(pi x), BC u(+-1,t) = 0."""
import torch


def initial_u(x, t=0):
    return -torch.sin(torch.pi * x)


def boundary_top(t, x=-1):
    return 0 * t


def boundary_bottom(t, x=1):
    return 0 * t


def _grid_points(Xi, Xf, Ti, Tf, n):
    x = torch.linspace(Xi, Xf, n)
    t = torch.linspace(Ti, Tf, n)
    return torch.hstack((x.repeat(n).unsqueeze(1), t.repeat_interleave(n).unsqueeze(1)))


def ICBC_data(Xi, Xf, Ti, Tf, BC_pts, IC_pts):
    x_ic = torch.linspace(Xi, Xf, IC_pts).unsqueeze(1)
    IC = torch.hstack((x_ic, torch.zeros(IC_pts, 1)))
    IC_u = initial_u(x_ic)
    xs = torch.linspace(Xi, Xf, BC_pts)
    t_bc = torch.linspace(Ti, Tf, BC_pts).unsqueeze(1)
    top = torch.hstack((torch.full((BC_pts, 1), float(xs[-1])), t_bc))
    bottom = torch.hstack((torch.full((BC_pts, 1), float(xs[0])), t_bc))
    xt_train_BC = torch.vstack((top, bottom))
    u_train_BC = torch.vstack((boundary_top(t_bc), boundary_bottom(t_bc)))
    return IC, IC_u, xt_train_BC, u_train_BC


def residual_data(Xi, Xf, Ti, Tf, Nc, N_test):
    xt_resid = _grid_points(Xi, Xf, Ti, Tf, Nc)
    f_hat_train = torch.zeros((xt_resid.shape[0], 1))
    xt_test = _grid_points(Xi, Xf, Ti, Tf, N_test)
    return xt_resid, f_hat_train, xt_test
