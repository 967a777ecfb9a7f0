"""Drop-in for Klein-Gordon/KG_data.py (host-side point sets and IC/BC targets; reference
KG_data.py:15-69).  x in [Xi,Xf], t in [Ti,Tf]; IC u(x,0)=x, u_t(x,0)=0; BC u(+-1,t)=+-cos t."""
import torch


def initial_u(x, t=0):
    return x


def initial_dudt(x, t=0):
    return 0 * x


def boundary_bottom(t, x=-1):
    return -torch.cos(t)


def boundary_top(t, x=1):
    return torch.cos(t)


def _grid_points(Xi, Xf, Ti, Tf, n):
    """n*n tensor-product points, t-major (all x for t0, then t1, ...) like the reference's transpose+flatten."""
    x = torch.linspace(Xi, Xf, n)
    t = torch.linspace(Ti, Tf, n)
    X = x.repeat(n).unsqueeze(1)
    T = t.repeat_interleave(n).unsqueeze(1)
    return torch.hstack((X, T))


def ICBC_data(Xi, Xf, Ti, Tf, BC_pts, IC_pts):
    x_ic = torch.linspace(Xi, Xf, IC_pts).unsqueeze(1)
    IC = torch.hstack((x_ic, torch.zeros(IC_pts, 1)))
    IC_u1 = initial_u(x_ic)
    IC_u2 = initial_dudt(x_ic)
    t_bc = torch.linspace(Ti, Tf, BC_pts).unsqueeze(1)
    top = torch.hstack((torch.full((BC_pts, 1), float(torch.linspace(Xi, Xf, BC_pts)[-1])), t_bc))
    bottom = torch.hstack((torch.full((BC_pts, 1), float(torch.linspace(Xi, Xf, BC_pts)[0])), t_bc))
    xt_train_BC = torch.vstack((top, bottom))
    u_train_BC = torch.vstack((boundary_top(t_bc), boundary_bottom(t_bc)))
    return IC, IC_u1, IC_u2, xt_train_BC, u_train_BC


def residual_data(Xi, Xf, Ti, Tf, Nc, N_test):
    xt_resid = _grid_points(Xi, Xf, Ti, Tf, Nc)
    f_hat_train = torch.zeros((xt_resid.shape[0], 1))
    xt_test = _grid_points(Xi, Xf, Ti, Tf, N_test)
    return xt_resid, f_hat_train, xt_test
