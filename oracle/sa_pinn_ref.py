"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the self-adaptive PINN the reference trains for Allen-Cahn.

The reference's trainer is TensorFlow code (Allen-Cahn/AC_main.py:204-339 = AC_SA_test.py:109-245, with
eager_lbfgs.py:14-200); TensorFlow and pyDOE are absent from this image, so this path cannot be pinned against a run of
the reference: **parity unpinned** (statistical only for whole trajectories).  What this file pins instead:

  * `sa_loss_autograd`  -- the loss exactly as AC_main.py:215-244 states it, differentiated by torch autograd (any dtype;
    the tests use float64).  The product's hand-derived forward/reverse pass is held to its gradients.
  * `TorchOps`          -- the channel recurrences (u, u_x, u_t, u_xx through tanh layers) and their adjoints written with
    torch tensor ops: the per-kernel reference for gpt-pinn_b200/csrc/sa_pinn.cu, and a CPU stand-in for host-logic tests.
  * `keras_adam_step`   -- tf.keras.optimizers.Adam (TF 2.10: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t*m/(sqrt(v)+eps), eps 1e-7).
  * `lbfgs_reference`   -- a literal port of eager_lbfgs.py:14-200 (fixed step, history 50, its break tests incl. the
    `tf.abs(f, f_old)` quirk which is |f| < tolX) over plain torch vectors.
"""
import math

import torch


def mlp_u(params, xt):
    a = xt
    for W, b in params[:-1]:
        a = torch.tanh(a @ W.t() + b)
    W, b = params[-1]
    return a @ W.t() + b


def sa_loss_autograd(params, xt_f, xt0, u0, xt_lb, xt_ub, wf, wu, lam, eps):
    """(total, mse_u, mse_b, mse_f) of AC_main.py:215-232; f = u_t - lam*u_xx + eps*u^3 - eps*u with the TRUE u_xx (:235-244)."""
    def with_dx(xt, second):
        x = xt[:, 0:1].clone().requires_grad_()
        t = xt[:, 1:2].clone().requires_grad_()
        u = mlp_u(params, torch.cat((x, t), 1))
        ux = torch.autograd.grad(u, x, torch.ones_like(u), create_graph=True)[0]
        if not second:
            return u, ux
        uxx = torch.autograd.grad(ux, x, torch.ones_like(ux), create_graph=True)[0]
        ut = torch.autograd.grad(u, t, torch.ones_like(u), create_graph=True)[0]
        return u, ux, ut, uxx
    u, ux, ut, uxx = with_dx(xt_f, True)
    f = ut - lam * uxx + eps * u * u * u - eps * u
    u0p = mlp_u(params, xt0)
    ulb, uxlb = with_dx(xt_lb, False)
    uub, uxub = with_dx(xt_ub, False)
    mse_0_u = torch.mean((wu * (u0 - u0p)) ** 2)
    mse_b = torch.mean((ulb - uub) ** 2) + torch.mean((uxlb - uxub) ** 2)
    mse_f_u = torch.mean((wf * f) ** 2)
    return mse_0_u + mse_b + mse_f_u, torch.mean((u0 - u0p) ** 2), mse_b, torch.mean(f ** 2)


class TorchOps:
    """Channel ops on [4, N, W] tensors (channels: 0 = u, 1 = d/dx, 2 = d/dt, 3 = d2/dx2), any device / dtype."""

    def layer0(self, xt, W0, b0, Z):
        Z[0] = xt @ W0.t() + b0
        Z[1] = W0[:, 0].expand_as(Z[1])
        Z[2] = W0[:, 1].expand_as(Z[2])
        Z[3] = 0

    def act_forward(self, Z, bias, A):
        if bias is not None:
            Z[0] += bias
        s0 = torch.tanh(Z[0])
        s1 = 1 - s0 * s0
        s2 = -2 * s0 * s1
        A[0] = s0
        A[1] = s1 * Z[1]
        A[2] = s1 * Z[2]
        A[3] = s2 * Z[1] * Z[1] + s1 * Z[3]

    def act_backward(self, Z, G):
        """G holds the adjoint of A on entry and the adjoint of Z on return."""
        s0 = torch.tanh(Z[0])
        s1 = 1 - s0 * s0
        s2 = -2 * s0 * s1
        s3 = s1 * (6 * s0 * s0 - 2)
        a0, a1, a2, a3 = G[0].clone(), G[1].clone(), G[2].clone(), G[3].clone()
        G[3] = s1 * a3
        G[2] = s1 * a2
        G[1] = s1 * a1 + 2 * s2 * Z[1] * a3
        G[0] = s1 * a0 + s2 * (Z[1] * a1 + Z[2] * a2 + Z[3] * a3) + s3 * Z[1] * Z[1] * a3

    def output_forward(self, A, Wout, bout, U):
        U[:] = (A @ Wout.reshape(-1))
        U[0] += bout.reshape(())

    def loss(self, U, NR, NI, NB, u0, wf, wu, lam, eps, Ubar, gwf, gwu, out4):
        u, ux, ut, uxx = U[0], U[1], U[2], U[3]
        Ubar.zero_()
        r = slice(0, NR)
        f = ut[r] - lam * uxx[r] + eps * u[r] ** 3 - eps * u[r]
        w = wf.reshape(-1)
        df = 2 * w * w * f / NR
        Ubar[2, r] = df
        Ubar[3, r] = -lam * df
        Ubar[0, r] = (3 * eps * u[r] * u[r] - eps) * df
        gwf.reshape(-1)[:] = 2 * w * f * f / NR
        i = slice(NR, NR + NI)
        res = u0.reshape(-1) - u[i]
        wu_ = wu.reshape(-1)
        Ubar[0, i] = -2 * wu_ * wu_ * res / NI
        gwu.reshape(-1)[:] = 2 * wu_ * res * res / NI
        lb, ub = slice(NR + NI, NR + NI + NB), slice(NR + NI + NB, NR + NI + 2 * NB)
        dv, dx = u[lb] - u[ub], ux[lb] - ux[ub]
        Ubar[0, lb] = 2 * dv / NB
        Ubar[0, ub] = -2 * dv / NB
        Ubar[1, lb] = 2 * dx / NB
        Ubar[1, ub] = -2 * dx / NB
        mse_b = torch.mean(dv * dv) + torch.mean(dx * dx)
        out4[0] = torch.mean((wu_ * res) ** 2) + mse_b + torch.mean((w * f) ** 2)
        out4[1] = torch.mean(res * res)
        out4[2] = mse_b
        out4[3] = torch.mean(f * f)

    def output_backward(self, Ubar, Wout, G):
        G[:] = Ubar[:, :, None] * Wout.reshape(1, 1, -1)

    def map_forward(self, A, Wn, bias, Z_out, A_out):
        W = A.shape[2]
        Z_out.reshape(-1, W)[:] = A.reshape(-1, W) @ Wn.t()
        self.act_forward(Z_out, bias, A_out)

    def map_adjoint(self, G, Wn, Zprev, G_out):
        W = G.shape[2]
        G_out.reshape(-1, W)[:] = G.reshape(-1, W) @ Wn
        self.act_backward(Zprev, G_out)

    def wgrad(self, G, A, out):
        W = G.shape[2]
        out[:] = G.reshape(-1, W).t() @ A.reshape(-1, W)

    def bias_grad(self, G, gb):
        gb[:] = G[0].sum(0)

    def layer0_grad(self, G, xt, gW0, gb0):
        # first map: inputs (x, t) with channels (x,1,0,0) and (t,0,1,0)
        gb0[:] = G[0].sum(0)
        gW0[:, 0] = G[0].t() @ xt[:, 0] + G[1].sum(0)
        gW0[:, 1] = G[0].t() @ xt[:, 1] + G[2].sum(0)

    def tick(self, step):
        step += 1

    def adam(self, p, m, v, g, step, lr, beta1, beta2, eps, sign):
        keras_adam_step(p, m, v, sign * g, int(step.item()), lr, beta1, beta2, eps)


def keras_adam_step(p, m, v, g, t, lr, beta1, beta2, eps):
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m += (g - m) * (1.0 - beta1)
    v += (g * g - v) * (1.0 - beta2)
    p -= lr_t * m / (torch.sqrt(v) + eps)


def lbfgs_reference(opfunc, x, max_iter, learning_rate, history=50, tolFun=1e-5, tolX=1e-9):
    """Literal port of eager_lbfgs.py:14-200.  opfunc(x) -> (f, g).  Returns (x, f_hist).  (tolFun / tolX are the reference's
    constants; tests move them to reach the break statements.)"""
    maxEval = max_iter * 1.25
    f, g = opfunc(x)
    f_hist = [float(f)]
    evals = 1
    if float(g.abs().sum()) <= tolFun:
        return x, f_hist
    old_dirs, old_stps, Hdiag = [], [], 1.0
    n_iter = 0
    d = t = g_old = None
    while n_iter < max_iter:
        n_iter += 1
        if n_iter == 1:
            d = -g
        else:
            y = g - g_old
            s = d * t
            ys = float(torch.dot(y, s))
            if ys > 1e-10:
                if len(old_dirs) == history:
                    del old_dirs[0]
                    del old_stps[0]
                old_dirs.append(s)
                old_stps.append(y)
                Hdiag = ys / float(torch.dot(y, y))
            k = len(old_dirs)
            ro = [1.0 / float(torch.dot(old_stps[i], old_dirs[i])) for i in range(k)]
            al = [0.0] * k
            q = -g
            for i in range(k - 1, -1, -1):
                al[i] = float(torch.dot(old_dirs[i], q)) * ro[i]
                q = q - al[i] * old_stps[i]
            r = q * Hdiag
            for i in range(k):
                be_i = float(torch.dot(old_stps[i], r)) * ro[i]
                r = r + (al[i] - be_i) * old_dirs[i]
            d = r
        g_old = g.clone()
        gtd = float(torch.dot(g, d))
        if gtd > -tolX:
            break
        t = min(1.0, 1.0 / float(g.abs().sum())) if n_iter == 1 else learning_rate
        x = x + t * d
        if n_iter != max_iter:
            f, g = opfunc(x)
        evals += 1
        f_hist.append(float(f))
        if n_iter == max_iter or evals >= maxEval:
            break
        if float(g.abs().sum()) <= tolFun:
            break
        if float((d * t).abs().sum()) <= tolX:
            break
        if abs(float(f)) < tolX:            # eager_lbfgs.py:171 `tf.abs(f, f_old) < tolX`: the second argument is the op name
            break
    return x, f_hist
