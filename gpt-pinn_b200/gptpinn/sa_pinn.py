"""Self-adaptive PINN trainer for Allen-Cahn without TensorFlow.

Reference: the TensorFlow / Keras code in Allen-Cahn/AC_main.py:204-339 (= AC_SA_test.py:109-245) and the fixed-step
L-BFGS of eager_lbfgs.py:14-200.  What is reproduced, statement by statement:
  * network   tanh MLP `layers` ([2,128,128,128,128,1]), Glorot-normal weights (truncated normal, seeded), zero biases;
  * loss      mean((w_u (u0 - u(x,0)))^2) + mean((u_lb - u_ub)^2) + mean((u_x,lb - u_x,ub)^2) + mean((w_f f)^2),
              f = u_t - lambda u_xx + eps u^3 - eps u with the true u_xx (AC_main.py:215-244);
  * fit       `tf_iter` Adam steps (lr 5e-3, beta_1 0.99, Keras epsilon 1e-7) on the weights with SIMULTANEOUS gradient
              ASCENT Adam steps on the self-adaptive weights w_f [N_f,1] ~ U(0,1), w_u [N_0,1] ~ 100 U(0,1)
              (AC_main.py:282-298), then up to `newton_iter` L-BFGS iterations (step 0.8, history 50, no line search) on
              the weights only.
How it runs on the B200: all points go through the network as one batch with four channels (u, u_x, u_t, u_xx) in
[4][N][W] tensors; the hidden->hidden maps with the activation in their epilogue (forward and adjoint), the weight- and
bias-gradient contractions over the 4N rows, the loss / seed computation, the Adam updates and the L-BFGS two-loop recursion
are the CUDA kernels of csrc/sa_pinn.cu (C ABI gptpinn_sa_*).  torch.mm is left for the 1 x 4N output-layer gradient and,
composed with the elementwise activation kernels, for hidden widths other than the reference's 128.  One whole Adam iteration (forward, reverse, three optimizer steps) is captured once into a CUDA graph
and replayed: no Python, no allocation, no host sync per step; the iteration counter of the bias correction lives on the
device.  L-BFGS keeps its history on the device (a ring of (s, y) pairs), gets its direction from one kernel launch and
synchronises once per iteration for the reference's scalar tests.  No TensorFlow, no autograd, no CPU path.
Parity: TensorFlow is not in this image, so trajectories are statistical only; the loss and every gradient are held to an
fp64 autograd restatement of the same loss in the test suite (tests/test_sa_pinn.py).
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib


class CudaOps:
    """The gptpinn_sa_* kernels on the current torch stream."""

    def __init__(self, device):
        self.dev = device
        L = _lib.lib()
        self.L = L
        self.ws = torch.zeros(int(L.gptpinn_sa_loss_workspace_floats()), dtype=torch.float32, device=device)
        self.ws2 = torch.empty(int(L.gptpinn_sa_workspace_floats()), dtype=torch.float32, device=device)     # wgrad / column-sum partials

    def _s(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def layer0(self, xt, W0, b0, Z):
        _lib.check(self.L.gptpinn_sa_layer0(self._p(xt), C.c_longlong(xt.shape[0]), self._p(W0), self._p(b0), C.c_int(W0.shape[0]),
                                            self._p(Z), self._s()), "gptpinn_sa_layer0")

    def act_forward(self, Z, bias, A):
        _lib.check(self.L.gptpinn_sa_act_forward(self._p(Z), self._p(bias) if bias is not None else None, self._p(A),
                                                 C.c_longlong(Z.shape[1]), C.c_int(Z.shape[2]), self._s()), "gptpinn_sa_act_forward")

    def act_backward(self, Z, G):
        _lib.check(self.L.gptpinn_sa_act_backward(self._p(Z), self._p(G), C.c_longlong(Z.shape[1]), C.c_int(Z.shape[2]), self._s()),
                   "gptpinn_sa_act_backward")

    def output_forward(self, A, Wout, bout, U):
        _lib.check(self.L.gptpinn_sa_output_forward(self._p(A), self._p(Wout), self._p(bout), C.c_longlong(A.shape[1]), C.c_int(A.shape[2]),
                                                    self._p(U), self._s()), "gptpinn_sa_output_forward")

    def loss(self, U, NR, NI, NB, u0, wf, wu, lam, eps, Ubar, gwf, gwu, out4):
        _lib.check(self.L.gptpinn_sa_loss(self._p(U), C.c_int(NR), C.c_int(NI), C.c_int(NB), self._p(u0), self._p(wf), self._p(wu),
                                          C.c_double(lam), C.c_double(eps), self._p(Ubar), self._p(gwf), self._p(gwu), self._p(out4),
                                          self._p(self.ws), self._s()), "gptpinn_sa_loss")

    def output_backward(self, Ubar, Wout, G):
        _lib.check(self.L.gptpinn_sa_output_backward(self._p(Ubar), self._p(Wout), C.c_longlong(G.shape[1]), C.c_int(G.shape[2]), self._p(G),
                                                     self._s()), "gptpinn_sa_output_backward")

    def tick(self, step):
        _lib.check(self.L.gptpinn_sa_tick(self._p(step), self._s()), "gptpinn_sa_tick")

    def map_forward(self, A, Wn, bias, Z_out, A_out):
        """Z_out = A Wn^T (+ bias on channel 0), A_out = channels of tanh(Z_out): one kernel for W = 128 (the reference's width)"""
        W = A.shape[2]
        if W == 128 and not os.environ.get("GPTPINN_SA_NO_FUSED_MAPS"):
            _lib.check(self.L.gptpinn_sa_map_forward(self._p(A), self._p(Wn), self._p(bias), C.c_longlong(A.shape[1]), C.c_int(W), self._p(Z_out),
                                                     self._p(A_out), self._s()), "gptpinn_sa_map_forward")
        else:                                         # other widths: library GEMM + the elementwise kernel
            torch.mm(A.view(-1, W), Wn.t(), out=Z_out.view(-1, W))
            self.act_forward(Z_out, bias, A_out)

    def map_adjoint(self, G, Wn, Zprev, G_out):
        """G_out = adjoint of the previous layer's pre-activations = act_backward(Zprev, G Wn)"""
        W = G.shape[2]
        if W == 128 and not os.environ.get("GPTPINN_SA_NO_FUSED_MAPS"):
            _lib.check(self.L.gptpinn_sa_map_adjoint(self._p(G), self._p(Wn), self._p(Zprev), C.c_longlong(G.shape[1]), C.c_int(W), self._p(G_out),
                                                     self._s()), "gptpinn_sa_map_adjoint")
        else:
            torch.mm(G.view(-1, W), Wn, out=G_out.view(-1, W))
            self.act_backward(Zprev, G_out)

    def wgrad(self, G, A, out):
        """out[j][i] = sum over channels and points of G[c][r][j] A[c][r][i]: the weight gradient of a hidden->hidden map"""
        W = G.shape[2]
        if W > 128:                                   # wider than the kernel's register tile: library GEMM
            torch.mm(G.view(-1, W).t(), A.view(-1, W), out=out)
            return
        _lib.check(self.L.gptpinn_sa_wgrad(self._p(G), self._p(A), C.c_longlong(G.shape[0] * G.shape[1]), C.c_int(W), self._p(out),
                                           self._p(self.ws2), self._s()), "gptpinn_sa_wgrad")

    def bias_grad(self, G, gb):
        """gb[j] = sum over the points of G[0][r][j]"""
        if G.shape[2] > 256:                          # wider than the kernel's column mapping: library reduction
            torch.sum(G[0], dim=0, out=gb)
            return
        _lib.check(self.L.gptpinn_sa_bias_grad(self._p(G), C.c_longlong(G.shape[1]), C.c_int(G.shape[2]), self._p(gb), self._p(self.ws2), self._s()),
                   "gptpinn_sa_bias_grad")

    def layer0_grad(self, G, xt, gW0, gb0):
        """gradients of the first map from the adjoints of its pre-activations (inputs (x, t), channels (x,1,0,0) / (t,0,1,0))"""
        if G.shape[2] > 256:
            torch.sum(G[0], dim=0, out=gb0)
            gW0[:, 0] = torch.mv(G[0].t(), xt[:, 0]) + torch.sum(G[1], dim=0)
            gW0[:, 1] = torch.mv(G[0].t(), xt[:, 1]) + torch.sum(G[2], dim=0)
            return
        _lib.check(self.L.gptpinn_sa_layer0_grad(self._p(G), self._p(xt), C.c_longlong(G.shape[1]), C.c_int(G.shape[2]), self._p(gW0), self._p(gb0),
                                                 self._p(self.ws2), self._s()), "gptpinn_sa_layer0_grad")

    def lbfgs_capacity(self):
        return int(self.L.gptpinn_sa_lbfgs_max_params())

    def lbfgs_direction(self, S, Y, ro, k, start, Hdiag, g, d):
        """d = -H g from the k most recent (s, y) pairs (ring rows start, start + 1, ... mod S.shape[0]): the whole two-loop
        recursion in one launch"""
        _lib.check(self.L.gptpinn_sa_lbfgs_direction(self._p(S), self._p(Y), self._p(ro), C.c_int(k), C.c_int(start), C.c_int(S.shape[0]),
                                                     C.c_longlong(S.stride(0)), C.c_longlong(g.numel()), C.c_double(Hdiag), self._p(g),
                                                     self._p(d), self._s()), "gptpinn_sa_lbfgs_direction")

    def adam(self, p, m, v, g, step, lr, beta1, beta2, eps, sign):
        _lib.check(self.L.gptpinn_sa_adam(self._p(p), self._p(m), self._p(v), self._p(g), C.c_longlong(p.numel()), self._p(step),
                                          C.c_double(lr), C.c_double(beta1), C.c_double(beta2), C.c_double(eps), C.c_int(1 if sign < 0 else 0),
                                          self._s()), "gptpinn_sa_adam")


def glorot_normal(fan_out, fan_in, generator):
    """Keras GlorotNormal: truncated normal (|z| <= 2) with stddev sqrt(2 / (fan_in + fan_out)) / 0.87962566103423978."""
    std = math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
    w = torch.empty(fan_out, fan_in)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=generator)
    return (w * std).float()


class SAPinn:
    """One self-adaptive PINN: network weights + self-adaptive weights + everything fit() needs, resident on the device."""

    def __init__(self, layers, xt_f, xt0, u0, xt_lb, xt_ub, lmbda, eps, col_weights=None, u_weights=None, seed=0, ops=None,
                 dtype=torch.float32):
        dev = xt_f.device
        if ops is None:
            if dev.type != "cuda":
                raise _lib.GptPinnError("SAPinn: tensors must live on the GPU (no CPU path)")
            ops = CudaOps(dev)
        self.ops, self.dev, self.dtype = ops, dev, dtype
        self.layers = [int(x) for x in layers]
        hidden = self.layers[1:-1]
        if self.layers[0] != 2 or self.layers[-1] != 1 or len(hidden) < 1 or len(set(hidden)) != 1:
            raise ValueError("SAPinn needs a [2, W, ..., W, 1] network")
        self.W, self.NH = hidden[0], len(hidden)
        self.lmbda, self.eps = float(lmbda), float(eps)
        self.NR, self.NI, self.NB = xt_f.shape[0], xt0.shape[0], xt_lb.shape[0]
        if xt_ub.shape[0] != self.NB:
            raise ValueError("lower and upper boundary sets must pair up")
        self.N = self.NR + self.NI + 2 * self.NB
        f = lambda t: t.detach().to(dev, dtype).contiguous()
        self.xt = torch.cat((f(xt_f), f(xt0), f(xt_lb), f(xt_ub)), 0).contiguous()
        self.u0 = f(u0).reshape(-1)
        g = torch.Generator().manual_seed(seed)
        # AC_main.py:347-348: col_weights ~ U(0,1) [N_f,1], u_weights ~ 100 U(0,1) [N0,1]
        self.wf = f(col_weights).reshape(-1).clone() if col_weights is not None else torch.rand(self.NR, generator=g).to(dev, dtype)
        self.wu = f(u_weights).reshape(-1).clone() if u_weights is not None else (100.0 * torch.rand(self.NI, generator=g)).to(dev, dtype)
        # flat parameter vector: per layer W (nn.Linear layout [out][in], row-major) then b
        sizes = []
        for fi, fo in zip(self.layers[:-1], self.layers[1:]):
            sizes += [fo * fi, fo]
        self.P = sum(sizes)
        self.theta = torch.zeros(self.P, dtype=dtype, device=dev)
        self.grad = torch.zeros(self.P, dtype=dtype, device=dev)
        self.Wv, self.bv, self.gW, self.gb = [], [], [], []
        off = 0
        for fi, fo in zip(self.layers[:-1], self.layers[1:]):
            self.Wv.append(self.theta[off:off + fo * fi].view(fo, fi)); self.gW.append(self.grad[off:off + fo * fi].view(fo, fi)); off += fo * fi
            self.bv.append(self.theta[off:off + fo]); self.gb.append(self.grad[off:off + fo]); off += fo
        for Wl in self.Wv:
            Wl.copy_(glorot_normal(Wl.shape[0], Wl.shape[1], g).to(dev, dtype))
        # workspaces (allocated once: an iteration never allocates, so it can live in a CUDA graph)
        N, W = self.N, self.W
        z = lambda *s: torch.zeros(*s, dtype=dtype, device=dev)
        self.Z = [z(4, N, W) for _ in range(self.NH)]
        self.A = [z(4, N, W) for _ in range(self.NH)]
        self.G = z(4, N, W)
        self.G2 = z(4, N, W)
        self.U, self.Ubar = z(4, N), z(4, N)
        self.gwf, self.gwu = z(self.NR), z(self.NI)
        self.loss4 = z(4)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.m_t, self.v_t = z(self.P), z(self.P)
        self.m_f, self.v_f = z(self.NR), z(self.NR)
        self.m_u, self.v_u = z(self.NI), z(self.NI)
        self._graph = None

    # ------------------------------------------------------------------ loss and gradients
    def loss_and_grad(self):
        """Fills self.loss4 = (total, mse_u, mse_b, mse_f), self.grad (flat, d total / d weights), self.gwf, self.gwu."""
        o, N, W, NH = self.ops, self.N, self.W, self.NH
        o.layer0(self.xt, self.Wv[0], self.bv[0], self.Z[0])
        o.act_forward(self.Z[0], None, self.A[0])
        for l in range(NH - 1):
            o.map_forward(self.A[l], self.Wv[l + 1], self.bv[l + 1], self.Z[l + 1], self.A[l + 1])
        o.output_forward(self.A[NH - 1], self.Wv[NH], self.bv[NH], self.U)
        o.loss(self.U, self.NR, self.NI, self.NB, self.u0, self.wf, self.wu, self.lmbda, self.eps, self.Ubar, self.gwf, self.gwu, self.loss4)
        # reverse
        torch.mm(self.Ubar.view(1, 4 * N), self.A[NH - 1].view(4 * N, W), out=self.gW[NH])
        torch.sum(self.Ubar[0], dim=0, keepdim=True, out=self.gb[NH])
        G, G2 = self.G, self.G2
        o.output_backward(self.Ubar, self.Wv[NH], G)
        o.act_backward(self.Z[NH - 1], G)                 # G: adjoint of the pre-activations of the last hidden layer
        for l in range(NH - 1, -1, -1):
            if l > 0:
                o.bias_grad(G, self.gb[l])
                o.wgrad(G, self.A[l - 1], self.gW[l])
                o.map_adjoint(G, self.Wv[l], self.Z[l - 1], G2)
                G, G2 = G2, G
            else:
                o.layer0_grad(G, self.xt, self.gW[0], self.gb[0])
        return self.loss4

    # ------------------------------------------------------------------ Adam phase (AC_main.py:282-298)
    def _adam_iteration(self, lr, beta1):
        self.loss_and_grad()
        self.ops.tick(self.step)
        self.ops.adam(self.theta, self.m_t, self.v_t, self.grad, self.step, lr, beta1, 0.999, 1e-7, +1)
        self.ops.adam(self.wf, self.m_f, self.v_f, self.gwf, self.step, lr, beta1, 0.999, 1e-7, -1)
        self.ops.adam(self.wu, self.m_u, self.v_u, self.gwu, self.step, lr, beta1, 0.999, 1e-7, -1)

    def adam_fit(self, iters, lr=0.005, beta1=0.99, record=False, use_graph=True, verbose=False):
        """`iters` simultaneous steps: Adam descent on the weights, Adam ascent on w_f and w_u.  Returns the recorded
        total losses (device tensor [iters]) when record, else None."""
        rec = torch.zeros(max(iters, 1), dtype=self.dtype, device=self.dev) if record else None

        def after(it):
            if record:
                rec[it] = self.loss4[0]
            if verbose and (it % 1000 == 0 or it == iters - 1):
                l4 = self.loss4.tolist()
                print(f"Epoch: {it}\nmse_0: {l4[1]} | mse_b: {l4[2]} | mse_f: {l4[3]} | Total Loss: {l4[0]}\n")

        it, graph = 0, None
        if use_graph and self.dev.type == "cuda" and iters > 8:
            # two real iterations on a side stream first (cuBLAS workspaces, lazy initialisation), then ONE iteration is
            # captured (capturing does not execute it) and replayed for the rest of the phase
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._adam_iteration(lr, beta1)
                    after(it)
                    it += 1
            torch.cuda.current_stream(self.dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._adam_iteration(lr, beta1)
        while it < iters:
            if graph is not None:
                graph.replay()
            else:
                self._adam_iteration(lr, beta1)
            after(it)
            it += 1
        return rec

    # ------------------------------------------------------------------ L-BFGS phase (eager_lbfgs.py:14-200)
    def lbfgs_fit(self, max_iter, learning_rate=0.8, history=50, record=False, verbose=False, use_fused=True, tolFun=1e-5, tolX=1e-9):
        """Fixed-step L-BFGS on the network weights (w_f, w_u frozen): the two-loop recursion and the stopping tests of
        eager_lbfgs.py:14-200, statement for statement in their effect, arranged so that an iteration reads the device ONCE:

          * the (s, y) pairs live in a ring of `history` rows (no shifting), and the direction d = -H g comes from one kernel
            launch (gptpinn_sa_lbfgs_direction) instead of 4k small tensor operations;
          * the step x += t d is taken and the loss / gradient at the new point evaluated BEFORE the host sees gtd = <g, d>;
            the six scalars an iteration decides on (gtd, sum|d|, f, sum|g|, <y, s>, <y, y>) come back in one transfer.  Should
            gtd fail the reference's progress test (gtd > -tolX: break before the step) the saved x is put back and the state
            re-evaluated there, which is what the reference's loop leaves behind.

        Returns the list of losses (f_hist)."""
        dt, dev, P = self.dtype, self.dev, self.P
        fused = (use_fused and hasattr(self.ops, "lbfgs_direction") and dt == torch.float32 and history <= 256
                 and P <= self.ops.lbfgs_capacity())
        x = self.theta
        ld = (P + 3) & ~3                                                          # 16-byte aligned rows for the kernel's float4 path
        S = torch.zeros(history, ld, dtype=dt, device=dev)[:, :P]                  # old_dirs  (ring)
        Y = torch.zeros(history, ld, dtype=dt, device=dev)[:, :P]                  # old_stps  (ring)
        ro = torch.zeros(history, dtype=dt, device=dev)
        k, head = 0, 0                                                             # pairs held; ring row of the oldest
        Hdiag = 1.0
        f_hist = []
        if max_iter <= 0:
            return f_hist
        self.loss_and_grad()
        g = self.grad.clone()
        f, gsum = torch.stack((self.loss4[0], g.abs().sum())).tolist()
        f_hist.append(f)
        evals = 1
        if gsum <= tolFun:
            return f_hist
        x_prev = torch.empty_like(x)
        s_buf = y_buf = None                                                       # s = d t, y = g - g_old of the step just taken
        ys = yy = 0.0
        n_iter = 0
        while n_iter < max_iter:
            n_iter += 1
            if n_iter == 1:
                d = -g
                t = min(1.0, 1.0 / gsum)
            else:
                if ys > 1e-10:
                    if k == history:                      # limited memory: the newest pair replaces the oldest
                        row, head = head, (head + 1) % history
                    else:
                        row, k = (head + k) % history, k + 1
                    S[row].copy_(s_buf); Y[row].copy_(y_buf)
                    ro[row] = 1.0 / ys
                    Hdiag = ys / yy
                if fused:
                    d = torch.empty_like(g)
                    self.ops.lbfgs_direction(S, Y, ro, k, head, Hdiag, g, d)
                else:
                    # two-loop recursion as tensor expressions (fp64 host-logic tests; vectors beyond one CTA's shared memory)
                    rows = [(head + i) % history for i in range(k)]
                    al = [None] * k
                    q = -g
                    for i in range(k - 1, -1, -1):
                        al[i] = torch.dot(S[rows[i]], q) * ro[rows[i]]
                        q = q - al[i] * Y[rows[i]]
                    r = q * Hdiag
                    for i in range(k):
                        be_i = torch.dot(Y[rows[i]], r) * ro[rows[i]]
                        r = r + (al[i] - be_i) * S[rows[i]]
                    d = r
                t = learning_rate
            last = n_iter == max_iter
            scal = [torch.dot(g, d), d.abs().sum()]
            x_prev.copy_(x)
            x.add_(d, alpha=t)
            if not last:                                  # eager_lbfgs.py:131: no re-evaluation in the last iteration
                self.loss_and_grad()
                g_new = self.grad.clone()
                y_buf = g_new - g
                s_buf = d * t
                scal += [self.loss4[0], g_new.abs().sum(), torch.dot(y_buf, s_buf), torch.dot(y_buf, y_buf)]
            vals = torch.stack(scal).tolist()                                      # the iteration's one host sync
            gtd_h, dsum_h = vals[0], vals[1]
            if gtd_h > -tolX:                             # no progress along d: the reference breaks BEFORE the step
                x.copy_(x_prev)
                if not last:
                    self.loss_and_grad()
                break
            if not last:
                f, gsum, ys, yy = vals[2:6]
                g = g_new
            evals += 1
            f_hist.append(f)
            if verbose and (n_iter % 1000 == 0 or n_iter == max_iter):
                print("Epoch: %3d | Loss: %6.5f " % (n_iter, f))
            if n_iter == max_iter or evals >= max_iter * 1.25:
                break
            if gsum <= tolFun:
                break
            if dsum_h * abs(t) <= tolX:
                break
            if abs(f) < tolX:                 # eager_lbfgs.py:171 `tf.abs(f, f_old) < tolX` == |f| < tolX
                break
        return f_hist

    def fit(self, tf_iter, newton_iter, lr_adam=0.005, lr_lbfgs=0.8, record=False, verbose=True):
        """AC_main.py:275-338 `fit`."""
        if verbose:
            print("Starting ADAM training")
        a = self.adam_fit(tf_iter, lr=lr_adam, record=record, verbose=verbose)
        if verbose:
            print("Starting L-BFGS training")
        b = self.lbfgs_fit(newton_iter, learning_rate=lr_lbfgs, record=record, verbose=verbose)
        return (a.cpu().numpy() if a is not None else None), np.asarray(b)

    # ------------------------------------------------------------------ hand-over to the GPT-PINN side
    def weights_for_P(self):
        """(SA_w, SA_b) exactly as AC_main.py:371-376 builds them for AC_models.P: torch [out, in] weights, [out] biases."""
        return [w.detach().clone() for w in self.Wv], [b.detach().clone() for b in self.bv]

    def predict(self, xt):
        """u at arbitrary points (AC_main.py:331-334 `predict`), through the same kernels."""
        from . import precomp
        u = torch.empty(xt.shape[0], dtype=torch.float32, device=self.dev)
        precomp.mlp_channels(list(zip(self.Wv, self.bv)), "tanh", xt.to(self.dev, torch.float32), v=u)
        return u
