"""Self-adaptive PINN trainer (reference Allen-Cahn/AC_main.py:204-339, eager_lbfgs.py:14-200).

TensorFlow is not in this image, so the reference trainer cannot be run: parity is per step against an fp64 autograd
restatement of the same loss (oracle/sa_pinn_ref.py) and, for L-BFGS, against a literal port of eager_lbfgs.py.  The CPU
tests drive the product's orchestration (gptpinn/sa_pinn.py) through a torch stand-in for the CUDA kernels; the GPU tests
hold every gptpinn_sa_* kernel and the whole loss / gradient to the same references."""
import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401  (puts the package on sys.path)


def _problem(NR, NI, NB, dtype, device="cpu", seed=0):
    g = torch.Generator().manual_seed(seed)
    xt_f = torch.rand(NR, 2, generator=g, dtype=torch.float64) * torch.tensor([2.0, 1.0]) - torch.tensor([1.0, 0.0])
    x0 = torch.rand(NI, 1, generator=g, dtype=torch.float64) * 2 - 1
    xt0 = torch.cat((x0, torch.zeros(NI, 1, dtype=torch.float64)), 1)
    u0 = x0 ** 2 * torch.cos(np.pi * x0)                     # the Allen-Cahn initial condition of AC.mat
    tb = torch.rand(NB, 1, generator=g, dtype=torch.float64)
    xt_lb = torch.cat((-torch.ones(NB, 1, dtype=torch.float64), tb), 1)
    xt_ub = torch.cat((torch.ones(NB, 1, dtype=torch.float64), tb), 1)
    return [t.to(device, dtype) for t in (xt_f, xt0, u0, xt_lb, xt_ub)]


def _autograd(m, prob, lam, eps, dtype=torch.float64):
    from oracle import sa_pinn_ref as R
    xt_f, xt0, u0, xt_lb, xt_ub = [t.detach().cpu().to(dtype) for t in prob]
    params = [(W.detach().cpu().to(dtype).clone().requires_grad_(), b.detach().cpu().to(dtype).clone().requires_grad_()) for W, b in zip(m.Wv, m.bv)]
    wf = m.wf.detach().cpu().to(dtype).reshape(-1, 1).clone().requires_grad_()
    wu = m.wu.detach().cpu().to(dtype).reshape(-1, 1).clone().requires_grad_()
    out = R.sa_loss_autograd(params, xt_f, xt0, u0, xt_lb, xt_ub, wf, wu, lam, eps)
    out[0].backward()
    flat = torch.cat([torch.cat((W.grad.reshape(-1), b.grad.reshape(-1))) for W, b in params])
    return [float(x) for x in out], flat, wf.grad.reshape(-1), wu.grad.reshape(-1)


def test_hand_derived_reverse_pass_equals_autograd_of_the_reference_loss_cpu_fp64():
    from gptpinn.sa_pinn import SAPinn
    from oracle import sa_pinn_ref as R
    prob = _problem(57, 13, 7, torch.float64)
    m = SAPinn([2, 16, 16, 16, 1], *prob, 5.5e-4, 3.0, seed=3, ops=R.TorchOps(), dtype=torch.float64)
    for b in m.bv:
        b.copy_(0.1 * torch.randn(b.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1)))
    m.loss_and_grad()
    l4, flat, gwf, gwu = _autograd(m, prob, 5.5e-4, 3.0)
    assert np.allclose(m.loss4.tolist(), l4, rtol=1e-12)
    assert float((flat - m.grad).abs().max() / flat.abs().max()) < 1e-12
    assert float((gwf - m.gwf).abs().max() / gwf.abs().max()) < 1e-12
    assert float((gwu - m.gwu).abs().max() / gwu.abs().max()) < 1e-12


def test_adam_phase_descends_on_weights_and_ascends_on_sa_weights_cpu():
    """AC_main.py:297-298: the network weights take Keras-Adam steps on +grad, the self-adaptive weights on -grad."""
    from gptpinn.sa_pinn import SAPinn
    from oracle import sa_pinn_ref as R
    prob = _problem(40, 9, 5, torch.float64)
    m = SAPinn([2, 8, 8, 1], *prob, 1e-3, 2.0, seed=1, ops=R.TorchOps(), dtype=torch.float64)
    th0, wf0, wu0 = m.theta.clone(), m.wf.clone(), m.wu.clone()
    m.loss_and_grad()
    g0, gf0, gu0 = m.grad.clone(), m.gwf.clone(), m.gwu.clone()
    m.adam_fit(1, use_graph=False)
    # first Keras-Adam step in closed form: lr_t m / (sqrt(v) + eps) = lr g / (|g| + eps / sqrt(1 - beta_2))
    first = lambda g: 0.005 * g / (g.abs() + 1e-7 / (1 - 0.999) ** 0.5)
    assert torch.allclose(m.theta - th0, -first(g0), rtol=1e-9, atol=1e-15)          # descent on the network weights
    assert torch.allclose(m.wf - wf0, first(gf0), rtol=1e-9, atol=1e-15)             # ASCENT on the collocation weights
    assert torch.allclose(m.wu - wu0, first(gu0), rtol=1e-9, atol=1e-15)             # ASCENT on the initial-condition weights
    rec = m.adam_fit(40, record=True, use_graph=False)
    assert rec.shape == (40,) and int(m.step.item()) == 41


def test_lbfgs_equals_the_literal_port_of_eager_lbfgs_cpu():
    from gptpinn.sa_pinn import SAPinn
    from oracle import sa_pinn_ref as R
    prob = _problem(57, 13, 7, torch.float64)
    m = SAPinn([2, 12, 12, 1], *prob, 5.5e-4, 3.0, seed=5, ops=R.TorchOps(), dtype=torch.float64)
    m.adam_fit(20, use_graph=False)
    theta0 = m.theta.clone()

    def opfunc(x):
        m.theta.copy_(x)
        m.loss_and_grad()
        return m.loss4[0].clone(), m.grad.clone()
    xr, fr = R.lbfgs_reference(opfunc, theta0.clone(), 70, 0.8, history=50)
    m.theta.copy_(theta0)
    fh = m.lbfgs_fit(70, 0.8, history=50, record=True)
    assert len(fh) == len(fr) and np.allclose(fh, fr, rtol=1e-8)
    assert float((m.theta - xr).abs().max()) < 1e-8
    # history shorter than the run: the ring shift of eager_lbfgs.py:70-73
    m.theta.copy_(theta0)
    xr, fr = R.lbfgs_reference(opfunc, theta0.clone(), 30, 0.8, history=5)
    m.theta.copy_(theta0)
    fh = m.lbfgs_fit(30, 0.8, history=5, record=True)
    assert np.allclose(fh, fr, rtol=1e-8)
    # the break statements (eager_lbfgs.py:118-120, 150-174), reached by moving the reference's two tolerances: no progress along
    # d in the first / a later iteration (the step taken ahead of the test is put back), small gradient, small step, and the
    # last iteration (no re-evaluation, eager_lbfgs.py:131)
    early = 0
    for kw, iters in ((dict(tolX=1e30), 20), (dict(tolFun=1e30), 20), (dict(tolFun=5000.0), 30), (dict(tolFun=1000.0), 30), (dict(tolX=0.5), 30),
                      (dict(tolX=5.0), 30), (dict(), 1), (dict(), 2), (dict(), 3)):
        xr, fr = R.lbfgs_reference(opfunc, theta0.clone(), iters, 0.8, history=7, **kw)
        m.theta.copy_(theta0)
        fh = m.lbfgs_fit(iters, 0.8, history=7, record=True, **kw)
        assert len(fh) == len(fr) and np.allclose(fh, fr, rtol=1e-9), (kw, iters, len(fh), len(fr))
        assert float((m.theta - xr).abs().max()) < 1e-12, (kw, iters)
        early += len(fh) < iters + 1
    assert early == 6


def test_sa_pinn_has_no_cpu_path():
    from gptpinn import GptPinnError
    from gptpinn.sa_pinn import SAPinn
    with pytest.raises(GptPinnError):
        SAPinn([2, 8, 8, 1], *_problem(10, 4, 3, torch.float32), 1e-3, 2.0)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_every_sa_kernel_matches_the_torch_recurrences_gpu():
    from gptpinn.sa_pinn import CudaOps
    from oracle import sa_pinn_ref as R
    dev = torch.device("cuda")
    cu, th = CudaOps(dev), R.TorchOps()
    g = torch.Generator().manual_seed(0)
    N, W = 777, 128
    r = lambda *s: torch.randn(*s, generator=g)
    xt, W0, b0 = torch.rand(N, 2, generator=g) * 2 - 1, r(W, 2), 0.1 * r(W)
    Zc, Zt = torch.zeros(4, N, W, device=dev), torch.zeros(4, N, W, dtype=torch.float64)
    cu.layer0(xt.to(dev), W0.to(dev), b0.to(dev), Zc)
    th.layer0(xt.double(), W0.double(), b0.double(), Zt)
    assert torch.allclose(Zc.cpu().double(), Zt, rtol=1e-6, atol=1e-6)
    Z = r(4, N, W)
    bias = 0.1 * r(W)
    Ac, At = torch.zeros(4, N, W, device=dev), torch.zeros(4, N, W, dtype=torch.float64)
    Zc, Zt = Z.clone().to(dev), Z.clone().double()
    cu.act_forward(Zc, bias.to(dev), Ac)
    th.act_forward(Zt, bias.double(), At)
    assert torch.allclose(Zc.cpu().double(), Zt, atol=1e-6) and torch.allclose(Ac.cpu().double(), At, rtol=1e-5, atol=2e-6)
    G = r(4, N, W)
    Gc, Gt = G.clone().to(dev), G.clone().double()
    cu.act_backward(Zc, Gc)
    th.act_backward(Zt, Gt)
    assert torch.allclose(Gc.cpu().double(), Gt, rtol=1e-5, atol=1e-5)
    Wout, bout = r(1, W) / W ** 0.5, r(1)
    Uc, Ut = torch.zeros(4, N, device=dev), torch.zeros(4, N, dtype=torch.float64)
    cu.output_forward(Ac, Wout.to(dev), bout.to(dev), Uc)
    th.output_forward(At, Wout.double(), bout.double(), Ut)
    assert torch.allclose(Uc.cpu().double(), Ut, rtol=1e-5, atol=1e-5)
    NR, NI, NB = 600, 97, 40
    assert NR + NI + 2 * NB == N
    u0, wf, wu = r(NI), torch.rand(NR, generator=g), 100 * torch.rand(NI, generator=g)
    outs_c = [torch.zeros(4, N, device=dev), torch.zeros(NR, device=dev), torch.zeros(NI, device=dev), torch.zeros(4, device=dev)]
    outs_t = [torch.zeros(4, N, dtype=torch.float64), torch.zeros(NR, dtype=torch.float64), torch.zeros(NI, dtype=torch.float64), torch.zeros(4, dtype=torch.float64)]
    cu.loss(Uc, NR, NI, NB, u0.to(dev), wf.to(dev), wu.to(dev), 5.5e-4, 3.0, *outs_c)
    th.loss(Uc.cpu().double(), NR, NI, NB, u0.double(), wf.double(), wu.double(), float(np.float32(5.5e-4)), 3.0, *outs_t)
    for a, b in zip(outs_c, outs_t):
        assert torch.allclose(a.cpu().double(), b, rtol=2e-5, atol=1e-6 * float(b.abs().max()))
    Gc2, Gt2 = torch.zeros(4, N, W, device=dev), torch.zeros(4, N, W, dtype=torch.float64)
    cu.output_backward(outs_c[0], Wout.to(dev), Gc2)
    th.output_backward(outs_c[0].cpu().double(), Wout.double(), Gt2)
    assert torch.allclose(Gc2.cpu().double(), Gt2, rtol=1e-6, atol=1e-9)
    # hidden->hidden maps with the activation in the epilogue (W = 128: one kernel; other widths: library GEMM + elementwise kernel),
    # forward and adjoint, ragged point counts, against the float64 composition of the same fp32 inputs
    for Nn, Ww in ((777, 128), (20712, 128), (33, 128), (1, 128), (32, 128), (300, 64)):
        Ain, Gin, Zp = 0.5 * r(4, Nn, Ww), r(4, Nn, Ww), r(4, Nn, Ww)
        Wn, bn = r(Ww, Ww) / Ww ** 0.5, 0.1 * r(Ww)
        Zo, Ao = torch.full((4, Nn, Ww), float("nan"), device=dev), torch.full((4, Nn, Ww), float("nan"), device=dev)
        cu.map_forward(Ain.to(dev), Wn.to(dev), bn.to(dev), Zo, Ao)
        Zr, Ar = torch.zeros(4, Nn, Ww, dtype=torch.float64), torch.zeros(4, Nn, Ww, dtype=torch.float64)
        th.map_forward(Ain.double(), Wn.double(), bn.double(), Zr, Ar)
        assert torch.allclose(Zo.cpu().double(), Zr, rtol=1e-5, atol=2e-6), (Nn, Ww)
        assert torch.allclose(Ao.cpu().double(), Ar, rtol=1e-5, atol=5e-6), (Nn, Ww)
        Go = torch.full((4, Nn, Ww), float("nan"), device=dev)
        cu.map_adjoint(Gin.to(dev), Wn.to(dev), Zp.to(dev), Go)
        Gr = torch.zeros(4, Nn, Ww, dtype=torch.float64)
        th.map_adjoint(Gin.double(), Wn.double(), Zp.double(), Gr)
        assert torch.allclose(Go.cpu().double(), Gr, rtol=1e-5, atol=2e-5), (Nn, Ww)
    # weight / bias gradient contractions: the register-tile kernel (W = 128, aligned), its zero-padded path (W < 128, ragged
    # row counts, fewer row tiles than CTAs), the library fallback (W > 128), against float64 sums of the same fp32 inputs
    for Nn, Ww in ((777, 128), (5003, 128), (20712, 128), (7, 128), (333, 48), (1, 12), (90, 160)):
        Gg, Aa = r(4, Nn, Ww).to(dev), r(4, Nn, Ww).to(dev)
        xy = (torch.rand(Nn, 2, generator=g) * 2 - 1).to(dev)
        oc, ot = torch.full((Ww, Ww), float("nan"), device=dev), torch.zeros(Ww, Ww, dtype=torch.float64)
        cu.wgrad(Gg, Aa, oc)
        th.wgrad(Gg.cpu().double(), Aa.cpu().double(), ot)
        oc2 = torch.empty_like(oc)
        cu.wgrad(Gg, Aa, oc2)
        assert torch.equal(oc, oc2) or Ww > 128                                            # fixed summation order
        assert float((oc.cpu().double() - ot).abs().max()) < 2e-6 * (4 * Nn) ** 0.5 * 4 + 1e-6, (Nn, Ww)
        bc, bt_ = torch.full((Ww,), float("nan"), device=dev), torch.zeros(Ww, dtype=torch.float64)
        cu.bias_grad(Gg, bc)
        th.bias_grad(Gg.cpu().double(), bt_)
        assert float((bc.cpu().double() - bt_).abs().max()) < 2e-6 * Nn ** 0.5 * 4 + 1e-6, (Nn, Ww)
        wc, wt_ = torch.full((Ww, 2), float("nan"), device=dev), torch.zeros(Ww, 2, dtype=torch.float64)
        bc2, bt2 = torch.full((Ww,), float("nan"), device=dev), torch.zeros(Ww, dtype=torch.float64)
        cu.layer0_grad(Gg, xy, wc, bc2)
        th.layer0_grad(Gg.cpu().double(), xy.cpu().double(), wt_, bt2)
        assert torch.equal(bc, bc2)
        assert float((wc.cpu().double() - wt_).abs().max()) < 2e-6 * Nn ** 0.5 * 8 + 1e-6, (Nn, Ww)
    # Keras Adam, descent and ascent, three consecutive steps (device-side step counter)
    n = 5001
    p, gr = r(n), r(n)
    for sign in (+1, -1):
        pc, mc, vc = p.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        pt, mt, vt = p.clone().double(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
        step = torch.zeros(1, dtype=torch.int32, device=dev)
        for t in (1, 2, 3):
            cu.tick(step)
            cu.adam(pc, mc, vc, (gr * t).to(dev), step, 0.005, 0.99, 0.999, 1e-7, sign)
            R.keras_adam_step(pt, mt, vt, sign * (gr * t).double(), t, 0.005, 0.99, 0.999, 1e-7)
        assert int(step.item()) == 3
        assert torch.allclose(pc.cpu().double(), pt, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_sa_loss_and_gradients_match_fp64_autograd_at_the_reference_sizes_gpu():
    """[2,128,128,128,128,1], N_f = 20 000, N_0 = 512, N_b = 100 (AC_main.py:49-51, :125): loss 1e-5, gradients 1e-4 of max|g|
    against autograd in float64 on the same weights (fp32 GEMM accumulation over 82 848 rows sets the gradient tolerance)."""
    from gptpinn.sa_pinn import SAPinn
    dev = torch.device("cuda")
    prob = _problem(20000, 512, 100, torch.float32, device=dev)
    m = SAPinn([2, 128, 128, 128, 128, 1], *prob, 5.5e-4, 3.0, seed=0)
    for b in m.bv:
        b.copy_(0.05 * torch.randn(b.shape, generator=torch.Generator().manual_seed(2)).to(dev))
    m.loss_and_grad()
    torch.cuda.synchronize()
    l4, flat, gwf, gwu = _autograd(m, prob, float(np.float32(5.5e-4)), 3.0)
    got = np.array(m.loss4.tolist())
    assert np.max(np.abs(got - np.array(l4)) / np.abs(np.array(l4))) < 1e-5
    assert float((flat - m.grad.cpu().double()).abs().max() / flat.abs().max()) < 1e-4
    assert float((gwf - m.gwf.cpu().double()).abs().max() / gwf.abs().max()) < 1e-4
    assert float((gwu - m.gwu.cpu().double()).abs().max() / gwu.abs().max()) < 1e-4


@pytest.mark.gpu
def test_sa_fit_graph_replay_equals_eager_and_training_makes_progress_gpu():
    from gptpinn.sa_pinn import SAPinn
    dev = torch.device("cuda")
    prob = _problem(2000, 128, 50, torch.float32, device=dev, seed=4)
    a = SAPinn([2, 64, 64, 64, 1], *prob, 5.5e-4, 3.0, seed=7)
    b = SAPinn([2, 64, 64, 64, 1], *prob, 5.5e-4, 3.0, seed=7)
    ra = a.adam_fit(60, record=True, use_graph=True)
    rb = b.adam_fit(60, record=True, use_graph=False)
    torch.cuda.synchronize()
    assert torch.allclose(ra, rb, rtol=1e-4) and torch.allclose(a.theta, b.theta, atol=1e-5)      # same kernels, same order
    assert int(a.step.item()) == 60 and float(ra[-1]) < float(ra[0])
    assert float(a.wf.mean()) > 0.5 - 0.05 and float((a.wf - b.wf).abs().max()) < 1e-4
    f_hist = a.lbfgs_fit(40, 0.8, record=True)       # (the fixed-step iteration is erratic later on: isolated loss spikes of 10-100x in fp64 as well)
    assert len(f_hist) >= 2 and np.isfinite(f_hist).all() and min(f_hist) < f_hist[0]
    w, bb = a.weights_for_P()
    assert [tuple(x.shape) for x in w] == [(64, 2), (64, 64), (64, 64), (1, 64)] and [tuple(x.shape) for x in bb] == [(64,), (64,), (64,), (1,)]
    import AC_models
    P = AC_models.P(w, bb, [2, 64, 64, 64, 1]).to(dev)
    xt = prob[0][:100]
    assert torch.allclose(P(xt)[:, 0], a.predict(xt), atol=1e-5)


@pytest.mark.gpu
def test_tf_free_allen_cahn_driver_writes_the_reference_result_files(tmp_path):
    """AC_main_b200.py (the flow of AC_main.py with gptpinn.sa_pinn in place of TensorFlow), shortened trainings: the greedy
    loop runs, neurons leave the grid, and the result files carry the names / shapes of AC_main.py:438-466."""
    import os
    mat = os.path.join(ROOT, "oracle", "_ref", "Allen-Cahn", "AC.mat")
    if not os.path.exists(mat):
        pytest.skip("AC.mat not available (oracle/_ref is made by oracle/fetch_ref.sh)")
    import AC_main_b200
    out = tmp_path / "ac_data"
    r = AC_main_b200.main(["--mat", mat, "--out", str(out), "--neurons", "3", "--adam", "200", "--lbfgs", "40",
                           "--epochs-gpt", "200", "--epochs-gpt-test", "300"])
    assert np.all(r["loss_list"] > 0) and len({tuple(x) for x in r["neurons"]}) == 3
    shapes = {"generation_time.dat": (3,), "loss_list.dat": (3,), "neurons.dat": (3, 2), "total_time.dat": (), "xt_resid.dat": (20000, 2),
              "ac_test.dat": (25, 2), "xt_test.dat": (3600, 2), "test_gpt_losses.dat": (301, 25), "test_gpt_soln.dat": (3600, 25),
              "test_gpt_time.dat": (25,)}              # AC_main.py:134 test_cases = ceil(0.2 * 121) = 25
    for name, shp in shapes.items():
        a = np.loadtxt(out / name)
        assert a.shape == shp and np.isfinite(a).all(), name
    p = np.load(out / "params.npy", allow_pickle=True).item()
    assert p["parameter size"] == 121 and p["layers_pinn"] == [2, 128, 128, 128, 128, 1]
    assert r["sa_final"][0][0] < 50.0            # 200 Adam + 40 L-BFGS steps already pull the weighted loss far below its start (~1e3)


@pytest.mark.gpu
def test_fused_lbfgs_direction_equals_the_two_loop_recursion_gpu():
    """gptpinn_sa_lbfgs_direction (one launch) against the two-loop recursion of eager_lbfgs.py:80-106 evaluated in fp64 on the
    same fp32 history (the dot-product order differs, so not bitwise): k = 0, a partly filled and a full ring, a ring that
    wraps, the reference-size vector (50 049 weights), 16-byte aligned rows (float4 path; from 8192 weights on the cluster
    kernel, with and without a scalar tail) and unaligned ones (scalar path);
    then a whole L-BFGS run with the kernel and with tensor expressions."""
    from gptpinn import GptPinnError
    from gptpinn.sa_pinn import CudaOps, SAPinn
    dev = torch.device("cuda")
    cu = CudaOps(dev)
    assert cu.lbfgs_capacity() >= 50049
    gen = torch.Generator().manual_seed(3)
    for P, hist, k, start, pad in ((50049, 50, 50, 0, True), (50049, 50, 50, 17, True), (50049, 50, 7, 46, False), (1234, 5, 0, 0, True),
                                   (777, 256, 256, 255, False), (31, 3, 2, 2, True), (4100, 6, 6, 3, True), (4099, 6, 1, 5, True),
                                   (8195, 4, 4, 1, True), (8193, 9, 5, 7, True), (cu.lbfgs_capacity(), 3, 3, 0, True)):
        S = torch.randn(hist, P, generator=gen, dtype=torch.float64) * 1e-2
        Y = S * (0.5 + torch.rand(hist, P, generator=gen, dtype=torch.float64)) + 1e-4 * torch.randn(hist, P, generator=gen, dtype=torch.float64)
        g = torch.randn(P, generator=gen, dtype=torch.float64)
        ld = (P + 3) & ~3 if pad else P
        S32 = torch.zeros(hist, ld, dtype=torch.float32, device=dev)[:, :P]
        Y32 = torch.zeros(hist, ld, dtype=torch.float32, device=dev)[:, :P]
        S32.copy_(S); Y32.copy_(Y)
        g32 = g.to(dev, torch.float32)
        Sd, Yd = S32.double(), Y32.double()
        ro32 = (1.0 / (Sd * Yd).sum(1)).float()
        rows = [(start + i) % hist for i in range(k)]
        Hd = float((Sd[rows[-1]] * Yd[rows[-1]]).sum() / (Yd[rows[-1]] ** 2).sum()) if k else 0.37
        q = -g32.double()
        al = [None] * k
        for i in range(k - 1, -1, -1):
            al[i] = torch.dot(Sd[rows[i]], q) * ro32[rows[i]].double()
            q = q - al[i] * Yd[rows[i]]
        r = q * float(np.float32(Hd))
        for i in range(k):
            r = r + (al[i] - torch.dot(Yd[rows[i]], r) * ro32[rows[i]].double()) * Sd[rows[i]]
        d = torch.full((P,), float("nan"), dtype=torch.float32, device=dev)
        cu.lbfgs_direction(S32, Y32, ro32, k, start, Hd, g32, d)
        d2 = torch.empty_like(d)
        cu.lbfgs_direction(S32, Y32, ro32, k, start, Hd, g32, d2)
        torch.cuda.synchronize()
        assert torch.equal(d, d2)                                                     # fixed summation order
        err = float((d.double() - r).norm() / r.norm())
        assert err < 2e-5, (P, hist, k, start, pad, err)      # fp32 rounding of 2k chained projections
    with pytest.raises(GptPinnError):
        big = torch.zeros(cu.lbfgs_capacity() + 1, dtype=torch.float32, device=dev)
        cu.lbfgs_direction(big[None], big[None], big[:1], 1, 0, 1.0, big, big.clone())
    prob = _problem(1500, 96, 40, torch.float32, device=dev, seed=9)
    a = SAPinn([2, 48, 48, 1], *prob, 5.5e-4, 3.0, seed=2)
    a.adam_fit(100)
    th = a.theta.clone()
    fa = a.lbfgs_fit(25, 0.8, history=6, record=True, use_fused=True)
    a.theta.copy_(th)
    fb = a.lbfgs_fit(25, 0.8, history=6, record=True, use_fused=False)
    n = min(len(fa), len(fb), 8)
    assert n >= 5 and np.allclose(fa[:n], fb[:n], rtol=5e-3)     # fixed-step L-BFGS amplifies rounding; the early iterations agree
