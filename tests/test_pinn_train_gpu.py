"""GPU: the fused full-PINN training kernel vs torch autograd + torch.optim.Adam on the same network
(the reference's NN.loss: KG_models.py:92-129, B_models.py:79-105).  Per-step parity: loss, every
gradient, one Adam update; then a short trajectory.  Tolerances are fp32 (written at each assert)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pack_grads(model):
    return torch.cat([torch.cat([l.weight.grad.reshape(-1), l.bias.grad.reshape(-1)]) for l in model.linears])


def _pack_params(model):
    return torch.cat([torch.cat([l.weight.data.reshape(-1), l.bias.data.reshape(-1)]) for l in model.linears])


def _kg_setup(nc=17, ic=45, bc=33, seed=3):
    import KG_data
    import KG_models
    import KG_precomp
    xt, f_hat, _ = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, nc, 4)
    IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, bc, ic)
    g = torch.Generator().manual_seed(seed)
    f_hat = 0.1 * torch.randn(f_hat.shape, generator=g)              # exercise the targets too
    IC_u2 = 0.1 * torch.randn(IC_u2.shape, generator=g)
    t = [x.cuda() for x in (xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u)]
    xcos = KG_precomp.xcos_term(t[0][:, 0:1], t[0][:, 1:2])
    params = (-1.3, 0.4, 0.7)

    def make():
        net = KG_models.NN(np.array([2, 40, 40, 1]), *params, xcos).cuda()
        gg = torch.Generator().manual_seed(seed + 1)
        for lin in net.linears:                                       # non-zero biases so their gradients are tested
            lin.bias.data = (0.1 * torch.randn(lin.bias.shape, generator=gg)).cuda()
        return net
    return make, params, t, xcos


def _torch_loss_kg(net, t):
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = t
    return net.loss(xt.clone().requires_grad_(), f_hat, IC_xt.clone().requires_grad_(), IC_u1, IC_u2, BC_xt, BC_u)


def test_kg_one_step_matches_autograd_and_adam():
    from gptpinn import pinn
    make, params, t, xcos = _kg_setup()
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = t
    ref = make()
    opt = torch.optim.Adam(ref.parameters(), lr=5e-4)
    loss = _torch_loss_kg(ref, t)
    opt.zero_grad()
    loss.backward()
    g_ref = _pack_grads(ref).clone()
    p0 = _pack_params(ref).clone()
    opt.step()
    p_ref = _pack_params(ref)

    net = make()
    r = pinn.train_fused(net, "kg", params, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, 1, 5e-4, xcos=xcos, want_grad=True)
    torch.cuda.synchronize()
    assert abs(float(r["loss"]) - float(loss)) <= 1e-5 * abs(float(loss))                      # loss: 1e-5 relative
    gerr = float((r["grad"] - g_ref).abs().max() / g_ref.abs().max())
    assert gerr < 2e-4, gerr                                                                   # gradients: 2e-4 of max |g|
    assert pinn.epochs_done(r) == 1
    dp_ref, dp = p_ref - p0, _pack_params(net) - p0
    big = g_ref.abs() > 1e-3 * g_ref.abs().max()                # Adam's first step is lr*sign(g): compare where g is resolved
    assert float((dp - dp_ref)[big].abs().max()) <= 2e-3 * 5e-4
    assert float(dp.abs().max()) <= 1.001 * 5e-4


def test_kg_short_trajectory_and_trace():
    from gptpinn import pinn
    make, params, t, xcos = _kg_setup(nc=13, ic=32, bc=20, seed=11)
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = t
    ref = make()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    ref_losses = []
    for _ in range(25):
        loss = _torch_loss_kg(ref, t)
        ref_losses.append(float(loss))
        opt.zero_grad()
        loss.backward()
        opt.step()
    net = make()
    r = pinn.train_fused(net, "kg", params, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, 25, 1e-3, xcos=xcos, trace_every=1)
    tr = r["trace"].cpu().numpy()
    assert tr.shape == (25,)
    assert np.max(np.abs(tr - np.array(ref_losses)) / np.array(ref_losses)) < 2e-3        # 25 chaotic-free steps: 2e-3 relative
    assert abs(float(r["loss"]) - ref_losses[-1]) <= 2e-3 * ref_losses[-1]
    assert tr[-1] < tr[0]
    perr = float((_pack_params(net) - _pack_params(ref)).abs().max())
    assert perr < 5e-4


def test_burgers_step_tolerance_and_ragged_sizes():
    import B_data
    import B_models
    from gptpinn import pinn
    xt, f_hat, _ = B_data.residual_data(-1.0, 1.0, 0.0, 1.0, 19, 4)       # 361 points: not a multiple of 32
    IC_xt, IC_u, BC_xt, BC_u = B_data.ICBC_data(-1.0, 1.0, 0.0, 1.0, 21, 37)
    t = [x.cuda() for x in (xt, f_hat, IC_xt, IC_u, BC_xt, BC_u)]
    xt, f_hat, IC_xt, IC_u, BC_xt, BC_u = t
    nu = 0.37

    def make():
        return B_models.NN(np.array([2, 20, 20, 20, 20, 1]), nu).cuda()
    ref = make()
    opt = torch.optim.Adam(ref.parameters(), lr=5e-3)
    loss = ref.loss(xt.clone().requires_grad_(), IC_xt, IC_u, BC_xt, BC_u, f_hat)
    opt.zero_grad()
    loss.backward()
    g_ref = _pack_grads(ref).clone()
    net = make()
    p0 = _pack_params(net).clone()
    r = pinn.train_fused(net, "burgers", (nu,), xt, f_hat, IC_xt, IC_u, None, BC_xt, BC_u, 1, 5e-3, want_grad=True)
    assert abs(float(r["loss"]) - float(loss)) <= 1e-5 * abs(float(loss))
    assert float((r["grad"] - g_ref).abs().max() / g_ref.abs().max()) < 2e-4
    assert not torch.equal(_pack_params(net), p0)
    # tolerance above the initial loss: the reference breaks BEFORE the first step (B_train.py:84)
    net2 = make()
    r2 = pinn.train_fused(net2, "burgers", (nu,), xt, f_hat, IC_xt, IC_u, None, BC_xt, BC_u, 50, 5e-3, tol=float(loss) * 2)
    assert pinn.epochs_done(r2) == 1 and torch.equal(_pack_params(net2), p0)
    assert abs(float(r2["loss"]) - float(loss)) <= 1e-5 * abs(float(loss))
    # and a run that converges below a reachable tolerance stops early
    net3 = make()
    r3 = pinn.train_fused(net3, "burgers", (nu,), xt, f_hat, IC_xt, IC_u, None, BC_xt, BC_u, 400, 5e-3, tol=float(loss) * 0.5)
    assert pinn.epochs_done(r3) < 400 and float(r3["loss"]) < float(loss) * 0.5


def test_dropin_pinn_train_prints_and_updates(capsys):
    import KG_train
    make, params, t, xcos = _kg_setup(nc=9, ic=16, bc=16, seed=5)
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = t
    net = make()
    p0 = _pack_params(net).clone()
    KG_train.pinn_train(net, *params, xt.requires_grad_(), IC_xt.requires_grad_(), IC_u1, IC_u2, BC_xt, BC_u, f_hat, 30, 5e-4)
    assert "PINN Final Loss:" in capsys.readouterr().out
    assert not torch.equal(_pack_params(net), p0)


def test_dropin_pinn_test_sweeps_shapes_and_first_records():
    """KG_test.pinn_test / pinn_test_loss and B_test.pinn_test_loss on the fused kernel: shapes, the
    record layout of the reference loops, and the recorded losses against a torch replay."""
    import B_data
    import B_models
    import B_test
    import KG_test
    make, params, t, xcos = _kg_setup(nc=9, ic=16, bc=16, seed=21)
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = t
    kg_test = np.array([[-1.2, 0.3, 0.6], [-1.7, 0.9, 0.1]])
    xt_test = xt[:50].contiguous()
    times, soln = KG_test.pinn_test(kg_test, np.array([2, 40, 40, 1]), xcos, xt, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, f_hat,
                                    40, 5e-4, xt_test)
    assert soln.shape == (50, 2) and times.shape == (2,) and np.isfinite(soln).all()
    losses = KG_test.pinn_test_loss(kg_test, np.array([2, 40, 40, 1]), xcos, xt, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, f_hat,
                                    1500, 5e-4)
    assert losses.shape == (101, 2)
    assert (losses[0] > 0).all() and (losses[1] > 0).all() and (losses[2:] == 0).all()
    assert (losses[1] < losses[0]).all()                       # training made progress; slot 1 = after 1000 (then 1500) steps
    # Burgers: replay with torch and compare the recorded values
    xtb, fb, _ = B_data.residual_data(-1.0, 1.0, 0.0, 1.0, 11, 4)
    ICx, ICu, BCx, BCu = B_data.ICBC_data(-1.0, 1.0, 0.0, 1.0, 12, 12)
    xtb, fb, ICx, ICu, BCx, BCu = (a.cuda() for a in (xtb, fb, ICx, ICu, BCx, BCu))
    layers = np.array([2, 20, 20, 20, 20, 1])
    got = B_test.pinn_test_loss(np.array([0.3]), layers, xtb, ICx, ICu, BCx, BCu, fb, 1000, 2e-3, 1e-9)
    ref = B_models.NN(layers, 0.3).cuda()
    opt = torch.optim.Adam(ref.parameters(), lr=2e-3)
    xr = xtb.clone().requires_grad_()
    want0 = want1000 = None
    for j in range(1, 1001):
        loss = ref.loss(xr, ICx, ICu, BCx, BCu, fb)
        if j == 1:
            want0 = float(loss)
        if j == 1000:
            want1000 = float(loss)
        opt.zero_grad()
        loss.backward()
        opt.step()
    assert got.shape == (61, 1)
    assert abs(got[0, 0] - want0) <= 1e-5 * want0
    # 1000 Adam steps at lr 2e-3 in fp32 are past the horizon over which two correct implementations track each other
    # (test_round2_gpu.py::test_trainer_step_and_trace_match_the_reference_fixture measures that horizon against torch itself):
    # only the level of the loss is comparable here
    assert 0.75 * want1000 <= got[1, 0] <= 1.25 * want1000


def test_pinn_test_and_pinn_test_loss_share_one_training_per_parameter(monkeypatch):
    """The reference's pinn_test and pinn_test_loss train the same networks twice (NN.__init__ reseeds torch).  The drop-in
    trains them once: the second call launches nothing and returns exactly what a re-run computes (checked with the memo
    switched off); an in-place change of any data tensor, other epochs or another order of the calls is a miss / still equal."""
    import B_data
    import B_test
    import KG_test
    from gptpinn import pinn
    make, params, t, xcos = _kg_setup(nc=9, ic=16, bc=16, seed=5)
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = t
    kg_test = np.array([[-1.2, 0.3, 0.6], [-1.7, 0.9, 0.1], [-1.1, 0.2, 0.9]])
    layers = np.array([2, 40, 40, 1])
    xt_test = xt[:40].contiguous()
    calls = []
    orig = pinn.train_fused_batch
    monkeypatch.setattr(pinn, "train_fused_batch", lambda *a, **k: (calls.append(1), orig(*a, **k))[1])
    args = (kg_test, layers, xcos, xt, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, f_hat, 2300, 5e-4)
    pinn.train_memo_clear()
    _, soln = KG_test.pinn_test(*args, xt_test)
    n1 = len(calls)
    losses = KG_test.pinn_test_loss(*args)
    assert n1 == 2 and len(calls) == n1                                   # one chunk: training + final-loss launch; then nothing
    assert (losses[:3] > 0).all() and (losses[3:] == 0).all()
    monkeypatch.setenv("GPTPINN_NO_TRAIN_MEMO", "1")
    losses2 = KG_test.pinn_test_loss(*args)
    _, soln2 = KG_test.pinn_test(*args, xt_test)
    assert len(calls) == n1 + 4 and np.array_equal(losses, losses2) and np.array_equal(soln, soln2)
    monkeypatch.delenv("GPTPINN_NO_TRAIN_MEMO")
    pinn.train_memo_clear()
    losses3 = KG_test.pinn_test_loss(*args)                               # the other order: losses first
    n3 = len(calls)
    _, soln3 = KG_test.pinn_test(*args, xt_test)
    assert len(calls) == n3 and np.array_equal(losses, losses3) and np.array_equal(soln, soln3)
    f_hat.add_(0.0)                                                       # in-place write: version counter moves -> miss
    KG_test.pinn_test_loss(*args)
    assert len(calls) == n3 + 2
    KG_test.pinn_test_loss(*args[:-2], 1200, 5e-4)                        # other epochs -> miss
    assert len(calls) == n3 + 4
    # Burgers (tolerance-stopped)
    xtb, fb, _ = B_data.residual_data(-1.0, 1.0, 0.0, 1.0, 11, 4)
    ICx, ICu, BCx, BCu = B_data.ICBC_data(-1.0, 1.0, 0.0, 1.0, 12, 12)
    xtb, fb, ICx, ICu, BCx, BCu = (a.cuda() for a in (xtb, fb, ICx, ICu, BCx, BCu))
    bl = np.array([2, 20, 20, 20, 20, 1])
    bargs = (np.array([0.3, 0.05]), bl, xtb, ICx, ICu, BCx, BCu, fb, 1500, 2e-3)
    pinn.train_memo_clear()
    n4 = len(calls)
    _, bs = B_test.pinn_test(*bargs, xtb[:20].contiguous(), 1e-9)
    bloss = B_test.pinn_test_loss(*bargs, 1e-9)
    assert len(calls) == n4 + 1
    monkeypatch.setenv("GPTPINN_NO_TRAIN_MEMO", "1")
    assert np.array_equal(bloss, B_test.pinn_test_loss(*bargs, 1e-9))
    assert np.array_equal(bs, B_test.pinn_test(*bargs, xtb[:20].contiguous(), 1e-9)[1])


def test_batched_trainings_match_single_trainings():
    """gptpinn_pinn_train_batch: several trainings in one launch = the same trainings one by one (the grid is cut into
    equal CTA groups, so only the order of the gradient partial sums differs: 1e-4 on a 25-step trajectory); each training
    keeps its own PDE parameters, trace and gradient; tolerance stops are per training."""
    import B_data
    import B_models
    import KG_models
    from gptpinn import pinn
    make, params, t, xcos = _kg_setup(nc=21, ic=40, bc=40)
    xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u = t
    mus = [(-1.3, 0.4, 0.7), (-1.9, 0.1, 0.0), (-1.0, 1.0, 1.0), (-1.5, 0.5, 0.5), (-1.2, 0.9, 0.3)]

    def fresh(mu):
        net = KG_models.NN(np.array([2, 40, 40, 1]), *mu, xcos).cuda()
        gg = torch.Generator().manual_seed(11)
        for lin in net.linears:
            lin.bias.data = (0.1 * torch.randn(lin.bias.shape, generator=gg)).cuda()
        return net
    singles, single_res = [], []
    for mu in mus:
        net = fresh(mu)
        single_res.append(pinn.train_fused(net, "kg", mu, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, 25, 5e-4, xcos=xcos,
                                           trace_every=1, want_grad=True))
        singles.append(_pack_params(net).clone())
    nets = [fresh(mu) for mu in mus]
    res = pinn.train_fused_batch(nets, "kg", mus, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, 25, 5e-4, xcos=xcos,
                                 trace_every=1, want_grad=True)
    for k in range(len(mus)):
        assert float((_pack_params(nets[k]) - singles[k]).abs().max() / singles[k].abs().max()) < 1e-4, k
        assert float((res[k]["trace"] - single_res[k]["trace"]).abs().max() / single_res[k]["trace"].abs().max()) < 1e-4, k
        assert float((res[k]["grad"] - single_res[k]["grad"]).abs().max() / single_res[k]["grad"].abs().max()) < 1e-3, k
        assert pinn.epochs_done(res[k]) == 25
    assert not torch.allclose(res[0]["trace"], res[1]["trace"])              # different PDE parameters, different runs
    with pytest.raises(ValueError):
        pinn.train_fused_batch([fresh(mus[0])] * 9, "kg", [mus[0]] * 9, xt, f_hat, IC_xt, IC_u1, IC_u2, BC_xt, BC_u, 1, 5e-4, xcos=xcos)

    # Burgers, per-training tolerance stop: trainings leave the launch at different epochs
    bxt, bf, _ = B_data.residual_data(-1.0, 1.0, 0.0, 1.0, 19, 4)
    bIC_xt, bIC_u, bBC_xt, bBC_u = B_data.ICBC_data(-1.0, 1.0, 0.0, 1.0, 21, 37)
    bxt, bf, bIC_xt, bIC_u, bBC_xt, bBC_u = (x.cuda() for x in (bxt, bf, bIC_xt, bIC_u, bBC_xt, bBC_u))
    nus = [0.9, 0.05, 0.37]
    mk = lambda nu: B_models.NN(np.array([2, 20, 20, 20, 20, 1]), nu).cuda()
    l0 = [float(pinn.train_fused(mk(nu), "burgers", (nu,), bxt, bf, bIC_xt, bIC_u, None, bBC_xt, bBC_u, 1, 0.0)["loss"]) for nu in nus]
    tol = 0.6 * min(l0)
    one = []
    for nu in nus:
        r = pinn.train_fused(mk(nu), "burgers", (nu,), bxt, bf, bIC_xt, bIC_u, None, bBC_xt, bBC_u, 300, 5e-3, tol=tol)
        one.append((pinn.epochs_done(r), float(r["loss"])))
    bnets = [mk(nu) for nu in nus]
    rb = pinn.train_fused_batch(bnets, "burgers", [(nu,) for nu in nus], bxt, bf, bIC_xt, bIC_u, None, bBC_xt, bBC_u, 300, 5e-3, tol=tol)
    done = [pinn.epochs_done(r) for r in rb]
    assert len(set(done)) > 1                                                  # they do stop at different epochs
    for k in range(len(nus)):
        assert abs(done[k] - one[k][0]) <= 2, (done, one)                      # a loss within 1e-4 of tol may flip by one epoch
        if done[k] < 300:
            assert float(rb[k]["loss"]) < tol
