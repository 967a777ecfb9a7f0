"""GPU: the drop-in modules (same call surface as the reference's KG_/B_/AC_ modules) against the
reference's own outputs for offline_generation / gpt_test / gpt_test_loss / inputs."""
import numpy as np
import pytest
import torch

from conftest import load_golden, lr_of, rel_scalar, selection_ok

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_kg_offline_generation_and_tests():
    import KG_test
    import KG_train
    g = load_golden("kg_small")
    d = {k: _cuda(g[k]) for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "out_test", "xcos",
                                  "f_hat", "IC_u1", "IC_u2", "BC_u")}
    N, NI, NB = d["out"].shape[0], d["out_IC"].shape[0], d["out_BC"].shape[0]
    for i, n in enumerate(g["ns"]):
        n = int(n)
        v = {k: d[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
        L, case = KG_train.offline_generation(g["grid"], torch.full((n,), 1.0 / n, device="cuda"), N, NI, NB,
                                              d["IC_u1"], d["IC_u2"], d["BC_u"], d["xcos"], v["out"], v["out_xx"],
                                              v["out_tt"], v["out_IC"], v["out_IC_t"], v["out_BC"], d["f_hat"],
                                              int(g["epochs"]), lr_of(g, i))
        iref = int(np.nonzero((g["grid"] == g[f"n{n}_case"]).all(1))[0][0])
        idx = int(np.nonzero((g["grid"] == np.asarray(case)).all(1))[0][0])
        assert selection_ok(g[f"n{n}_f"], iref, idx), n              # n = 1: two parameters tie to 1 ulp
        assert all(isinstance(x, np.float64) for x in case)
        assert abs(float(L) - float(g[f"n{n}_L"])) <= TOL * float(L)
        loss_list = np.zeros(1)
        loss_list[0] = L                                            # what KG_main.py:132 does with it
    c0 = torch.full((15,), 1.0 / 15, device="cuda")
    times, soln = KG_test.gpt_test(g["test_mu"], d["out"], d["out_xx"], d["out_tt"], d["out_IC"], d["out_IC_t"],
                                   d["out_BC"], d["f_hat"], d["xcos"], N, NI, NB, d["IC_u1"], d["IC_u2"], d["BC_u"],
                                   c0, int(g["test_epochs"]), lr_of(g, 0), d["out_test"])
    assert soln.shape == g["test_soln"].shape and soln.dtype == np.float64
    assert np.max(np.abs(soln - g["test_soln"])) <= 2e-5 * np.max(np.abs(g["test_soln"]))
    assert times.shape == (len(g["test_mu"]),) and np.all(np.diff(times) > 0)
    losses = KG_test.gpt_test_loss(g["test_mu"], d["out"], d["out_xx"], d["out_tt"], d["out_IC"], d["out_IC_t"],
                                   d["out_BC"], d["f_hat"], d["xcos"], N, NI, NB, d["IC_u1"], d["IC_u2"], d["BC_u"],
                                   c0, int(g["test_epochs"]), lr_of(g, 0))
    nrec = int(g["test_epochs"]) // 500 + 1
    assert losses.shape == (11, len(g["test_mu"]))
    assert rel_scalar(losses[:nrec], g["test_losses"][:nrec]) < TOL


def test_kg_single_step_classes():
    """grad_descent.update and GPT.loss, one call at a time, as user code may drive them."""
    import KG_models
    import KG_optimizer
    import KG_precomp
    g = load_golden("kg_small")
    n = 7
    d = {k: _cuda(g[k]) for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat",
                                  "IC_u1", "IC_u2", "BC_u")}
    v = {k: d[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
    a, b, gm = g["grid"][11]
    T = KG_precomp.Ptt_aPxx_bP(a, b, v["out_tt"], v["out_xx"], v["out"])
    G2 = KG_precomp.gamma2_P(gm, v["out"])
    N, NI, NB = d["out"].shape[0], d["out_IC"].shape[0], d["out_BC"].shape[0]
    GD = KG_optimizer.grad_descent(a, b, gm, N, NI, NB, d["IC_u1"], d["IC_u2"], d["BC_u"], d["xcos"], T, G2,
                                   v["out"], v["out_IC"], v["out_IC_t"], v["out_BC"], 50, 0.025)
    NNg = KG_models.GPT(a, b, gm, v["out"], v["out_IC"], v["out_IC_t"], v["out_BC"], d["xcos"], d["f_hat"], T,
                        d["IC_u1"], d["IC_u2"], d["BC_u"])
    c = torch.full((n,), 1.0 / n, device="cuda")
    l0 = NNg.loss(c.unsqueeze(1))
    assert abs(float(l0) - float(g["n7_trace"][11, 0])) <= TOL * float(l0)
    parts = NNg.lossR(c.unsqueeze(1)) + NNg.lossIC1(c.unsqueeze(1)) + NNg.lossIC2(c.unsqueeze(1)) + NNg.lossBC(c.unsqueeze(1))
    assert abs(float(parts) - float(l0)) <= TOL * float(l0)
    for _ in range(50):
        c = GD.update(c, c.unsqueeze(1))
    assert abs(float(NNg.loss(c.unsqueeze(1))) - float(g["n7_trace"][11, 1])) <= TOL * float(l0)


def test_kg_inputs_fills_columns():
    import KG_models
    import KG_precomp
    g = load_golden("kg_small")
    names = ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "out_test")
    buf = {k: torch.ones(g[k].shape, device="cuda") for k in names}
    xt, IC_xt, BC_xt, xt_test = (_cuda(g[k]) for k in ("xt_resid", "IC_xt", "BC_xt", "xt_test"))
    xt = xt.requires_grad_()
    IC_xt = IC_xt.requires_grad_()
    for i in (0, 1, 2):
        pinn = KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, None).to("cuda")
        for l, lin in enumerate(pinn.linears):
            lin.weight.data = _cuda(g[f"pinn{i}_W{l}"])
            lin.bias.data = _cuda(g[f"pinn{i}_b{l}"])
        views = KG_precomp.inputs(pinn, xt, buf["out"], buf["out_xx"], buf["out_tt"], buf["out_IC_t"], buf["out_IC"],
                                  buf["out_BC"], IC_xt, BC_xt, i, buf["out_test"], xt_test)
        assert [tuple(t.shape) for t in views] == [(g["out"].shape[0], i + 1)] * 3 + [(g["out_IC"].shape[0], i + 1)] * 2 + [(g["out_BC"].shape[0], i + 1)]
        assert views[3].data_ptr() == buf["out_IC_t"].data_ptr()      # order: out, xx, tt, IC_t, IC, BC
    for k in names:
        got, want = buf[k][:, :3].cpu().numpy(), g[k][:, :3]
        assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want)), k
        if k != "out_test":
            assert (buf[k][:, 3:] == 1).all()


def test_burgers_dropin():
    import B_precomp
    import B_models
    import B_test
    import B_train
    g = load_golden("b_small")
    d = {k: _cuda(g[k]) for k in ("out", "out_x", "out_t", "out_xx", "out_IC", "out_BC", "out_test", "f_hat",
                                  "IC_u", "BC_u")}
    N, NI, NB = d["out"].shape[0], d["out_IC"].shape[0], d["out_BC"].shape[0]
    for i, n in enumerate(g["ns"]):
        n = int(n)
        v = {k: d[k][:, :n] for k in ("out", "out_x", "out_t", "out_xx", "out_IC", "out_BC")}
        c0 = B_train.initial_coefficients(g["grid"], g["neurons"][:n], n)
        assert np.array_equal(c0, g[f"n{n}_c0"])
        L, case = B_train.offline_generation(g["grid"], N, NI, NB, d["IC_u"], d["BC_u"], v["out"], v["out_x"],
                                             v["out_t"], v["out_xx"], v["out_IC"], v["out_BC"], d["f_hat"],
                                             int(g["epochs"]), lr_of(g, i), g["neurons"], n - 1)
        assert case == float(g[f"n{n}_case"]) and abs(float(L) - float(g[f"n{n}_L"])) <= TOL * float(L)
    times, soln = B_test.gpt_test(g["test_mu"], N, NI, NB, d["IC_u"], d["BC_u"], d["out"], d["out_x"], d["out_t"],
                                  d["out_xx"], d["out_IC"], d["out_BC"], d["f_hat"], int(g["test_epochs"]), 0.02,
                                  g["neurons"], d["out_test"])
    assert np.max(np.abs(soln - g["test_soln"])) <= 2e-5 * np.max(np.abs(g["test_soln"]))
    losses = B_test.gpt_test_loss(g["test_mu"], N, NI, NB, d["IC_u"], d["BC_u"], d["out"], d["out_x"], d["out_t"],
                                  d["out_xx"], d["out_IC"], d["out_BC"], d["f_hat"], int(g["test_epochs"]), 0.02,
                                  g["neurons"])
    assert rel_scalar(losses[:3], g["test_losses"][:3]) < TOL
    # inputs(): column fill + the 20 % largest-|P_xx| rows zeroed, idx_list recorded
    names = ("out", "out_t", "out_x", "out_xx", "out_IC", "out_BC", "out_test")
    buf = {k: torch.zeros(g[k].shape, device="cuda") for k in names}
    num_mag = g["idx_list"].shape[1]
    idx_list = torch.zeros((9, num_mag), dtype=torch.long)
    xt = _cuda(g["xt_resid"]).requires_grad_()
    for i in (0, 1):
        pinn = B_models.NN(np.array([2, 20, 20, 20, 20, 1]), 0.5).to("cuda")
        for l, lin in enumerate(pinn.linears):
            lin.weight.data = _cuda(g[f"pinn{i}_W{l}"])
            lin.bias.data = _cuda(g[f"pinn{i}_b{l}"])
        B_precomp.inputs(pinn, xt, buf["out"], buf["out_t"], buf["out_x"], buf["out_xx"], buf["out_IC"], buf["out_BC"],
                         _cuda(g["IC_xt"]), _cuda(g["BC_xt"]), i, buf["out_test"], _cuda(g["xt_test"]), N, num_mag,
                         idx_list)
        assert set(idx_list[i].tolist()) == set(g["idx_list"][i].tolist())
    for k in names:
        got, want = buf[k][:, :2].cpu().numpy(), g[k][:, :2]
        assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want)), k


def test_allen_cahn_dropin():
    import AC_test
    import AC_train
    from oracle import oracle
    g = load_golden("ac_small")
    names = ("out", "out_t", "out_xx", "out_IC", "out_BC_ub", "out_BC_lb", "out_BC_diff", "out_BC_ub_x",
             "out_BC_lb_x", "out_BC_diff_x")
    d = {k: _cuda(g[k]) for k in names + ("out_test", "f_hat", "IC_u")}
    N, NI, NB = d["out"].shape[0], d["out_IC"].shape[0], d["out_BC_ub"].shape[0]
    for i, n in enumerate(g["ns"]):
        n = int(n)
        v = {k: d[k][:, :n] for k in names}
        L, case, tc = AC_train.offline_generation(g["grid"], torch.full((n,), 1.0 / n, device="cuda"), N, NI, NB,
                                                  d["IC_u"], v["out"], v["out_t"], v["out_xx"], v["out_IC"],
                                                  v["out_BC_ub"], v["out_BC_lb"], v["out_BC_diff"], v["out_BC_ub_x"],
                                                  v["out_BC_lb_x"], v["out_BC_diff_x"], d["f_hat"], int(g["epochs"]),
                                                  lr_of(g, i))
        _, iref = oracle.argmax_scan(g[f"n{n}_f"], g[f"n{n}_m"])
        idx = int(np.where(np.all(g["grid"] == np.array(case), axis=1))[0][0])
        assert selection_ok(g[f"n{n}_f"], iref, idx)
        if idx == iref:
            assert np.linalg.norm(tc.cpu().numpy() - g[f"n{n}_trained_c"]) <= TOL * np.linalg.norm(g[f"n{n}_trained_c"])
    c0 = torch.full((9,), 1.0 / 9, device="cuda")
    args = (d["out"], d["out_t"], d["out_xx"], d["out_IC"], d["f_hat"], N, NI, NB, d["IC_u"], c0, int(g["test_epochs"]),
            0.0025, d["out_test"], d["out_BC_ub"], d["out_BC_lb"], d["out_BC_diff"], d["out_BC_ub_x"], d["out_BC_lb_x"],
            d["out_BC_diff_x"])
    times, soln = AC_test.gpt_test(g["test_mu"], *args)
    assert np.max(np.abs(soln - g["test_soln"])) <= 2e-5 * np.max(np.abs(g["test_soln"]))
    losses = AC_test.gpt_test_loss(g["test_mu"], *args)
    assert losses.shape == g["test_losses"].shape == (int(g["test_epochs"]) + 1, len(g["test_mu"]))
    assert rel_scalar(losses, g["test_losses"]) < TOL


def test_kg_greedy_outer_loop_selects_the_reference_neuron_sequence():
    """KG_main.py:100-138 with fixed synthetic PINN weights per stage (golden from the reference's own
    inputs() + offline_generation): the drop-in inputs() + offline_generation must hand over the identical
    neuron sequence, stage by stage, with the grid shrinking exactly as the reference's does."""
    import KG_models
    import KG_precomp
    import KG_train
    g = load_golden("kg_greedy")
    s = load_golden("kg_small")
    nmax = int(g["nmax"])
    xt, IC_xt, BC_xt, xt_test = (_cuda(s[k]) for k in ("xt_resid", "IC_xt", "BC_xt", "xt_test"))
    d = {k: _cuda(s[k]) for k in ("xcos", "f_hat", "IC_u1", "IC_u2", "BC_u")}
    N, NI, NB, NT = xt.shape[0], IC_xt.shape[0], BC_xt.shape[0], xt_test.shape[0]
    buf = {k: torch.ones((r, nmax), device="cuda") for k, r in
           (("out", N), ("out_xx", N), ("out_tt", N), ("out_IC", NI), ("out_IC_t", NI), ("out_BC", NB))}
    out_test = torch.zeros((NT, nmax), device="cuda")
    alpha, beta, gamma = np.linspace(-2, -1, 4), np.linspace(0, 1, 3), np.linspace(0, 1, 5)
    kg_train = np.array(np.meshgrid(alpha, beta, gamma)).T.reshape(-1, 3)
    neurons = np.zeros((nmax, 3))
    neurons[0] = (np.median(alpha), np.median(beta), np.median(gamma))
    loss_list = np.zeros(nmax)
    for i, neuron in enumerate(neurons):
        kg_train = np.delete(kg_train, np.where(np.all(kg_train == neuron, axis=1))[0], axis=0)
        assert len(kg_train) == int(g["sizes"][i])
        pinn = KG_models.NN(np.array([2, 40, 40, 1]), *neuron, d["xcos"]).to("cuda")
        for l, lin in enumerate(pinn.linears):                         # stage i's "trained" network = fixed weights
            lin.weight.data = _cuda(s[f"pinn{i}_W{l}"])
            lin.bias.data = _cuda(s[f"pinn{i}_b{l}"])
        v = KG_precomp.inputs(pinn, xt, buf["out"], buf["out_xx"], buf["out_tt"], buf["out_IC_t"], buf["out_IC"],
                              buf["out_BC"], IC_xt, BC_xt, i, out_test, xt_test)
        train_out, train_out_xx, train_out_tt, train_out_IC_t, train_out_IC, train_out_BC = v
        c_initial = torch.full((1, i + 1), 1 / (i + 1)).to("cuda")[0]
        L, case = KG_train.offline_generation(kg_train, c_initial, N, NI, NB, d["IC_u1"], d["IC_u2"], d["BC_u"], d["xcos"],
                                              train_out, train_out_xx, train_out_tt, train_out_IC, train_out_IC_t,
                                              train_out_BC, d["f_hat"], int(g["epochs"]), float(g["lr"]))
        loss_list[i] = L
        if i + 1 < nmax:
            neurons[i + 1] = case
    assert np.array_equal(neurons, g["neurons"])
    assert np.max(np.abs(loss_list - g["loss_list"]) / g["loss_list"]) < 2e-5
