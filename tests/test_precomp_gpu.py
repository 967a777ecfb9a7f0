"""GPU parity: closed-form MLP channel kernel vs the reference's autograd inputs() (golden) and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
PRE_TOL = 1e-5          # of the column's max-abs; fp64 adjudication in test_round2_gpu.py


def _col_err(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _layers(g, i):
    out, l = [], 0
    while f"pinn{i}_W{l}" in g:
        out.append((torch.from_numpy(g[f"pinn{i}_W{l}"]).cuda(), torch.from_numpy(g[f"pinn{i}_b{l}"]).cuda()))
        l += 1
    return out


def _eval(layers, act, xt, names):
    from gptpinn import precomp
    xt = torch.from_numpy(np.ascontiguousarray(xt)).cuda()
    # strided destinations: column 3 of a [N, 7] buffer, as inputs() writes column i of [N, neurons]
    bufs = {k: torch.full((xt.shape[0], 7), -7.0, device="cuda") for k in names}
    precomp.mlp_channels(layers, act, xt, **{k: bufs[k][:, 3] for k in names})
    torch.cuda.synchronize()
    for k in names:
        assert (bufs[k][:, [0, 1, 2, 4, 5, 6]] == -7.0).all()      # neighbours untouched
    return {k: bufs[k][:, 3].cpu().numpy() for k in names}


def test_kg_channels():
    g = load_golden("kg_small")
    for i in (0, 7, 14):
        L = _layers(g, i)
        ch = _eval(L, "cos", g["xt_resid"], "vxtqr")
        assert _col_err(ch["v"], g["out"][:, i]) < PRE_TOL
        assert _col_err(ch["q"], g["out_xx"][:, i]) < PRE_TOL
        assert _col_err(ch["r"], g["out_tt"][:, i]) < PRE_TOL
        ic = _eval(L, "cos", g["IC_xt"], "vt")
        assert _col_err(ic["v"], g["out_IC"][:, i]) < PRE_TOL
        assert _col_err(ic["t"], g["out_IC_t"][:, i]) < PRE_TOL
        assert _col_err(_eval(L, "cos", g["BC_xt"], "v")["v"], g["out_BC"][:, i]) < PRE_TOL


def test_burgers_and_ac_channels_vs_oracle_and_golden():
    from oracle import oracle
    g = load_golden("b_small")
    L = _layers(g, 2)
    ch = _eval(L, "tanh", g["xt_resid"], "vxtq")
    o = oracle.mlp_channels([2, 20, 20, 20, 20, 1], [w.cpu().numpy() for w, _ in L], [b.cpu().numpy() for _, b in L],
                            "tanh", g["xt_resid"])
    for k in "vxtq":
        assert _col_err(ch[k], o[k]) < PRE_TOL
    g = load_golden("ac_small")
    L = _layers(g, 1)
    ch = _eval(L, "tanh", g["xt_resid"], "vtq")
    assert _col_err(ch["v"], g["out"][:, 1]) < PRE_TOL
    assert _col_err(ch["t"], g["out_t"][:, 1]) < PRE_TOL
    assert _col_err(ch["q"], g["out_xx"][:, 1]) < PRE_TOL
    ub = _eval(L, "tanh", g["BC_xt_ub"], "vx")
    assert _col_err(ub["x"], g["out_BC_ub_x"][:, 1]) < PRE_TOL


def test_edge_sizes():
    g = load_golden("kg_small")
    L = _layers(g, 0)
    for N in (1, 31, 129):
        ch = _eval(L, "cos", g["xt_resid"][:N], "vq")
        assert _col_err(ch["v"], g["out"][:N, 0]) < PRE_TOL


def test_dense_million_points_config5():
    """BASELINE config 5 size (N_R = 10^6): spot-check 4096 random points against the oracle, check that the
    result does not depend on the batch it was computed in (per-point determinism)."""
    from oracle import oracle
    from gptpinn import precomp
    g = load_golden("kg_small")
    L = _layers(g, 3)
    nc = 1000
    x = torch.linspace(-1.0, 1.0, nc, device="cuda")
    t = torch.linspace(0.0, 5.0, nc, device="cuda")
    xt = torch.stack((x.repeat(nc), t.repeat_interleave(nc)), dim=1).contiguous()
    bufs = {k: torch.empty((nc * nc, 15), device="cuda") for k in "vqr"}
    precomp.mlp_channels(L, "cos", xt, **{k: bufs[k][:, 7] for k in "vqr"})
    torch.cuda.synchronize()
    idx = torch.from_numpy(np.random.default_rng(0).choice(nc * nc, 4096, replace=False)).cuda()
    o = oracle.mlp_channels([2, 40, 40, 1], [w.cpu().numpy() for w, _ in L], [b.cpu().numpy() for _, b in L], "cos",
                            xt[idx].cpu().numpy())
    for k in "vqr":
        assert _col_err(bufs[k][idx, 7].cpu().numpy(), o[k]) < PRE_TOL
    sub = {k: torch.empty(4096, device="cuda") for k in "vqr"}
    precomp.mlp_channels(L, "cos", xt[idx].contiguous(), **sub)
    for k in "vqr":
        assert torch.equal(sub[k], bufs[k][idx, 7])
