#!/usr/bin/env python
"""bench.py -- GPT-PINN offline sweep throughput (param*epochs/s) on 1..8 B200.

Workload (BASELINE.json configs[3], SURVEY.md 8d "config 4"): Klein-Gordon, synthetic 10^5-parameter
grid linspace(-2,-1,50) x linspace(0,1,50) x linspace(0,1,40) in the reference's meshgrid order,
N_R = 100x100 collocation points, 512 IC and 1024 BC points, 2000 epochs of the reduced-model
gradient descent per parameter, lr 0.025.  The activation matrices come from 15 seeded random
[2,40,40,1] cos networks pushed through the precompute on the reference point sets.

  step      = the 15-neuron stage: the whole grid swept once with no early exit (2*10^8 param*epochs,
              fixed work) + the exact arg-max.  `value` = param*epochs/s with inputs resident in HBM,
              `e2e` = the same through the drop-in KG_train.offline_generation with HOST buffers.
  full      = all 15 greedy stages n = 1..15 through KG_train.offline_generation (grid shrinking by the
              chosen parameter as KG_main.py:100-138 does): per-stage ms and roofline fraction, the
              sum-of-stages roofline, and the same loop once more with the reference's early-exit
              semantics as saved work (identical selections required) -> wall time to the same neurons.
  verified  = the 2000-epoch selection of the step is re-derived from the fp64 CPU oracle on the
              top-64 + 64 random parameters (checker only; never on the product path).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--epochs E] [--grid a,b,g]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the grid is sharded by (alpha,beta) groups,
no collective on the data path, two 8-byte all-reduces for the arg-max per step.
`--impl reference` times the UNMODIFIED reference (oracle/_ref, copied by oracle/fetch_ref.sh) on the
host cores: its own inputs(), gpt_test (fixed work) and offline_generation (with its early exit) on a
bounded sample of the same workload; with a GPU present also the same modules with device = cuda.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gpt-pinn_b200")
REF_KG = os.path.join(ROOT, "oracle", "_ref", "Klein-Gordon")

N_NEURONS = 15
NC, IC_PTS, BC_PTS, N_TEST = 100, 512, 512, 40
LR = 0.025
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12          # 74.5
INPUTS = "15 seeded synthetic [2,40,40,1] cos PINNs (Xavier-normal x0.6 + N(0,0.05^2), zero bias) on the reference point sets"


def flops_per_param_epoch(n, NR, NI, NB):
    """Algorithmic flop of one reference update: 8n*N_R + 4n*(N_BC + 2*N_IC)."""
    return 8 * n * NR + 4 * n * (NB + 2 * NI)


def kg_grid(na, nb, ng):
    a, b, g = np.linspace(-2, -1, na), np.linspace(0, 1, nb), np.linspace(0, 1, ng)
    return np.array(np.meshgrid(a, b, g)).T.reshape(-1, 3)       # KG_main.py:46-49 ordering


def synthetic_weights(seed):
    """Seeded [2,40,40,1] weights: Xavier-normal x 0.6 + N(0, 0.05^2), zero biases."""
    g = torch.Generator().manual_seed(seed)
    layers = [2, 40, 40, 1]
    out = []
    for fi, fo in zip(layers[:-1], layers[1:]):
        std = (2.0 / (fi + fo)) ** 0.5
        w = torch.randn(fo, fi, generator=g) * std * 0.6 + torch.randn(fo, fi, generator=g) * 0.05
        out.append((w.float(), torch.zeros(fo)))
    return out


def workload_name(na, nb, ng, epochs):
    return (f"KG n=15, grid {na}x{nb}x{ng}={na * nb * ng} params x {epochs} epochs, N_R={NC * NC}, N_IC={IC_PTS}, "
            f"N_BC={2 * BC_PTS}")


# =====================================================================================================================
# B200 arm helpers
# =====================================================================================================================
def build_problem(device, nc=NC):
    """Point sets (reference KG_data semantics) + the seven [N, 15] activation buffers, filled column by
    column by the CUDA precompute kernel.  Returns dict of device tensors and the precompute time."""
    import KG_data
    import KG_precomp
    from gptpinn import precomp
    xt_resid, f_hat, xt_test = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, nc, N_TEST)
    IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, BC_PTS, IC_PTS)
    d = dict(xt_resid=xt_resid, f_hat=f_hat, xt_test=xt_test, IC_xt=IC_xt, IC_u1=IC_u1, IC_u2=IC_u2, BC_xt=BC_xt, BC_u=BC_u)
    d = {k: v.to(device) for k, v in d.items()}
    d["xcos"] = KG_precomp.xcos_term(d["xt_resid"][:, 0:1], d["xt_resid"][:, 1:2])
    N, NI, NB, NT = xt_resid.shape[0], IC_xt.shape[0], BC_xt.shape[0], xt_test.shape[0]
    for k, r in (("out", N), ("out_xx", N), ("out_tt", N), ("out_IC", NI), ("out_IC_t", NI), ("out_BC", NB), ("out_test", NT)):
        d[k] = torch.ones((r, N_NEURONS), device=device)
    t_pre = []
    for i in range(N_NEURONS):
        layers = [(w.to(device), b.to(device)) for w, b in synthetic_weights(1000 + i)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        precomp.mlp_channels(layers, "cos", d["xt_resid"], v=d["out"][:, i], q=d["out_xx"][:, i], r=d["out_tt"][:, i])
        precomp.mlp_channels(layers, "cos", d["IC_xt"], v=d["out_IC"][:, i], t=d["out_IC_t"][:, i])
        precomp.mlp_channels(layers, "cos", d["BC_xt"], v=d["out_BC"][:, i])
        precomp.mlp_channels(layers, "cos", d["xt_test"], v=d["out_test"][:, i])
        e1.record()
        torch.cuda.synchronize()
        t_pre.append(e0.elapsed_time(e1) * 1e-3)
    return d, float(np.median(t_pre))


def _timed(fn, reps=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps, r


def aux_measurements(device, d, ffma_peak):
    """Secondary rows: the other BASELINE configs as short fixed-work sweeps with their own roofline fractions
    (configs 1, 2, 3 = default-size grids of KG / Burgers / Allen-Cahn; config 5 = dense KG: precompute at
    N_R = 10^6 + the gpt_test sweep), and the fused full-PINN trainer."""
    import KG_models
    from gptpinn import pinn, precomp, sweep
    out = {}
    g = torch.Generator(device=device).manual_seed(5)
    mk = lambda r, n: torch.randn(r, n, device=device, generator=g) * 0.3
    vec = lambda r: torch.randn(r, 1, device=device, generator=g) * 0.3
    zer = lambda r: torch.zeros(r, 1, device=device)
    E = 2000
    # ---- config 1: KG default grid (10^3 parameters), n = 15, the real activation matrices of this run
    grid1 = kg_grid(10, 10, 10)
    c15 = torch.full((N_NEURONS,), 1.0 / N_NEURONS, device=device)
    run1 = lambda: sweep.kg_sweep(grid1, c15, d["out"], d["out_xx"], d["out_tt"], d["out_IC"], d["out_IC_t"], d["out_BC"], d["xcos"],
                                  d["f_hat"], d["IC_u1"], d["IC_u2"], d["BC_u"], E, LR, want_c=False)
    run1()
    dt, r = _timed(run1, 3)
    W = flops_per_param_epoch(15, NC * NC, IC_PTS, 2 * BC_PTS)
    out["config1_kg_default_grid_n15"] = {"params": 1000, "epochs": E, "ms_per_stage": 1e3 * dt, "param_epochs_per_s": 1000 * E / dt,
                                          "roofline_frac": W * 1000 * E / dt / 1e12 / ffma_peak, "mapping": r["info"]}
    # ---- config 2: Burgers default grid (128 nu), n = 9
    NR, NI, NB, n = 10000, 200, 400, 9
    mats = [mk(NR, n) for _ in range(4)] + [mk(NI, n), mk(NB, n)]
    gridb = np.linspace(0.005, 1, 129)[:128]
    cb = torch.full((128, n), 1.0 / n, device=device)
    fh, iu, bu = zer(NR), vec(NI), zer(NB)
    run2 = lambda: sweep.b_sweep(gridb, cb, *mats, fh, iu, bu, E, 0.001, want_c=False)
    run2()
    dt, r = _timed(run2, 3)
    W = 12 * n * NR + 4 * n * (NB + NI)
    out["config2_burgers_default_grid_n9"] = {"params": 128, "epochs": E, "ms_per_stage": 1e3 * dt, "param_epochs_per_s": 128 * E / dt,
                                              "roofline_frac": W * 128 * E / dt / 1e12 / ffma_peak, "mapping": r["info"]}
    # ---- config 3: Allen-Cahn default grid (120 (lambda, eps)), n = 9
    NR, NI, NB = 20000, 512, 100
    mats = [mk(NR, n) for _ in range(3)] + [mk(NI, n)] + [mk(NB, n) for _ in range(6)]
    lm, ep = np.linspace(1e-3, 1e-4, 11), np.linspace(1, 5, 11)
    grida = np.array(np.meshgrid(lm, ep)).T.reshape(-1, 2)[:120]
    ca = torch.full((n,), 1.0 / n, device=device)
    fh, iu = zer(NR), vec(NI)
    run3 = lambda: sweep.ac_sweep(grida, ca, *mats, fh, iu, E, 0.0001, want_c=False)
    run3()
    dt, r = _timed(run3, 3)
    W = 8 * n * NR + 4 * n * NI + 12 * n * NB
    out["config3_allen_cahn_default_grid_n9"] = {"params": 120, "epochs": E, "ms_per_stage": 1e3 * dt, "param_epochs_per_s": 120 * E / dt,
                                                 "roofline_frac": W * 120 * E / dt / 1e12 / ffma_peak, "mapping": r["info"]}
    del mats
    # ---- config 5: dense KG, N_R = 10^6: precompute of one neuron + the gpt_test sweep (200 test parameters, n = 15)
    nc = 1000
    x = torch.linspace(-1.0, 1.0, nc, device=device)
    t = torch.linspace(0.0, 5.0, nc, device=device)
    xt = torch.stack((x.repeat(nc), t.repeat_interleave(nc)), dim=1).contiguous()
    layers = [(w.to(device), b.to(device)) for w, b in synthetic_weights(1000)]
    bufs = [torch.empty((nc * nc, N_NEURONS), device=device) for _ in range(3)]
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        precomp.mlp_channels(layers, "cos", xt, v=bufs[0][:, 3], q=bufs[1][:, 3], r=bufs[2][:, 3])
        precomp.mlp_channels(layers, "cos", d["IC_xt"], v=d["out_IC"][:, 0], t=d["out_IC_t"][:, 0])
        precomp.mlp_channels(layers, "cos", d["BC_xt"], v=d["out_BC"][:, 0])
        precomp.mlp_channels(layers, "cos", d["xt_test"], v=d["out_test"][:, 0])
        e1.record()
        torch.cuda.synchronize()
        out["precompute_1e6_s"] = e0.elapsed_time(e1) * 1e-3
    for b_ in bufs:
        b_.copy_(torch.randn(b_.shape, device=device, generator=g) * 0.3)
    xcos6 = torch.randn((nc * nc, 1), device=device, generator=g) * 0.3
    test_mu = grid1[np.random.default_rng(1234).choice(1000, 200, replace=False)]
    z6 = torch.zeros_like(xcos6)
    dense = lambda ep: sweep.kg_sweep(test_mu, c15, bufs[0], bufs[1], bufs[2], d["out_IC"], d["out_IC_t"], d["out_BC"], xcos6,
                                      z6, d["IC_u1"], d["IC_u2"], d["BC_u"], ep, 0.001, want_c=False)
    dense(2)
    shots = [_timed(lambda: dense(20)) for _ in range(5)]          # single 20-epoch shots spread by +-15 % from run to run: median of five
    dt, r = sorted(shots, key=lambda x: x[0])[2]
    W = flops_per_param_epoch(15, nc * nc, IC_PTS, 2 * BC_PTS)
    out["config5_dense_1e6_test_sweep"] = {"params": 200, "epochs": 20, "shots_ms": [round(1e3 * x[0], 2) for x in shots], "param_epochs_per_s": 200 * 20 / dt,
                                           "roofline_frac": W * 200 * 20 / dt / 1e12 / ffma_peak, "mapping": r["info"]}
    out["dense_1e6_test_sweep_param_epochs_per_s"] = 200 * 20 / dt
    del bufs, xt, xcos6, z6
    # ---- fused full-PINN trainer
    net = KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, d["xcos"]).to(device)
    args = (net, "kg", (-1.5, 0.5, 0.5), d["xt_resid"], d["f_hat"], d["IC_xt"], d["IC_u1"], d["IC_u2"], d["BC_xt"], d["BC_u"])
    pinn.train_fused(*args, 10, 5e-4, xcos=d["xcos"])
    n_ep = 1000
    dt, r = _timed(lambda: pinn.train_fused(*args, n_ep, 5e-4, xcos=d["xcos"]))
    out["pinn_train_us_per_epoch"] = 1e6 * dt / n_ep
    out["pinn_train_loss_after_1010_epochs"] = float(r["loss"])
    mus = [(-1.5 + 0.05 * k, 0.5, 0.5) for k in range(8)]
    nets = [KG_models.NN(np.array([2, 40, 40, 1]), *mu, d["xcos"]).to(device) for mu in mus]
    bargs = (nets, "kg", mus, d["xt_resid"], d["f_hat"], d["IC_xt"], d["IC_u1"], d["IC_u2"], d["BC_xt"], d["BC_u"])
    pinn.train_fused_batch(*bargs, 10, 5e-4, xcos=d["xcos"])
    dt, _ = _timed(lambda: pinn.train_fused_batch(*bargs, n_ep // 2, 5e-4, xcos=d["xcos"]))
    out["pinn_train_batch8_us_per_epoch_per_training"] = 1e6 * dt / (n_ep // 2) / 8
    # ---- SA-PINN trainer (config 3's full-order solver, AC_main.py:204-339) at the reference sizes; single process only
    try:
        import torch.distributed as _dist
        if not (_dist.is_available() and _dist.is_initialized()):
            out.update(_sa_pinn_timings(device))
    except Exception as e:                             # a secondary row must never hide the headline
        out["sa_pinn_error"] = repr(e)
    return out


def _sa_pinn_timings(device):
    """ms per Adam iteration (CUDA graph replay) and per L-BFGS iteration of gptpinn.sa_pinn.SAPinn: [2,128,128,128,128,1],
    20 000 collocation + 512 initial + 2 x 100 boundary points (AC_main.py:49-51, 125)."""
    from gptpinn.sa_pinn import SAPinn
    g = torch.Generator().manual_seed(0)
    NRs, NIs, NBs = 20000, 512, 100
    xt_f = torch.stack((2 * torch.rand(NRs, generator=g) - 1, torch.rand(NRs, generator=g)), 1).to(device)
    x0 = torch.linspace(-1, 1, NIs)
    xt0 = torch.stack((x0, torch.zeros(NIs)), 1).to(device)
    u0 = (x0 ** 2 * torch.cos(np.pi * x0)).to(device)
    tb = torch.rand(NBs, generator=g)
    xt_lb, xt_ub = torch.stack((-torch.ones(NBs), tb), 1).to(device), torch.stack((torch.ones(NBs), tb), 1).to(device)
    m = SAPinn([2, 128, 128, 128, 128, 1], xt_f, xt0, u0, xt_lb, xt_ub, 1e-4, 5.0, seed=1)
    m.adam_fit(20)
    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    m.adam_fit(200)
    e1.record()
    torch.cuda.synchronize()
    res["sa_pinn_adam_ms_per_iter"] = e0.elapsed_time(e1) / 200
    m.lbfgs_fit(10, 0.8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fh = m.lbfgs_fit(100, 0.8)
    torch.cuda.synchronize()
    res["sa_pinn_lbfgs_ms_per_iter"] = 1e3 * (time.perf_counter() - t0) / max(len(fh) - 1, 1)
    res["sa_pinn_loss_after_320_iterations"] = float(fh[-1])
    return res


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": float(np.median(pw)) if pw else None, "samples": len(sm), "reasons": reasons}


def host_mats(d):
    names = ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat", "IC_u1", "IC_u2", "BC_u")
    return [d[k].detach().cpu().contiguous() for k in names]


def verify_selection(d, grid, f, m, L, idx, epochs, n_top=64, n_rand=64):
    """Checker leg (CPU oracle, fp64 adjudicator): re-derive the step's answer on the parameters that decide it.
    Top-`n_top` by final loss + `n_rand` random rows are re-run in the fp64 C oracle; their losses must agree to 1e-5
    and the exact scan over the GPU arrays with those entries replaced by the oracle's must return the same index."""
    from oracle import oracle
    fh, mh = f.detach().cpu().numpy().copy(), m.detach().cpu().numpy().copy()
    order = np.argsort(-np.where(np.isnan(fh), -np.inf, fh), kind="stable")
    rng = np.random.default_rng(7)
    rows = np.unique(np.concatenate([order[:n_top], rng.choice(len(grid), min(n_rand, len(grid)), replace=False), [int(idx)] if idx >= 0 else []]).astype(np.int64))
    pr = oracle.kg_problem(N_NEURONS, *[t.numpy() for t in host_mats(d)])
    t0 = time.perf_counter()
    o = oracle.sweep("kg", pr, grid[rows], np.full(N_NEURONS, 1.0 / N_NEURONS, np.float32), epochs, LR, 0, "f64")
    dt = time.perf_counter() - t0
    ef = float(np.max(np.abs(fh[rows] - o["f"]) / np.abs(o["f"])))
    em = float(np.max(np.abs(mh[rows] - o["m"]) / np.abs(o["m"])))
    f2, m2 = fh.copy(), mh.copy()
    f2[rows], m2[rows] = o["f"], o["m"]
    Lo, io = oracle.argmax_scan(f2, m2)
    Lg, ig = oracle.argmax_scan(fh, mh)
    ok = bool(ef <= 1e-5 and em <= 1e-5 and io == int(idx) and ig == int(idx) and abs(Lo - float(L)) <= 1e-5 * abs(Lo))
    return {"selection_verified": ok, "rows_checked": int(len(rows)), "max_rel_err_final_loss": ef, "max_rel_err_min_prefix_loss": em,
            "oracle": "fp64 C restatement (oracle/), OpenMP", "oracle_seconds": dt, "oracle_index": int(io), "gpu_index": int(idx),
            "oracle_largest_loss": float(Lo), "gpu_largest_loss": float(L)}


def full_offline(d, grid, epochs, ffma_peak, world, early_exit):
    """All 15 greedy stages through the drop-in KG_train.offline_generation (grid shrinking by the chosen parameter,
    KG_main.py:100-138); fixed synthetic activation columns.  Returns per-stage records."""
    import KG_train
    from gptpinn import offline
    N, NI, NB = d["out"].shape[0], d["out_IC"].shape[0], d["out_BC"].shape[0]
    g = grid.copy()
    stages, cases = [], []
    offline.EARLY_EXIT = bool(early_exit)
    try:
        t_all = time.perf_counter()
        for i in range(N_NEURONS):
            n = i + 1
            c0 = torch.full((n,), 1.0 / n, device=d["out"].device)
            v = {k: d[k][:, :n] for k in ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC")}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L, case = KG_train.offline_generation(g, c0, N, NI, NB, d["IC_u1"], d["IC_u2"], d["BC_u"], d["xcos"], v["out"], v["out_xx"],
                                                  v["out_tt"], v["out_IC"], v["out_IC_t"], v["out_BC"], d["f_hat"], epochs, LR)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            pe = float(len(g)) * epochs
            W = flops_per_param_epoch(n, N, NI, NB)
            rec = {"n": n, "params": int(len(g)), "ms": ms, "largest_loss": float(L), "case": [float(x) for x in case]}
            if early_exit:
                rec["stats"] = {k: v_ for k, v_ in offline.last_stats.items() if isinstance(v_, (int, float, str))}
            else:
                rec["param_epochs_per_s"] = pe / (ms * 1e-3)
                rec["frac"] = W * pe / (ms * 1e-3) / 1e12 / world / ffma_peak
            stages.append(rec)
            cases.append(tuple(case))
            g = np.delete(g, np.where(np.all(g == np.asarray(case), axis=1))[0], axis=0)      # KG_main.py:102
        wall = time.perf_counter() - t_all
    finally:
        offline.EARLY_EXIT = False
    return stages, cases, wall


# =====================================================================================================================
# reference arm: the UNMODIFIED reference on the host cores (and on the GPU through torch when there is one)
# =====================================================================================================================
def reference_arm(args, emit, workload, grid):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not os.path.isdir(REF_KG):
        emit({"impl": "reference", "unavailable": "oracle/_ref missing: run `sh oracle/fetch_ref.sh` where /root/reference exists"})
        return
    sys.path.insert(0, REF_KG)
    import KG_data, KG_models, KG_precomp, KG_test, KG_train        # noqa: E401  (the reference's own modules)
    assert os.path.dirname(os.path.abspath(KG_train.__file__)) == REF_KG
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)

    def problem(device):
        """the same synthetic workload as the B200 arm, built with the reference's own inputs() (autograd)"""
        for mod in (KG_models, KG_precomp, KG_test, KG_train):
            mod.device = device
        xt_resid, f_hat, xt_test = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, NC, N_TEST)
        IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, BC_PTS, IC_PTS)
        t = dict(xt_resid=xt_resid, f_hat=f_hat, xt_test=xt_test, IC_xt=IC_xt, IC_u1=IC_u1, IC_u2=IC_u2, BC_xt=BC_xt, BC_u=BC_u)
        t = {k: v.to(device) for k, v in t.items()}
        t["xcos"] = KG_precomp.xcos_term(t["xt_resid"][:, 0].unsqueeze(1), t["xt_resid"][:, 1].unsqueeze(1))
        t["xt_resid"].requires_grad_()
        t["IC_xt"].requires_grad_()
        N, NI, NB, NT = xt_resid.shape[0], IC_xt.shape[0], BC_xt.shape[0], xt_test.shape[0]
        for k, r in (("out", N), ("out_xx", N), ("out_tt", N), ("out_IC", NI), ("out_IC_t", NI), ("out_BC", NB)):
            t[k] = torch.ones((r, N_NEURONS)).to(device)
        t["out_test"] = torch.zeros((NT, N_NEURONS)).to(device)
        t0 = time.perf_counter()
        for i in range(N_NEURONS):
            net = KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, t["xcos"]).to(device)
            for lin, (w, b) in zip(net.linears, synthetic_weights(1000 + i)):
                lin.weight.data = w.to(device)
                lin.bias.data = b.to(device)
            KG_precomp.inputs(net, t["xt_resid"], t["out"], t["out_xx"], t["out_tt"], t["out_IC_t"], t["out_IC"], t["out_BC"],
                              t["IC_xt"], t["BC_xt"], i, t["out_test"], t["xt_test"])
        if device.type == "cuda":
            torch.cuda.synchronize()
        t["precompute_s_per_neuron"] = (time.perf_counter() - t0) / N_NEURONS
        t["sizes"] = (N, NI, NB)
        return t

    def fixed_work(t, rows, ep):
        """reference KG_test.gpt_test: `ep` updates for each row, no early exit (KG_test.py:10-44)"""
        N, NI, NB = t["sizes"]
        c0 = torch.full((1, N_NEURONS), 1 / N_NEURONS).to(t["out"].device)[0]
        t0 = time.perf_counter()
        KG_test.gpt_test(grid[rows], t["out"], t["out_xx"], t["out_tt"], t["out_IC"], t["out_IC_t"], t["out_BC"], t["f_hat"], t["xcos"],
                         N, NI, NB, t["IC_u1"], t["IC_u2"], t["BC_u"], c0, ep, LR, t["out_test"])
        return time.perf_counter() - t0

    cpu = torch.device("cpu")
    tc = problem(cpu)
    n_params, ep, n_sub = (int(x) for x in args.ref_sample.split(","))     # bounded sample: a few seconds of CPU work per step
    rows = np.linspace(0, len(grid) - 1, n_params).astype(np.int64)
    for _ in range(args.warmup):
        fixed_work(tc, rows[:1], 50)
    total = 0.0
    for _ in range(max(args.steps, 1)):
        total += fixed_work(tc, rows, ep)
    steps = max(args.steps, 1)
    value = n_params * ep * steps / total
    sample = (f"{n_params} evenly spaced grid rows x {ep} epochs per step through the reference's own KG_test.gpt_test (update only, "
              f"no early exit) on the same synthetic matrices (built by the reference's inputs()), torch CPU, {threads} threads")
    extra = {"precompute_s_per_neuron_reference_cpu": tc["precompute_s_per_neuron"]}
    # the reference's offline_generation itself (with its early exit) on a subsample of the grid: wall time per greedy stage
    sub = np.linspace(0, len(grid) - 1, n_sub).astype(np.int64)
    N, NI, NB = tc["sizes"]
    c0 = torch.full((1, N_NEURONS), 1 / N_NEURONS)[0]
    t0 = time.perf_counter()
    Lr, case_r = KG_train.offline_generation(grid[sub], c0, N, NI, NB, tc["IC_u1"], tc["IC_u2"], tc["BC_u"], tc["xcos"], tc["out"], tc["out_xx"],
                                             tc["out_tt"], tc["out_IC"], tc["out_IC_t"], tc["out_BC"], tc["f_hat"], args.epochs, LR)
    dt = time.perf_counter() - t0
    extra["offline_generation_early_exit"] = {
        "rows": int(len(sub)), "epochs_cap": args.epochs, "seconds": dt, "seconds_per_row": dt / len(sub),
        "extrapolated_seconds_per_stage_full_grid": dt / len(sub) * len(grid), "largest_loss": float(Lr),
        "what": f"unmodified KG_train.offline_generation (breaks when loss < running max) on {n_sub} evenly spaced rows, n = 15"}
    if torch.cuda.is_available():                                  # BASELINE.md 3, second row: same modules, device = cuda
        try:
            tg = problem(torch.device("cuda"))
            fixed_work(tg, rows[:1], 20)
            torch.cuda.synchronize()
            dtg = fixed_work(tg, rows[:4], 250)
            torch.cuda.synchronize()
            extra["reference_on_b200"] = {"value": 4 * 250 / dtg, "unit": "param*epochs/s", "sample": "4 rows x 250 epochs, reference gpt_test, device=cuda",
                                          "precompute_s_per_neuron": tg["precompute_s_per_neuron"]}
        except Exception as e:                                     # the CPU number is the arm; the GPU row is a bonus
            extra["reference_on_b200"] = {"error": repr(e)}
    emit({
        "impl": "reference", "metric": "param_epochs_per_sec", "value": value, "unit": "param*epochs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "inputs": INPUTS},
        "cpu_baseline": {"value": value, "unit": "param*epochs/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "param*epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "reference": extra})


# =====================================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--epochs", type=int, default=2000)
    ap.add_argument("--grid", default="50,50,40")
    ap.add_argument("--cfg", default="", help="M:SL:W:mode[:K] override of the kernel mapping (tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the 15-stage runs and the aux configs")
    ap.add_argument("--ref-sample", default="16,500,24", help="reference arm: rows,epochs of the fixed-work sample per step, rows of the early-exit sample")
    args = ap.parse_args()

    # stdout carries exactly ONE JSON line: anything libraries print meanwhile (e.g. NCCL's version banner at communicator
    # creation) is sent to stderr; the real stdout is restored for the result line only
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    na, nb, ng = (int(x) for x in args.grid.split(","))
    grid = kg_grid(na, nb, ng)
    workload = workload_name(na, nb, ng, args.epochs)
    W15 = flops_per_param_epoch(N_NEURONS, NC * NC, IC_PTS, 2 * BC_PTS)

    if args.impl == "reference":
        reference_arm(args, emit, workload, grid)
        return

    # ------------------------------------------------------------------ B200 arm
    for _p in (ROOT, PKG, os.path.join(PKG, "dropin")):
        if _p not in sys.path:
            sys.path.insert(0, _p)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import KG_train
    from gptpinn import _lib, offline, sweep

    d, precompute_s = build_problem(device)
    N, NI, NB = d["out"].shape[0], d["out_IC"].shape[0], d["out_BC"].shape[0]
    c0 = torch.full((N_NEURONS,), 1.0 / N_NEURONS, device=device)
    cfg = None
    if args.cfg:
        v_ = [int(x) for x in args.cfg.split(":")] + [0] * 5
        cfg = dict(M=v_[0], SL=v_[1], W=v_[2], mode=v_[3], K=v_[4])

    tf = C.c_double()
    _lib.check(_lib.lib().gptpinn_ffma_peak(C.byref(tf), 3, None), "gptpinn_ffma_peak")
    ffma_peak = tf.value

    infos, paths = [], []

    def run15(rows, _c0, epochs=args.epochs, **kw):
        r = sweep.kg_sweep(rows, c0, d["out"], d["out_xx"], d["out_tt"], d["out_IC"], d["out_IC_t"], d["out_BC"],
                           d["xcos"], d["f_hat"], d["IC_u1"], d["IC_u2"], d["BC_u"], epochs, LR, want_c=False, cfg=cfg, **kw)
        infos.append(r["info"])
        return r

    def step_device():
        """the hot path with inputs resident in HBM: sharded sweep + exact arg-max (packed-key all-reduces when sharded)"""
        L_, idx_, _ = offline.greedy_select(run15, grid, args.epochs, group_cols=(0, 1), early_exit=False)
        paths.append(offline.last_stats.get("path"))
        return L_, idx_

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)     # > 126 MB L2

    # first warm-up step keeps the full (f, m) arrays for the selection check
    full_res = offline.sharded_sweep(run15, grid, group_cols=(0, 1))
    L0, idx0 = offline.select(full_res["f"], full_res["m"])
    for _ in range(max(args.warmup - 1, 0)):
        L, idx = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(float(k))                     # evict L2 between timed steps (outside the event pair)
        ev[k][0].record()
        L, idx = step_device()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    total_pe = float(len(grid)) * args.epochs * args.steps
    value = total_pe / (ms * 1e-3)
    same_answer = (int(idx) == int(idx0)) and abs(float(L) - float(L0)) <= 1e-6 * abs(float(L0))

    # ---- end to end through the drop-in API with HOST buffers (pinned), copies inside the timed region
    names = ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat", "IC_u1", "IC_u2", "BC_u")
    host = {k: d[k].detach().cpu().pin_memory() for k in names}
    h2d = sum(v.numel() * 4 for v in host.values()) + N_NEURONS * 4
    c0_host = torch.full((N_NEURONS,), 1.0 / N_NEURONS).pin_memory()

    def step_e2e():
        dv = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        cc = c0_host.to(device, non_blocking=True)
        Lx, case = KG_train.offline_generation(grid, cc, N, NI, NB, dv["IC_u1"], dv["IC_u2"], dv["BC_u"], dv["xcos"], dv["out"],
                                               dv["out_xx"], dv["out_tt"], dv["out_IC"], dv["out_IC_t"], dv["out_BC"], dv["f_hat"],
                                               args.epochs, LR)
        return float(Lx), case                    # device -> host read of the result

    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        Le, case_e = step_e2e()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(len(grid)) * args.epochs * e2e_steps / float(te.item())

    # ---- all 15 greedy stages, fixed work and with early-exit semantics (every rank takes part; rank 0 reports)
    full = None
    if not args.no_full:
        barrier()
        st_fixed, cases_fixed, wall_fixed = full_offline(d, grid, args.epochs, ffma_peak, world, early_exit=False)
        barrier()
        st_exit, cases_exit, wall_exit = full_offline(d, grid, args.epochs, ffma_peak, world, early_exit=True)
        barrier()
        sumW = sum(flops_per_param_epoch(s["n"], N, NI, NB) * s["params"] * float(args.epochs) for s in st_fixed)
        t_fixed = sum(s["ms"] for s in st_fixed) * 1e-3
        full = {"stages": st_fixed, "full_offline_s": t_fixed, "full_offline_wall_s": wall_fixed,
                "param_epochs": float(sum(s["params"] for s in st_fixed)) * args.epochs,
                "param_epochs_per_s": float(sum(s["params"] for s in st_fixed)) * args.epochs / t_fixed,
                "sum_of_stages_frac": sumW / t_fixed / 1e12 / world / ffma_peak,
                "sum_of_stages_frac_of_nominal_74.5": sumW / t_fixed / 1e12 / world / FP32_NOMINAL_TFLOPS,
                "early_exit": {"full_offline_s": sum(s["ms"] for s in st_exit) * 1e-3, "full_offline_wall_s": wall_exit,
                               "identical_selection": cases_exit == cases_fixed,
                               "stages": [{"n": s["n"], "ms": s["ms"], **s.get("stats", {})} for s in st_exit],
                               "what": "same 15 stages with the reference's early-exit semantics as saved work (scout pass, seed groups, "
                                       "in-kernel exit below a bound <= the reference's running maximum); selections must equal the fixed-work run"}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    info = infos[-1]
    launches_per_step = 3 + (1 if info.get("kernels_launched", 2) >= 3 else 0)   # pack + [T arrays] + sweep + arg-max scan (memsets / copies / all-reduces are not our kernels)
    sweep_ms = ms / args.steps                     # the sweep kernel is > 99.9 % of a step (profiles/)
    achieved = W15 * value / 1e12 / world          # per-GPU algorithmic TFLOP/s
    roof = {"bound": "fp32", "achieved": achieved, "peak": ffma_peak, "unit": "TFLOP/s", "frac": achieved / ffma_peak,
            "peak_source": "FFMA microbenchmark measured in this run (gptpinn_ffma_peak); MEASURED_PEAKS.json has no FP32 figure",
            "frac_of_nominal_74.5": achieved / FP32_NOMINAL_TFLOPS,
            "flop_per_param_epoch": W15, "kernel": "gp::sweep_kernel<KG,15,M,W,PRIV>", "kernel_ms_per_launch": sweep_ms,
            "traffic": None}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        alg_bytes = (3 * N + NB + 2 * NI) * N_NEURONS * 4 + (len(grid) // world) * 8
        roof["hbm"] = {"algorithmic_bytes_per_launch": alg_bytes, "achieved_GBps": alg_bytes / (sweep_ms * 1e-3) / 1e9,
                       "peak_GBps": peaks["hbm_gbs"], "frac": alg_bytes / (sweep_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    except Exception:
        pass
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            roof["traffic"] = tj["dram_bytes_per_pass"] * (args.epochs + 1) if "dram_bytes_per_pass" in tj else tj["dram_bytes_per_launch"]
            roof["traffic_note"] = ("DRAM bytes per launch = ncu bytes per pass (profiles/traffic.json) x (epochs + 1) passes: the precomputed T arrays "
                                    "(1.6 GB at 10^5 parameters) are streamed from HBM once per epoch by design -- idle HBM bandwidth traded for FP32 issue slots")
        except Exception:
            pass

    verified = None
    try:
        verified = verify_selection(d, grid, full_res["f"], full_res["m"], L0, idx0, args.epochs)
        verified["timed_steps_returned_the_same_answer"] = bool(same_answer)
        verified["selection_verified"] = bool(verified["selection_verified"] and same_answer)
    except Exception as e:                        # the oracle is test infrastructure; its absence must not hide the measurement
        verified = {"selection_verified": False, "error": repr(e)}

    cpu = None
    if not args.no_cpu_baseline:
        # the unmodified reference on this box's host cores, in its own process (its module names collide with the drop-in's)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1",
                                "--epochs", str(args.epochs), "--grid", args.grid],
                               capture_output=True, text=True, timeout=900, env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            ref = json.loads(p.stdout.strip().splitlines()[-1])
            cpu = dict(ref.get("cpu_baseline") or {"error": ref.get("unavailable")})
            cpu["reference"] = ref.get("reference")
        except Exception as e:
            cpu = {"error": repr(e)}

    aux = aux_measurements(device, d, ffma_peak) if not args.no_full else {}
    out = {
        "metric": "param_epochs_per_sec", "value": value, "unit": "param*epochs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "inputs": INPUTS},
        "run": {"parallelism": f"grid sharded over {world} rank(s) by (alpha,beta) group; selection path {paths[-1]}",
                "l2": "flushed between timed steps (256 MiB write); the 2.3 MB tile stream is L2-resident by design",
                "mapping": info, "selected_index": int(idx), "largest_loss": float(L)},
        "roofline": roof, "cpu_baseline": cpu, "selection": verified,
        "e2e": {"value": e2e_value, "unit": "param*epochs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "api": "KG_train.offline_generation (drop-in) with pinned host activation matrices"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks, "precompute_s": {"value": precompute_s, "what": f"one neuron, all KG point sets (N_R={N}), CUDA closed-form kernel"},
        "wall_s_timed_region": t_wall, "ffma_peak_tflops_measured": ffma_peak, "full_offline": full, "aux": aux,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
