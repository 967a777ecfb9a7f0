#!/usr/bin/env python
"""bench.py -- GPT-PINN offline sweep throughput (param*epochs/s) on 1..8 B200.

Workload (BASELINE.json configs[3], SURVEY.md 8d "config 4"): Klein-Gordon, 15 neurons, synthetic
10^5-parameter grid linspace(-2,-1,50) x linspace(0,1,50) x linspace(0,1,40) in the reference's
meshgrid order, N_R = 100x100 collocation points, 512 IC and 1024 BC points, 2000 epochs of the
reduced-model gradient descent per parameter, lr 0.025, no early exit (fixed work).  One "step" =
one neuron stage = the whole grid swept once (2*10^8 param*epochs) + the exact arg-max.  The
activation matrices come from 15 seeded random [2,40,40,1] cos networks pushed through this
repo's closed-form precompute kernel on the reference point sets.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--epochs E] [--grid a,b,g]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the grid is sharded by (alpha,beta) groups,
no collective on the data path, one 8-byte-per-parameter all-gather for the arg-max per step.
`--impl reference` times the reference's own torch CPU code path (oracle/torch_port.py, the op-for-op
restatement pinned bit-exactly against the reference) on the host cores.
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
for _p in (ROOT, PKG, os.path.join(PKG, "dropin")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_NEURONS = 15
NC, IC_PTS, BC_PTS, N_TEST = 100, 512, 512, 40
LR = 0.025
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12          # 74.5


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


def aux_measurements(device, d):
    """Secondary rows of the metric: precompute seconds at N_R = 10^6 (BASELINE config 5) and the fused
    full-PINN training step (us / Adam epoch at the KG default sizes)."""
    import KG_models
    from gptpinn import pinn, precomp
    out = {}
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
    # config 5, second half: the GPT-PINN test sweep (KG_test.gpt_test: 200 test parameters from the default grid) on
    # N_R = 10^6 residual rows, n = 15; a short run, fixed work per epoch
    from gptpinn import sweep
    g = torch.Generator(device=device).manual_seed(5)
    for b_ in bufs:
        b_.copy_(torch.randn(b_.shape, device=device, generator=g) * 0.3)
    xcos6 = torch.randn((nc * nc, 1), device=device, generator=g) * 0.3
    a_, b2_, g_ = np.linspace(-2, -1, 10), np.linspace(0, 1, 10), np.linspace(0, 1, 10)
    kg_grid = np.array(np.meshgrid(a_, b2_, g_)).T.reshape(-1, 3)
    test_mu = kg_grid[np.random.default_rng(1234).choice(1000, 200, replace=False)]
    c0 = torch.full((N_NEURONS,), 1.0 / N_NEURONS, device=device)
    dense = lambda ep: sweep.kg_sweep(test_mu, c0, bufs[0], bufs[1], bufs[2], d["out_IC"], d["out_IC_t"], d["out_BC"], xcos6,
                                      torch.zeros_like(xcos6), d["IC_u1"], d["IC_u2"], d["BC_u"], ep, 0.001, want_c=False)
    dense(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dense(20)
    e1.record()
    torch.cuda.synchronize()
    out["dense_1e6_test_sweep_param_epochs_per_s"] = 200 * 20 / (e0.elapsed_time(e1) * 1e-3)
    del bufs, xt, xcos6
    net = KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, d["xcos"]).to(device)
    args = (net, "kg", (-1.5, 0.5, 0.5), d["xt_resid"], d["f_hat"], d["IC_xt"], d["IC_u1"], d["IC_u2"], d["BC_xt"], d["BC_u"])
    pinn.train_fused(*args, 10, 5e-4, xcos=d["xcos"])
    torch.cuda.synchronize()
    n_ep = 1000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = pinn.train_fused(*args, n_ep, 5e-4, xcos=d["xcos"])
    e1.record()
    torch.cuda.synchronize()
    out["pinn_train_us_per_epoch"] = 1e3 * e0.elapsed_time(e1) / n_ep
    out["pinn_train_loss_after_1010_epochs"] = float(r["loss"])
    # pinn_test's shape: 8 test parameters per cooperative launch
    mus = [(-1.5 + 0.05 * k, 0.5, 0.5) for k in range(8)]
    nets = [KG_models.NN(np.array([2, 40, 40, 1]), *mu, d["xcos"]).to(device) for mu in mus]
    bargs = (nets, "kg", mus, d["xt_resid"], d["f_hat"], d["IC_xt"], d["IC_u1"], d["IC_u2"], d["BC_xt"], d["BC_u"])
    pinn.train_fused_batch(*bargs, 10, 5e-4, xcos=d["xcos"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pinn.train_fused_batch(*bargs, n_ep // 2, 5e-4, xcos=d["xcos"])
    e1.record()
    torch.cuda.synchronize()
    out["pinn_train_batch8_us_per_epoch_per_training"] = 1e3 * e0.elapsed_time(e1) / (n_ep // 2) / 8
    return out


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


def cpu_reference_rate(mats_cpu, grid, n_params, epochs, threads):
    """Reference torch CPU path (oracle/torch_port.py), fixed work: n_params x epochs updates incl. the loss
    the offline loop evaluates each epoch.  Returns (param_epochs_per_s, seconds)."""
    from oracle import torch_port
    torch.set_num_threads(threads)
    c0 = torch.full((N_NEURONS,), 1.0 / N_NEURONS)
    t0 = time.perf_counter()
    _, _, done = torch_port.kg_offline_generation(grid[:n_params], c0, mats_cpu, epochs, LR, early_exit=False)
    dt = time.perf_counter() - t0
    return done / dt, dt


def c_oracle_rate(mats_cpu, grid, n_params, epochs):
    """The C restatement with OpenMP over parameters (all host threads): a much faster CPU port than the reference."""
    from oracle import oracle
    pr = oracle.kg_problem(N_NEURONS, *[m.numpy() for m in mats_cpu])
    t0 = time.perf_counter()
    oracle.sweep("kg", pr, grid[:n_params], np.full(N_NEURONS, 1.0 / N_NEURONS, np.float32), epochs, LR, 0, "f32")
    dt = time.perf_counter() - t0
    return n_params * epochs / dt, dt


def host_mats(d):
    names = ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat", "IC_u1", "IC_u2", "BC_u")
    return [d[k].detach().cpu().contiguous() for k in names]


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
    workload = f"KG n=15, grid {na}x{nb}x{ng}={len(grid)} params x {args.epochs} epochs, N_R={NC * NC}, N_IC={IC_PTS}, N_BC={2 * BC_PTS}"
    W15 = flops_per_param_epoch(N_NEURONS, NC * NC, IC_PTS, 2 * BC_PTS)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        torch.manual_seed(0)
        # the reference path is size-independent per param*epoch on CPU (dispatch bound); inputs of the named shape
        g = torch.Generator().manual_seed(0)
        mk = lambda r: torch.randn(r, N_NEURONS, generator=g) * 0.3
        mats = [mk(NC * NC), mk(NC * NC), mk(NC * NC), mk(IC_PTS), mk(IC_PTS), mk(2 * BC_PTS),
                torch.randn(NC * NC, 1, generator=g) * 0.3, torch.zeros(NC * NC, 1), torch.randn(IC_PTS, 1, generator=g),
                torch.zeros(IC_PTS, 1), torch.randn(2 * BC_PTS, 1, generator=g)]
        threads = os.cpu_count() or 1
        n_params, ep = 8, 500                                        # ~2-4 s of CPU work per step
        for _ in range(args.warmup):
            cpu_reference_rate(mats, grid, 1, 50, threads)
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            r, dt = cpu_reference_rate(mats, grid, n_params, ep, threads)
            done += n_params * ep
        total = time.perf_counter() - t0
        value = done / total
        sample = f"{n_params} params x {ep} epochs per step (update + loss, no early exit), torch CPU, {threads} threads"
        emit({
            "impl": "reference", "metric": "param_epochs_per_sec", "value": value, "unit": "param*epochs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sampled": sample},
            "cpu_baseline": {"value": value, "unit": "param*epochs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "param*epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return

    # ------------------------------------------------------------------ B200 arm
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

    infos = []

    def step_device():
        """the hot path with inputs resident in HBM: sharded sweep + all-gather + exact arg-max"""
        def run(rows, _c0):
            r = sweep.kg_sweep(rows, c0, d["out"], d["out_xx"], d["out_tt"], d["out_IC"], d["out_IC_t"], d["out_BC"],
                               d["xcos"], d["f_hat"], d["IC_u1"], d["IC_u2"], d["BC_u"], args.epochs, LR, want_c=False, cfg=cfg)
            infos.append(r["info"])
            return r
        res = offline.sharded_sweep(run, grid, group_cols=(0, 1))
        return offline.select(res["f"], res["m"])

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)     # > 126 MB L2

    for _ in range(args.warmup):
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

    # ---- end to end through the drop-in API with HOST buffers (pinned), copies inside the timed region
    names = ("out", "out_xx", "out_tt", "out_IC", "out_IC_t", "out_BC", "xcos", "f_hat", "IC_u1", "IC_u2", "BC_u")
    host = {k: d[k].detach().cpu().pin_memory() for k in names}
    h2d = sum(v.numel() * 4 for v in host.values()) + N_NEURONS * 4
    c0_host = torch.full((N_NEURONS,), 1.0 / N_NEURONS).pin_memory()

    def step_e2e():
        dv = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        cc = c0_host.to(device, non_blocking=True)
        lo, hi = offline.shard_bounds(len(grid))
        Lx, case = KG_train.offline_generation(grid, cc, N, NI, NB, dv["IC_u1"], dv["IC_u2"], dv["BC_u"], dv["xcos"], dv["out"],
                                               dv["out_xx"], dv["out_tt"], dv["out_IC"], dv["out_IC_t"], dv["out_BC"], dv["f_hat"],
                                               args.epochs, LR)
        return float(Lx), case                    # device -> host read of the result

    step_e2e()
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    info = infos[-1]
    launches_per_step = 3                          # pack + sweep + arg-max scan (memsets / copies are not kernels)
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
            roof["traffic"] = json.load(open(prof))["dram_bytes_per_launch"]
        except Exception:
            pass

    cpu = None
    if not args.no_cpu_baseline:
        mats = host_mats(d)
        threads = os.cpu_count() or 1
        n_params, ep = 32, 1000
        v_ref, dt_ref = cpu_reference_rate(mats, grid, n_params, ep, threads)
        cpu = {"value": v_ref, "unit": "param*epochs/s", "cores": threads, "kind": "port",
               "sample": f"{n_params} params x {ep} epochs (update + loss each epoch, no early exit) of the same workload, "
                         f"torch CPU op-for-op port of the reference, {threads} threads, {dt_ref:.1f} s"}
        try:
            v_c, dt_c = c_oracle_rate(mats, grid, 4 * threads, 200)
            cpu["c_oracle_openmp"] = {"value": v_c, "unit": "param*epochs/s", "cores": threads,
                                      "sample": f"{4 * threads} params x 200 epochs, C restatement, OpenMP over parameters, {dt_c:.1f} s"}
        except Exception as e:   # the C oracle is optional test infrastructure
            cpu["c_oracle_openmp"] = {"error": str(e)}

    aux = aux_measurements(device, d) if not args.no_cpu_baseline else {}
    out = {
        "metric": "param_epochs_per_sec", "value": value, "unit": "param*epochs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "parallelism": f"grid sharded over {world} rank(s) by (alpha,beta) group",
                   "l2": "flushed between timed steps (256 MiB write); the 2.3 MB tile stream is L2-resident by design",
                   "mapping": info, "selected_index": int(idx), "largest_loss": float(L)},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "param*epochs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "api": "KG_train.offline_generation (drop-in) with pinned host activation matrices"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks, "precompute_s": {"value": precompute_s, "what": f"one neuron, all KG point sets (N_R={N}), CUDA closed-form kernel"},
        "wall_s_timed_region": t_wall, "ffma_peak_tflops_measured": ffma_peak, "aux": aux,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
