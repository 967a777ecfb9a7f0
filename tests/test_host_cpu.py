"""CPU-only: C-ABI surface, host-side logic (grid sharding over gloo, Burgers initial coefficients,
data modules) -- nothing here launches a kernel."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

LIB = os.path.join(ROOT, "gpt-pinn_b200", "gptpinn", "lib", "libgptpinn_b200.so")


def _ensure_lib():
    if not os.path.exists(LIB):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return LIB


def test_c_abi_exports_every_declared_symbol():
    lib = ctypes.CDLL(_ensure_lib())
    hdr = open(os.path.join(ROOT, "include", "gptpinn_b200.h")).read()
    declared = set(re.findall(r"\b(gptpinn_[a-z0-9_]+)\s*\(", hdr))
    assert {"gptpinn_kg_sweep", "gptpinn_b_sweep", "gptpinn_ac_sweep", "gptpinn_argmax_scan",
            "gptpinn_mlp_channels", "gptpinn_create", "gptpinn_destroy", "gptpinn_last_error"} <= declared
    for name in declared:
        assert hasattr(lib, name), name
    lib.gptpinn_version.restype = ctypes.c_int
    assert lib.gptpinn_version() >= 100
    from gptpinn import _lib
    assert set(_lib.EXPORTS) <= declared


def test_sass_is_sm100_with_tma_bulk_copy():
    out = subprocess.run(["cuobjdump", "-lelf", _ensure_lib()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN2gp12sweep_kernelILi0ELi15ELi5ELi8ELb1EEEvNS_11SweepParamsE",
                           _ensure_lib()], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass and "FFMA" in sass


def test_no_cpu_fallback():
    from gptpinn import GptPinnError, sweep
    t = torch.zeros(4, 2)
    with pytest.raises(GptPinnError):
        sweep.kg_sweep(np.zeros((1, 3)), torch.zeros(2), t, t, t, t, t, t, torch.zeros(4), None, torch.zeros(4),
                       None, torch.zeros(4), 1, 0.1)


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "gpt-pinn_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("SURVEY", ""), os.path.join(dp, f)


def test_burgers_initial_coefficients_match_reference():
    import B_train
    g = load_golden("b_small")
    for n in (1, 2, 5, 9):
        assert np.array_equal(B_train.initial_coefficients(g["grid"], g["neurons"][:n], n), g[f"n{n}_c0"])
    # a grid value that IS a neuron gets a one-hot (B_train.py:28-31)
    c0 = B_train.initial_coefficients(np.array([g["neurons"][1]]), g["neurons"][:3], 3)
    assert c0.tolist() == [[0.0, 1.0, 0.0]]


def test_data_modules_match_reference_point_sets():
    import B_data
    import KG_data
    g = load_golden("kg_small")
    xt_resid, f_hat, xt_test = KG_data.residual_data(-1.0, 1.0, 0.0, 5.0, 24, 10)
    IC_xt, IC_u1, IC_u2, BC_xt, BC_u = KG_data.ICBC_data(-1.0, 1.0, 0.0, 5.0, 64, 64)
    for got, key in ((xt_resid, "xt_resid"), (xt_test, "xt_test"), (IC_xt, "IC_xt"), (BC_xt, "BC_xt"),
                     (IC_u1, "IC_u1"), (IC_u2, "IC_u2"), (BC_u, "BC_u"), (f_hat, "f_hat")):
        assert np.array_equal(got.numpy(), g[key]), key
    g = load_golden("b_small")
    xt_resid, f_hat, xt_test = B_data.residual_data(-1.0, 1.0, 0.0, 1.0, 24, 10)
    IC_xt, IC_u, BC_xt, BC_u = B_data.ICBC_data(-1.0, 1.0, 0.0, 1.0, 40, 40)
    for got, key in ((xt_resid, "xt_resid"), (xt_test, "xt_test"), (IC_xt, "IC_xt"), (BC_xt, "BC_xt"),
                     (IC_u, "IC_u"), (BC_u, "BC_u")):
        assert np.array_equal(got.numpy(), g[key]), key


def test_kg_precomp_helpers_match_reference_formulas():
    import KG_precomp
    g = load_golden("kg_small")
    xt = torch.from_numpy(g["xt_resid"])
    assert np.array_equal(KG_precomp.xcos_term(xt[:, 0:1], xt[:, 1:2]).numpy(), g["xcos"])
    out, oxx, ott = (torch.from_numpy(g[k][:, :3]) for k in ("out", "out_xx", "out_tt"))
    T = KG_precomp.Ptt_aPxx_bP(-1.25, 0.5, ott, oxx, out)
    assert torch.equal(T, (ott + (-1.25) * oxx) + 0.5 * out)
    assert torch.equal(KG_precomp.gamma2_P(0.3, out), (2 * 0.3) * out)


def test_shard_bounds_cover_grid():
    from gptpinn import offline
    for n in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            spans = [offline.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gptpinn import offline
    from oracle import oracle
    rng = np.random.default_rng(0)
    Np, n = 37, 4
    grid = rng.random((Np, 3))
    f_all = rng.random(Np).astype(np.float32)
    m_all = (f_all * rng.choice([0.5, 1.0], Np)).astype(np.float32)
    c_all = rng.random((Np, n)).astype(np.float32)
    seen = {}

    grid[:, 0] = rng.integers(0, 3, Np)          # few distinct (alpha, beta): sibling groups
    grid[:, 1] = rng.integers(0, 2, Np)
    c0_all = torch.from_numpy(rng.random((Np, n)).astype(np.float32))

    def run(rows, c0_rows):          # test double for the CUDA sweep: looks the rows up in the known answer
        ids = [int(np.nonzero(np.all(grid == r, axis=1))[0][0]) for r in rows]
        seen["c0"] = None if c0_rows is None else tuple(c0_rows.shape)
        seen["c0_ok"] = c0_rows is None or torch.equal(c0_rows, c0_all[ids])
        seen["groups"] = {(r[0], r[1]) for r in rows}
        return dict(f=torch.from_numpy(f_all[ids]), m=torch.from_numpy(m_all[ids]),
                    c=torch.from_numpy(c_all[ids]), info={})

    res = offline.sharded_sweep(run, grid, c0=c0_all, want_c=True, gather_c=True, group_cols=(0, 1))
    ok0 = seen["c0_ok"] and len(seen["groups"]) <= 4          # whole sibling groups stay on one rank (6 groups / 2)
    ok = ok0 and np.array_equal(res["f"].numpy(), f_all) and np.array_equal(res["m"].numpy(), m_all)
    ok = ok and np.array_equal(res["c"].numpy(), c_all)
    lo, hi = offline.shard_bounds(Np)
    ok = ok and seen["c0"] == (hi - lo, n)
    L, idx = oracle.argmax_scan(res["f"].numpy(), res["m"].numpy())      # every rank scans the same full arrays
    row = offline.winner_row(res, idx, Np)
    ok = ok and np.array_equal(row.numpy(), c_all[idx])
    q.put((rank, bool(ok), idx))
    dist.destroy_process_group()


def test_grid_sharding_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in out)
    assert len({idx for _, _, idx in out}) == 1


def test_planner_assigns_every_parameter_once():
    """Host-only planner (no GPU): whatever the grid shape, mapping override or team size, every parameter lands in
    exactly one work item, teams divide the CTA, and the configuration fits one SM."""
    from gptpinn import _lib
    rng = np.random.default_rng(3)
    a = np.linspace(-2, -1, 50); b = np.linspace(0, 1, 50); gm = np.linspace(0, 1, 40)
    big = np.array(np.meshgrid(a, b, gm)).T.reshape(-1, 3)
    order = np.lexsort((big[:, 1], big[:, 0]))
    grids = {"config4": big, "shard_1_of_8": big[order[12500:25000]], "default": big[::100],
             "ragged": np.column_stack([rng.integers(0, 7, 333) * 0.1, rng.integers(0, 3, 333) * 0.5, rng.random(333)]),
             "one_group": np.column_stack([np.full(500, -1.5), np.full(500, 0.5), rng.random(500)]),
             "all_distinct": rng.random((257, 3)), "single": np.array([[-1.5, 0.5, 0.5]])}
    for name, grid in grids.items():
        for cfg in (None, dict(K=1), dict(mode=2, K=2), dict(mode=2, K=8), dict(mode=1), dict(M=4), dict(M=1, mode=2, K=4)):
            for n in (1, 15):
                try:
                    d = _lib.kg_plan(grid, n, cfg=cfg)
                except _lib.GptPinnError:
                    assert cfg and cfg.get("K", 0) > 1, (name, cfg, n)      # a forced team size may be infeasible, nothing else
                    continue
                assert d["covered"] == len(grid), (name, cfg, n)
                assert d["W"] % d["K"] == 0 and 1 <= d["ctas"] <= 148, (name, cfg, n, d)
                assert sum(d["sl_hist"].values()) == d["items"]
                if cfg and "K" in cfg and d["priv"]:
                    assert d["K"] == cfg["K"]
    shard = _lib.kg_plan(grids["shard_1_of_8"], 15)
    assert shard["priv"] == 1 and shard["K"] == 8          # 313 sibling groups for 1184 warps: whole CTAs team up


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm): exactly one stdout line, valid JSON,
    the keys of the bench contract, and the same metric / unit / workload naming as the B200 arm uses."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--epochs", "200", "--ref-sample", "4,100,6"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "param_epochs_per_sec" and d["unit"] == "param*epochs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert d["cpu_baseline"]["kind"] == "reference"          # the unmodified reference (oracle/_ref), not a port
    assert d["reference"]["offline_generation_early_exit"]["rows"] == 6
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_batched_trainer_rejects_bad_batches_without_a_gpu():
    """Host-side checks of train_fused_batch: batch size limits, one parameter tuple per model, no CPU fallback."""
    import KG_models
    from gptpinn import _lib, pinn
    xcos = torch.zeros(4, 1)
    nets = [KG_models.NN(np.array([2, 40, 40, 1]), -1.5, 0.5, 0.5, xcos) for _ in range(2)]
    xt = torch.zeros(4, 2)
    one = torch.zeros(4, 1)
    args = (xt, one, xt, one, one, xt, one, 1, 5e-4)
    with pytest.raises(ValueError):
        pinn.train_fused_batch(nets, "kg", [(-1.5, 0.5, 0.5)], *args, xcos=xcos)              # 2 models, 1 parameter tuple
    with pytest.raises(ValueError):
        pinn.train_fused_batch([], "kg", [], *args, xcos=xcos)
    with pytest.raises(ValueError):
        pinn.train_fused_batch(nets * 5, "kg", [(-1.5, 0.5, 0.5)] * 10, *args, xcos=xcos)     # more than MAX_BATCH
    with pytest.raises(_lib.GptPinnError):
        pinn.train_fused_batch(nets, "kg", [(-1.5, 0.5, 0.5)] * 2, *args, xcos=xcos)          # CPU tensors: no fallback


def test_header_is_plain_c_and_the_c_demo_links():
    """include/gptpinn_b200.h is a C header (no C++, no torch types): examples/c_abi_demo.c compiles with gcc and links
    against the shared library.  (It needs a GPU to run; on a box without one gptpinn_create fails with an error code.)"""
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    out = os.path.join(ROOT, "build", "c_abi_demo")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    libdir = os.path.join(ROOT, "gpt-pinn_b200", "gptpinn", "lib")
    cmd = ["/usr/bin/gcc", "-O1", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", out, "-L", libdir, "-lgptpinn_b200",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not torch.cuda.is_available():
        run = subprocess.run([out], capture_output=True, text=True)
        assert run.returncode != 0 and "gptpinn_create" in run.stderr          # fails loudly, no fallback


# ---------------------------------------------------------------------------------------------------------------
# selection logic added in round 2: packed-key fast path (two 8-byte all-reduces) and the early-exit scheme
# ---------------------------------------------------------------------------------------------------------------
def _cpu_scan(f, m):
    """stand-in for the CUDA scan kernel in host-logic tests: the oracle's literal N1 scan"""
    from oracle import oracle
    L, idx = oracle.argmax_scan(f.numpy(), m.numpy())
    return torch.tensor(L, dtype=torch.float32), idx


def _literal_reference_loop(traj):
    """KG_train.py:12-43 on recorded loss trajectories traj[p, j] (loss after j updates)."""
    L, case = 0.0, -1
    E = traj.shape[1] - 1
    for p in range(traj.shape[0]):
        loss = traj[p, 0]
        for i in range(1, E + 1):
            if loss < L:
                break
            loss = traj[p, i]
        if loss > L:
            L, case = float(loss), p
    return L, case


def _trajectories(rng, Np, E, kind):
    base = rng.random(Np).astype(np.float32) * 0.9 + 0.05
    j = np.arange(E + 1, dtype=np.float32)
    rate = (rng.random(Np).astype(np.float32) * 0.2 + 0.01)[:, None]
    traj = base[:, None] * (1.0 + 3.0 * np.exp(-rate * j[None, :]))
    if kind == "nonmono":
        wob = 1.0 + 0.6 * np.sin(j[None, :] * rng.random((Np, 1)).astype(np.float32) * 0.7) * (rng.random((Np, 1)) < 0.5)
        traj = traj * wob.astype(np.float32)
        traj[rng.choice(Np, 3, replace=False), E // 2:] = np.nan
    return traj.astype(np.float32)


def _fake_run(grid, traj):
    def run(rows, _c0, epochs=None, bound=None):
        ids = np.array([int(np.nonzero(np.all(grid == r, axis=1))[0][0]) for r in rows], dtype=np.int64)
        f = np.empty(len(ids), np.float32)
        m = np.empty(len(ids), np.float32)
        b = None if bound is None else bound.numpy()
        for k, p in enumerate(ids):
            t = traj[p, :epochs + 1]
            f[k], mm = t[epochs], np.inf
            for jj in range(epochs):
                if b is not None and t[jj] < b[k]:
                    f[k] = mm = t[jj]
                    break
                if t[jj] < mm:
                    mm = t[jj]
            m[k] = mm
        return dict(f=torch.from_numpy(f), m=torch.from_numpy(m), c=None, info={})
    return run


@pytest.mark.parametrize("kind", ["mono", "nonmono"])
def test_early_exit_select_equals_the_literal_reference_loop(kind, monkeypatch):
    from gptpinn import offline
    monkeypatch.setattr(offline, "select", _cpu_scan)
    rng = np.random.default_rng(11 if kind == "mono" else 12)
    a, b, g = np.linspace(-2, -1, 6), np.linspace(0, 1, 5), np.linspace(0, 1, 8)
    grid = np.array(np.meshgrid(a, b, g)).T.reshape(-1, 3)
    E = 128
    traj = _trajectories(rng, len(grid), E, kind)
    want = _literal_reference_loop(traj)
    for group_cols in ((0, 1), None):
        L, idx, stats = offline.early_exit_select(_fake_run(grid, traj), grid, E, group_cols=group_cols, scout_epochs=4,
                                                  seed_frac=0.05)
        assert idx == want[1] and float(L) == np.float32(want[0]), (kind, group_cols, stats)
        assert stats["seeds"] + stats["proven_by_scout"] + stats["bounded"] == len(grid)
        if kind == "mono":
            assert stats["bounded"] - stats.get("ran_to_the_end", 0) + stats["proven_by_scout"] > len(grid) // 2   # work was saved


def test_exit_bounds_are_exclusive_prefix_maxima():
    from gptpinn import offline
    f = torch.tensor([0.5, float("nan"), 0.9, 0.2])
    m = torch.tensor([0.7, 0.1, 0.8, 0.3])
    rows = torch.tensor([1, 3, 6, 8])
    b = offline.exit_bounds(f, m, rows, 10, 0.0).numpy()
    assert list(b[:2]) == [-1.0, -1.0]                     # nothing finished before rows 0, 1
    assert np.allclose(b[2:7], 0.5) and np.allclose(b[7:], 0.8)   # NaN row is no source; min(f, m) of row 6 = 0.8


def _key_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from gptpinn import offline
    offline.select = _cpu_scan
    assert offline._dist() is not None and dist.is_initialized()          # lazily created from the launcher's environment
    rng = np.random.default_rng(5)
    out = []
    for case in range(6):
        Np = 41 + case
        grid = rng.random((Np, 3))
        grid[:, 0] = rng.integers(0, 3, Np); grid[:, 1] = rng.integers(0, 2, Np)
        f = rng.random(Np).astype(np.float32)
        m = (f * rng.choice([0.3, 1.0, 1.5], Np)).astype(np.float32)
        if case == 1:
            f[rng.choice(Np, 4, replace=False)] = np.nan
        if case == 2:
            f[:] = np.float32(0.25)                                       # exact ties: first index wins
        if case == 3:
            w = int(np.nanargmax(f)); m[w] = 0.0                          # the arg-max dipped early: guard fails -> gather
        if case == 4:
            f[:] = 0.0                                                    # nothing exceeds 0

        def run(rows, _c0, epochs=None, key_index=None):
            ids = [int(np.nonzero(np.all(grid == r, axis=1))[0][0]) for r in rows]
            assert key_index is not None and key_index.tolist() == ids      # the kernel epilogue gets the GLOBAL row indices
            key = None
            if case % 2 == 0:                                                # emulate the fused epilogue on half of the cases
                ff = torch.from_numpy(f[ids])
                key = offline.pack_keys(ff, key_index.to(torch.int64)).max().reshape(1) if len(ids) else torch.zeros(1, dtype=torch.int64)
            return dict(f=torch.from_numpy(f[ids]), m=torch.from_numpy(m[ids]), c=None, key=key, info={})

        res = offline.sharded_sweep(run, grid, group_cols=(0, 1), gather=False)
        L, idx, path = offline.select_sharded(res)
        Lw, iw = _cpu_scan(torch.from_numpy(f), torch.from_numpy(m))
        out.append((idx == iw and (iw < 0 or float(L) == float(Lw)), path))
    q.put((rank, out))
    dist.destroy_process_group()


def test_packed_key_selection_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_key_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    for _, out in outs:
        assert all(ok for ok, _ in out), out
        assert out[3][1] == "gather" and out[0][1] == "key"
    assert outs[0][1] == outs[1][1]


def test_result_file_checker_and_stage_checkpoint(tmp_path):
    """f-3: the on-disk format of KG_main.py:175-205 / B_main.py:190-224 as the plotting scripts read it."""
    from gptpinn import results
    rng = np.random.default_rng(0)
    n, C, NR, NT = 3, 4, 16, 9
    d = tmp_path / "kg_data"
    d.mkdir()
    params = {"Device": "cuda", "Domain": {}, "Data sizes": {"Nc": 4, "N_test": 3, "BC_pts": 2, "IC_pts": 2}, "layers_pinn": [2, 4, 1],
              "lr_pinn": 1e-3, "epochs_pinn": 75000, "parameter size": 1000, "number_of_neurons": n, "lr_gpt": 0.025,
              "epochs_gpt_train": 2000, "test_cases": C, "epochs_gpt_test": 5000}
    np.save(d / "params.npy", params)
    neurons = rng.random((n, 3))
    files = {"generation_time.dat": rng.random(n), "loss_list.dat": rng.random(n), "neurons.dat": neurons, "total_time.dat": np.array([0.5]),
             "xt_resid.dat": rng.random((NR, 2)), "kg_test.dat": rng.random((C, 3)), "xt_test.dat": rng.random((NT, 2)),
             "gpt_test_losses.dat": rng.random((11, C)), "gpt_test_soln.dat": rng.random((NT, C)), "gpt_test_time.dat": 0.5 + np.arange(1, C + 1) * 0.01,
             "pinn_test_losses.dat": rng.random((101, C)), "pinn_test_soln.dat": rng.random((NT, C)), "pinn_test_time.dat": np.arange(1, C + 1) * 0.1}
    for k, v in files.items():
        np.savetxt(d / k, v)
    got = results.check_results(str(d), "kg")
    assert np.allclose(got["neurons.dat"], neurons)
    np.savetxt(d / "gpt_test_losses.dat", rng.random((10, C)))
    with pytest.raises(results.ResultFormatError, match="gpt_test_losses"):
        results.check_results(str(d), "kg")
    np.savetxt(d / "gpt_test_losses.dat", files["gpt_test_losses.dat"])
    neurons[2] = neurons[0]
    np.savetxt(d / "neurons.dat", neurons)
    with pytest.raises(results.ResultFormatError, match="repeated"):
        results.check_results(str(d), "kg")
    os.remove(d / "neurons.dat")
    with pytest.raises(results.ResultFormatError, match="missing"):
        results.check_results(str(d), "kg")
    # per-stage checkpoint (the reference has none)
    ck = tmp_path / "ck"
    assert results.load_stage(str(ck)) is None
    results.save_stage(str(ck), 0, neurons[:1], [0.3], rng.random((5, 3)))
    results.save_stage(str(ck), 1, neurons[:2], [0.3, 0.2], rng.random((4, 3)), extra={"W0": rng.random((4, 2))})
    st = results.load_stage(str(ck))
    assert st["stage"] == 1 and st["grid"].shape == (4, 3) and st["x_W0"].shape == (4, 2) and list(st["loss_list"]) == [0.3, 0.2]


def test_reference_copy_is_byte_identical():
    """oracle/_ref (made by oracle/fetch_ref.sh, git-ignored) must be the unmodified reference."""
    import hashlib
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(ref):
        if not os.path.isdir("/root/reference"):
            pytest.skip("no reference checkout and no copy")
        subprocess.check_call(["sh", os.path.join(ROOT, "oracle", "fetch_ref.sh")])
    n = 0
    for line in open(os.path.join(ref, "MANIFEST.sha256")):
        h, name = line.split()
        data = open(os.path.join(ref, name), "rb").read()
        assert hashlib.sha256(data).hexdigest() == h, name
        src = os.path.join("/root/reference", name)
        if os.path.exists(src):
            assert open(src, "rb").read() == data, name
        n += 1
    assert n >= 25
    tracked = subprocess.run(["git", "-C", ROOT, "ls-files", "oracle/_ref"], capture_output=True, text=True).stdout.strip()
    assert tracked == "", "reference sources must not enter the history"


def test_training_memo_key_follows_identity_version_and_arguments():
    """gptpinn.pinn memo of finished test-set trainings: a hit needs the same parameters, epochs, lr, tol, batching and the
    same data tensors (object identity AND in-place version); a freed tensor invalidates the entry."""
    import torch
    from gptpinn import pinn
    a, b = torch.zeros(5, 2), torch.ones(7)
    mus = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    key = lambda **kw: pinn.train_memo_key(kw.get("pde", "kg"), kw.get("layers", [2, 40, 40, 1]), kw.get("mus", mus), kw.get("epochs", 100),
                                           kw.get("lr", 5e-4), kw.get("tol"), kw.get("chunks", [(0, 2)]), kw.get("tensors", (a, b, None)))
    pinn.train_memo_clear()
    k0 = key()
    assert key() == k0 and hash(k0) is not None
    for other in (key(pde="burgers"), key(layers=[2, 20, 1]), key(mus=mus + 1e-12), key(epochs=101), key(lr=1e-3), key(tol=1e-4),
                  key(chunks=[(0, 1), (1, 2)]), key(tensors=(a.clone(), b, None)), key(tensors=(a, b, b))):
        assert other != k0
    pinn.train_memo_put(k0, (a, b, None), ["rec"])
    assert pinn.train_memo_get(k0) == ["rec"]
    a.add_(1.0)                                           # in-place write moves the version counter
    assert key() != k0 and pinn.train_memo_get(key()) is None
    k1 = key()
    pinn.train_memo_put(k1, (a, b, None), ["rec1"])
    assert pinn.train_memo_get(k1) == ["rec1"]
    os.environ["GPTPINN_NO_TRAIN_MEMO"] = "1"
    try:
        assert pinn.train_memo_get(k1) is None
    finally:
        del os.environ["GPTPINN_NO_TRAIN_MEMO"]
    del b                                                 # the caller's tensor goes away: the entry can never be hit again
    import gc
    gc.collect()
    b = torch.ones(7)
    assert pinn.train_memo_get(key()) is None
    for i in range(10):
        pinn.train_memo_put(("k", i), (a,), [i])
    assert len(pinn._train_memo) <= pinn.TRAIN_MEMO_ENTRIES
    pinn.train_memo_clear()
