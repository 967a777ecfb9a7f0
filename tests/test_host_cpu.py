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
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
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
