"""Greedy selection on top of the sweep: grid sharding across ranks and the exact arg-max.

Reference semantics (Klein-Gordon/KG_train.py:12-43, Burgers/B_train.py:18-71,
Allen-Cahn/AC_train.py:10-40): visit the grid in order, train c for each parameter with an
early exit against the running maximum, keep the arg-max.  Here every parameter is trained to
the end in one fused launch, and the reference's (largest_loss, largest_case) is recovered
*exactly* from (final loss, min-prefix loss) by `gptpinn_argmax_scan`.

Multi-GPU (one process per GPU, torch.distributed): the grid rows are split into contiguous
blocks, one per rank, with no data-path collective during the sweep.  Selection afterwards:
  * fast path (north-star wording): every rank reduces its shard to ONE 64-bit key
    (float bits of f << 32 | ~global index; losses are >= 0 so bit order = numeric order, the
    complemented index makes the first index win ties), an 8-byte MAX all-reduce gives (F, w), and a
    second 8-byte MAX all-reduce checks the N1 guard `m_w >= max_{p<w} f_p`;
  * fallback (always exact, taken when the guard fails): one packed all-gather of (f, m)
    (8 bytes per parameter) and the same scan on every rank.
Either way all ranks agree on the next neuron without a broadcast.

`early_exit_select` reproduces the reference's early exit (KG_train.py:31-33) as *saved work*:
a short scout pass, a few seed groups trained to the end, then one launch in which every other
parameter stops as soon as its loss falls below a bound that is provably <= the reference's
running maximum at that grid row -- the selection is identical to the fixed-work path.
"""
import os

import numpy as np
import torch

from . import sweep as _sweep


def _dist():
    """torch.distributed when this process is one rank of several, else None.

    Under a multi-process launcher (torchrun sets WORLD_SIZE / RANK / LOCAL_RANK / MASTER_*) the process group is
    created here on first use, so the reference's `*_main.py` run UNCHANGED on N GPUs: no init call has to be added to
    the script.  NCCL when CUDA is there (one rank per GPU), gloo otherwise (host-logic tests)."""
    import torch.distributed as dist
    if not dist.is_available():
        return None
    if not dist.is_initialized():
        if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or "RANK" not in os.environ:
            return None
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if torch.cuda.is_available():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    if dist.get_world_size() > 1:
        return dist
    return None


def shard_bounds(n_rows, rank=None, world=None):
    """Contiguous block [lo, hi) of `n_rows` grid rows owned by `rank` (remainder to the first ranks)."""
    d = _dist()
    if world is None:
        world = d.get_world_size() if d else 1
    if rank is None:
        rank = d.get_rank() if d else 0
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_rows):
    """All-gather per-parameter rows (first dim) from every rank's contiguous shard -> full [n_rows, ...]."""
    d = _dist()
    if d is None:
        return local
    world = d.get_world_size()
    cap = (n_rows + world - 1) // world
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    d.all_gather_into_tensor(full, pad)
    if n_rows == world * cap:
        return full
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_rows, r, world)
        out.append(full[r * cap: r * cap + hi - lo])
    return torch.cat(out, dim=0)


def shard_order(grid, group_cols=None):
    """Row order in which the grid is cut into per-rank blocks.  With `group_cols` (KG: (0, 1) =
    alpha, beta) rows that share the linear operator T stay together, so every rank keeps whole
    sibling groups (the kernel amortises the T build over the siblings); otherwise grid order."""
    n_rows = grid.shape[0]
    if group_cols is None or n_rows == 0 or _dist() is None:
        return None
    keys = tuple(grid[:, c] for c in reversed(group_cols))
    return np.lexsort(keys).astype(np.int64)            # stable: siblings keep their grid order


def sharded_sweep(run, grid, c0=None, want_c=False, gather_c=False, group_cols=None, gather=True, bound=None, **kw):
    """run(grid_rows, c0_rows_or_None, **kw) -> sweep result dict for this rank's rows.

    Returns dict(f, m [, c]) over the FULL grid, in grid order, on every rank (f and m travel in one packed
    all-gather; c gathered only when asked).  With gather=False nothing is exchanged: `f`, `m` are this rank's rows
    only (`rows` = their grid indices) -- the caller reduces them itself (select_sharded).  `bound` ([n_rows] device
    tensor, grid order) is cut to the rank's rows and passed on as run(..., bound=rows_of_bound); other keyword
    arguments go to `run` unchanged."""
    grid = np.asarray(grid, dtype=np.float64)
    n_rows = grid.shape[0]
    order = shard_order(grid, group_cols)
    lo, hi = shard_bounds(n_rows)
    rows = np.arange(lo, hi) if order is None else order[lo:hi]
    c0_loc = c0
    if c0 is not None and c0.dim() == 2:
        c0_loc = c0[lo:hi] if order is None else c0[torch.from_numpy(rows).to(c0.device)]
    if not gather and _dist() is not None:
        kw["key_index"] = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int32))      # fused arg-max epilogue in the kernel
    if bound is not None:
        kw["bound"] = (bound[lo:hi] if order is None else bound[torch.from_numpy(rows).to(bound.device)]).contiguous()
    res = run(grid[rows], c0_loc, **kw)
    out = dict(info=res["info"], lo=lo, hi=hi, rows=rows, order=order, local=res, n_rows=n_rows)
    if not gather:
        out.update(f=res["f"], m=res["m"])
        if want_c:
            out["c"] = res["c"]
        return out

    def full(local):
        g = gather_rows(local, n_rows)
        if order is None:
            return g
        o = torch.empty_like(g)
        o[torch.from_numpy(order).to(g.device)] = g
        return o

    if _dist() is None:
        out.update(f=res["f"], m=res["m"])
    else:
        fm = full(torch.stack((res["f"], res["m"]), dim=1))          # one collective for both scalars
        out.update(f=fm[:, 0].contiguous(), m=fm[:, 1].contiguous())
    if want_c:
        out["c"] = full(res["c"]) if gather_c else res["c"]
    return out


def select(f, m):
    """(largest_loss, index): exact offline_generation semantics; index -1 when nothing exceeds 0."""
    return _sweep.argmax_scan(f, m)


def pack_keys(f, gidx):
    """int64 keys (float_bits(f) << 32) | (0xFFFFFFFF - global index); NaN or non-positive losses map to 0 (never win)."""
    bits = f.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    key = (bits << 32) | (0xFFFFFFFF - gidx.to(torch.int64))
    return torch.where(f > 0, key, torch.zeros_like(key))           # NaN > 0 is False


def select_sharded(res):
    """Selection for a `sharded_sweep(..., gather=False)` result: two 8-byte MAX all-reduces in the common case.

    (F, w) = first arg-max of f over all ranks via the packed key; it is the reference's answer iff the N1 guard
    `!(m_w < max(0, max_{p<w} f_p))` holds -- checked with one more all-reduce that carries both
    the prefix maximum and m_w.  Otherwise falls back to the packed all-gather + exact scan.  Returns
    (largest_loss 0-dim tensor, index, path) with path in {"key", "gather"}."""
    d = _dist()
    f, m = res["f"], res["m"]
    dev = f.device
    n_rows = res["n_rows"]
    gidx = torch.from_numpy(np.asarray(res["rows"], dtype=np.int64)).to(dev)
    if res.get("local") is not None and res["local"].get("key") is not None:
        key = res["local"]["key"].clone()                            # reduced by the sweep kernel's epilogue (atomicMax)
    elif f.numel():
        key = pack_keys(f, gidx).max().reshape(1)
    else:
        key = torch.zeros(1, dtype=torch.int64, device=dev)
    if d is not None:
        d.all_reduce(key, op=d.ReduceOp.MAX)
    k = int(key.item())
    if k == 0:
        return torch.zeros((), dtype=torch.float32, device=dev), -1, "key"
    w = 0xFFFFFFFF - (k & 0xFFFFFFFF)
    F = torch.tensor([k >> 32], dtype=torch.int64, device=dev).to(torch.int32).view(torch.float32)[0]
    # guard: [max_{p<w} f_p (NaN skipped), -m_w] -> MAX all-reduce; only the owner of w contributes m_w
    ninf = torch.full((2,), float("-inf"), dtype=torch.float32, device=dev)
    if f.numel():
        before = (gidx < w) & (f == f)
        ninf[0] = torch.where(before, f, torch.full_like(f, float("-inf"))).max()
        mine = (gidx == w)
        ninf[1] = torch.where(mine, -m, torch.full_like(m, float("-inf"))).max()
    if d is not None:
        d.all_reduce(ninf, op=d.ReduceOp.MAX)
    pm, mw = float(ninf[0].item()), -float(ninf[1].item())
    if not (mw < max(0.0, pm)):
        return F, int(w), "key"
    # guard failed (a non-monotone trajectory dipped below an earlier maximum): exact scan on the gathered arrays
    fm = torch.stack((f, m), dim=1)
    g = gather_rows(fm, n_rows)
    if res["order"] is not None:
        o = torch.empty_like(g)
        o[torch.from_numpy(res["order"]).to(dev)] = g
        g = o
    L, idx = select(g[:, 0].contiguous(), g[:, 1].contiguous())
    return L, idx, "gather"


def winner_row(res, idx, n_rows):
    """c of grid row `idx` (AC's trained_c), fetched from the owning rank when sharded."""
    d = _dist()
    if d is None:
        return res["local"]["c"][idx - res["lo"]].clone()
    world = d.get_world_size()
    pos = idx if res.get("order") is None else int(np.nonzero(res["order"] == idx)[0][0])   # position in shard order
    owner = next(r for r in range(world) if shard_bounds(n_rows, r, world)[0] <= pos < shard_bounds(n_rows, r, world)[1])
    ncol = res["local"]["c"].shape[1]
    row = torch.empty(ncol, dtype=torch.float32, device=res["f"].device)
    if d.get_rank() == owner:
        row.copy_(res["local"]["c"][pos - res["lo"]])
    d.broadcast(row, src=owner)
    return row


def spread_hours(total_seconds, count):
    """Cumulative wall-clock hours per test case, as gpt_test reports them (KG_test.py:37-41): the
    batched sweep has one total, spread linearly over the cases."""
    if count == 0:
        return np.zeros(0)
    return (total_seconds / 3600.0) * (np.arange(1, count + 1) / count)


# ---------------------------------------------------------------------------------------------------------------
# early exit with the reference's semantics, as saved work
# ---------------------------------------------------------------------------------------------------------------
def exit_bounds(f_seed, m_seed, seed_rows, n_rows, margin):
    """Per-row early-exit bound from rows that were trained to the end.

    For every finished row p' the reference's running maximum satisfies L_seq(p) >= min(m_p', f_p') for all p > p'
    (if p' won, L became f_p'; if it lost, either m_p' < L or f_p' <= L already).  So
        B_p = (1 - margin) * max_{seed p' < p} min(m_p', f_p')
    is a valid bound; rows with NaN are skipped as sources.  Rows with no seed before them get -1
    (never exit).  `margin` keeps the decision robust against the last-bit differences between launch plans."""
    dev = f_seed.device
    g = torch.minimum(f_seed, m_seed)
    g = torch.where(g == g, g, torch.full_like(g, float("-inf")))
    arr = torch.full((n_rows + 1,), float("-inf"), dtype=torch.float32, device=dev)
    arr[seed_rows + 1] = g                                           # shifted by one: exclusive prefix
    pref = torch.cummax(arr, dim=0).values[:n_rows]
    b = pref * (1.0 - margin)
    return torch.where(b > 0, b, torch.full_like(b, -1.0))


def early_exit_select(run, grid, epochs, c0=None, group_cols=None, scout_epochs=None, seed_frac=0.02, margin=1e-4):
    """Exact (largest_loss, index) of the reference's offline_generation with most of its skipped work skipped.

    run(rows, c0_rows, epochs=..., bound=...) is the caller's sweep closure.  Three launches per rank:
      1. scout : every row for `scout_epochs` (default epochs/64, at least 4): g1_p = min_j loss_j so far;
      2. seeds : the `seed_frac` of the grid with the largest g1 (whole sibling groups when `group_cols` is given,
                 so the kernel keeps its T sharing) trained for all epochs -> exact (f, m) and the bounds B_p;
      3. rest  : rows already below their bound after the scout are done (they report f = m = g1 < B_p); all others
                 run from scratch with the in-kernel exit `loss_j < B_p` (gptpinn_sweep_args.bound_dev).
    Every row ends up either with its exact (f, m) or with a proof m_p < B_p <= L_seq(p), which the exact scan
    treats like the reference's `break`.  Returns (L, idx, stats)."""
    grid = np.asarray(grid, dtype=np.float64)
    n_rows = grid.shape[0]
    if n_rows == 0:
        return torch.zeros(()), -1, dict(param_epochs=0)
    E1 = int(scout_epochs) if scout_epochs else max(4, epochs // 64)
    if E1 >= epochs or n_rows < 64:
        r = sharded_sweep(run, grid, c0=c0, group_cols=group_cols, epochs=epochs)
        L, idx = select(r["f"], r["m"])
        return L, idx, dict(param_epochs=n_rows * epochs, scout=0, seeds=n_rows, bounded=0, proven_by_scout=0)
    dev = None

    def sub(rows_idx, ep, bound=None):
        """sharded sweep over the grid rows `rows_idx` (ascending grid indices); returns full-length f, m for them"""
        cc = c0
        if c0 is not None and c0.dim() == 2:
            cc = c0[torch.from_numpy(rows_idx).to(c0.device)]
        return sharded_sweep(run, grid[rows_idx], c0=cc, group_cols=group_cols, epochs=ep, bound=bound)

    all_rows = np.arange(n_rows)
    s1 = sub(all_rows, E1)
    dev = s1["f"].device
    g1 = torch.minimum(s1["f"], s1["m"])
    g1 = torch.where(g1 == g1, g1, torch.full_like(g1, float("inf")))        # NaN: never proven by the scout
    g1h = g1.cpu().numpy()
    want = max(1, int(np.ceil(seed_frac * n_rows)))
    if group_cols is not None:
        keys = np.ascontiguousarray(grid[:, list(group_cols)])
        _, inv = np.unique(keys, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        score = np.full(inv.max() + 1, -np.inf)
        np.maximum.at(score, inv, g1h)
        size = np.bincount(inv)
        pick, have = [], 0
        for gi in np.argsort(-score, kind="stable"):
            pick.append(gi)
            have += size[gi]
            if have >= want:
                break
        seed_mask = np.isin(inv, np.asarray(pick))
    else:
        seed_mask = np.zeros(n_rows, dtype=bool)
        seed_mask[np.argsort(-g1h, kind="stable")[:want]] = True
    seed_rows = all_rows[seed_mask]
    s2 = sub(seed_rows, epochs)
    seed_rows_t = torch.from_numpy(seed_rows).to(dev)
    bound = exit_bounds(s2["f"], s2["m"], seed_rows_t, n_rows, margin)
    f = torch.empty(n_rows, dtype=torch.float32, device=dev)
    m = torch.empty(n_rows, dtype=torch.float32, device=dev)
    f[seed_rows_t] = s2["f"]
    m[seed_rows_t] = s2["m"]
    rest_mask = torch.from_numpy(~seed_mask).to(dev)
    proven = rest_mask & (g1 < bound)
    f[proven] = g1[proven]
    m[proven] = g1[proven]
    todo = rest_mask & ~proven
    todo_rows = torch.nonzero(todo).reshape(-1)
    n_todo = int(todo_rows.numel())
    stats = dict(scout=n_rows, scout_epochs=E1, seeds=int(seed_rows.size), proven_by_scout=int(proven.sum().item()), bounded=n_todo)
    if n_todo:
        rows_h = todo_rows.cpu().numpy()
        s3 = sub(rows_h, epochs, bound=bound[todo_rows].contiguous())
        f[todo_rows] = s3["f"]
        m[todo_rows] = s3["m"]
        stats["ran_to_the_end"] = int((~(s3["m"] < bound[todo_rows])).sum().item())
    L, idx = select(f, m)
    stats["param_epochs_upper_bound"] = n_rows * E1 + int(seed_rows.size) * epochs + n_todo * epochs
    return L, idx, stats


EARLY_EXIT = os.environ.get("GPTPINN_EARLY_EXIT", "0") == "1"   # drop-ins: reference early-exit semantics as saved work
last_stats = {}                                                 # what the last greedy_select did (bench / tests)


def greedy_select(run, grid, epochs, c0=None, group_cols=None, want_c=False, early_exit=None):
    """One greedy stage for the drop-in offline_generation functions: sweep every grid row (sharded over the ranks),
    pick (largest_loss, index) with the reference's semantics.  Returns (L, idx, res); res carries `c` of this
    rank's rows when want_c (AC's trained_c is fetched with winner_row)."""
    global last_stats
    grid = np.asarray(grid, dtype=np.float64)
    if early_exit is None:
        early_exit = EARLY_EXIT
    if early_exit and not want_c:
        L, idx, stats = early_exit_select(run, grid, epochs, c0=c0, group_cols=group_cols)
        last_stats = dict(stats, path="early_exit")
        return L, idx, None
    sharded = _dist() is not None
    res = sharded_sweep(run, grid, c0=c0, want_c=want_c, group_cols=group_cols, gather=not sharded, epochs=epochs)
    if sharded:
        L, idx, path = select_sharded(res)
    else:
        (L, idx), path = select(res["f"], res["m"]), "scan"
    last_stats = dict(path=path, info=res["info"], param_epochs=grid.shape[0] * epochs)
    return L, idx, res
