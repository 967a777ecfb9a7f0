"""Greedy selection on top of the sweep: grid sharding across ranks and the exact arg-max.

Reference semantics (Klein-Gordon/KG_train.py:12-43, Burgers/B_train.py:18-71,
Allen-Cahn/AC_train.py:10-40): visit the grid in order, train c for each parameter with an
early exit against the running maximum, keep the arg-max.  Here every parameter is trained to
the end in one fused launch, and the reference's (largest_loss, largest_case) is recovered
*exactly* from (final loss, min-prefix loss) by `gptpinn_argmax_scan`.

Multi-GPU (one process per GPU, torch.distributed): the grid rows are split into contiguous
blocks, one per rank, with no data-path collective during the sweep; afterwards the two
per-parameter scalars are all-gathered (8 bytes per parameter) and every rank runs the same
scan, so all ranks agree on the next neuron without a broadcast.
"""
import numpy as np
import torch

from . import sweep as _sweep


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
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
    parts = [torch.empty_like(pad) for _ in range(world)]
    d.all_gather(parts, pad)
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_rows, r, world)
        out.append(parts[r][: hi - lo])
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


def sharded_sweep(run, grid, c0=None, want_c=False, gather_c=False, group_cols=None):
    """run(grid_rows, c0_rows_or_None) -> sweep result dict for this rank's rows.

    Returns dict(f, m [, c]) over the FULL grid, in grid order, on every rank (f, m gathered; c
    gathered only when asked)."""
    grid = np.asarray(grid, dtype=np.float64)
    n_rows = grid.shape[0]
    order = shard_order(grid, group_cols)
    lo, hi = shard_bounds(n_rows)
    rows = np.arange(lo, hi) if order is None else order[lo:hi]
    c0_loc = c0
    if c0 is not None and c0.dim() == 2:
        c0_loc = c0[lo:hi] if order is None else c0[torch.from_numpy(rows).to(c0.device)]
    res = run(grid[rows], c0_loc)

    def full(local):
        g = gather_rows(local, n_rows)
        if order is None:
            return g
        out = torch.empty_like(g)
        out[torch.from_numpy(order).to(g.device)] = g
        return out

    out = dict(f=full(res["f"]), m=full(res["m"]), info=res["info"], lo=lo, hi=hi, rows=rows, order=order, local=res)
    if want_c:
        out["c"] = full(res["c"]) if gather_c else res["c"]
    return out


def select(f, m):
    """(largest_loss, index): exact offline_generation semantics; index -1 when nothing exceeds 0."""
    return _sweep.argmax_scan(f, m)


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
