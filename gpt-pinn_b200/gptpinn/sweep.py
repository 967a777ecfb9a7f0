"""Tensor-level host API over the C ABI: the coefficient sweep and the exact arg-max scan.

Inputs are the tensors the reference passes around (fp32 CUDA tensors, possibly column-slice
views `out[:, 0:n]` with row stride = number_of_neurons) and the host float64 parameter grid.
Everything is enqueued on the current torch CUDA stream; `argmax_scan` synchronises once to read back
(L, index).  The optional targets f_hat / IC_u2 / BC_u are all-zero in every shipped problem and are then
passed as NULL; whether a tensor is all-zero is checked once per tensor object and version (one tiny
reduction + sync the first time a given tensor is seen, none afterwards).
"""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import ACProblem, BProblem, KGProblem, Mat, SweepArgs, SweepInfo


def _req_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise _lib.GptPinnError(f"{name} is on {t.device}: the B200 path takes CUDA tensors only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t


def _mat(t, name, keep, n=None):
    t = _req_cuda(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D, got shape {tuple(t.shape)}")
    if n is not None and t.shape[1] != n:
        raise ValueError(f"{name} has {t.shape[1]} columns, expected {n}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    if ld < t.shape[1]:
        t = t.contiguous()
        ld = t.shape[1]
    keep.append(t)
    return Mat(t.data_ptr(), ld)


def _vec(t, name, keep, rows):
    t = _req_cuda(t, name).reshape(-1)
    if t.numel() != rows:
        raise ValueError(f"{name} has {t.numel()} entries, expected {rows}")
    t = t.contiguous()
    keep.append(t)
    return t.data_ptr()


_zero_cache = {}


def _is_all_zero(t):
    """f_hat / IC_u2 / BC_u are zero in every shipped problem; an all-zero target is passed as NULL so the
    kernel skips it.  The answer is cached per tensor object and autograd version counter (in-place writes bump
    it), so a greedy loop pays the reduction + host sync once per target, not once per call."""
    ent = _zero_cache.get(id(t))
    if ent is not None and ent[0]() is t and ent[1] == t._version:
        return ent[2]
    res = not bool(torch.any(t != 0).item())
    if len(_zero_cache) > 256:
        for k in [k for k, v in _zero_cache.items() if v[0]() is None]:
            del _zero_cache[k]
    _zero_cache[id(t)] = (weakref.ref(t), t._version, res)
    return res


def _opt_vec(t, name, keep, rows):
    if t is None:
        return None
    t = _req_cuda(t, name)
    if _is_all_zero(t):
        return None
    return _vec(t, name, keep, rows)


def _run(fn_name, prob, keep, grid, gcols, n, c0, epochs, lr, rec_every, want_c, cfg, device, bound=None, key_index=None):
    grid = np.ascontiguousarray(np.asarray(grid, dtype=np.float64).reshape(-1, gcols))
    Np = grid.shape[0]
    c0 = _req_cuda(c0, "c_initial")
    if c0.dim() == 1:
        per = 0
        if c0.numel() != n:
            raise ValueError(f"c_initial has {c0.numel()} entries, expected {n}")
    else:
        per = 1
        if tuple(c0.shape) != (Np, n):
            raise ValueError(f"per-parameter c_initial must be [{Np}, {n}], got {tuple(c0.shape)}")
    c0 = c0.contiguous()
    f = torch.empty(Np, dtype=torch.float32, device=device)
    m = torch.empty(Np, dtype=torch.float32, device=device)
    c = torch.empty((Np, n), dtype=torch.float32, device=device) if want_c else None
    trace = None
    if rec_every and rec_every > 0:
        trace = torch.zeros((Np, epochs // rec_every + 1), dtype=torch.float32, device=device)
    args = SweepArgs()
    args.grid_host = grid.ctypes.data_as(C.POINTER(C.c_double))
    args.Np = Np
    args.c0_dev = c0.data_ptr()
    args.c0_per_param = per
    args.epochs = int(epochs)
    args.lr = float(lr)
    args.rec_every = int(rec_every or 0)
    args.c_out_dev = c.data_ptr() if c is not None else None
    args.f_out_dev = f.data_ptr()
    args.m_out_dev = m.data_ptr()
    args.trace_dev = trace.data_ptr() if trace is not None else None
    if bound is not None:
        bound = _req_cuda(bound, "bound").reshape(-1).contiguous()
        if bound.numel() != Np:
            raise ValueError(f"bound has {bound.numel()} entries, expected {Np}")
        args.bound_dev = bound.data_ptr()
    key = None
    if key_index is not None:
        # fused arg-max epilogue: the kernel maximises (float_bits(f) << 32) | ~global_index over this call's rows
        key_index = key_index.to(device=device, dtype=torch.int32).contiguous()
        if key_index.numel() != Np:
            raise ValueError(f"key_index has {key_index.numel()} entries, expected {Np}")
        key = torch.zeros(1, dtype=torch.int64, device=device)
        args.key_out_dev, args.key_index_dev = key.data_ptr(), key_index.data_ptr()
    if cfg:
        args.cfg_M, args.cfg_SL, args.cfg_W = (int(cfg.get("M", 0)), int(cfg.get("SL", 0)), int(cfg.get("W", 0)))
        args.cfg_mode = int(cfg.get("mode", 0))
        args.cfg_K = int(cfg.get("K", 0))
    info = SweepInfo()
    with torch.cuda.device(device):
        h = _lib.handle(device.index if device.index is not None else torch.cuda.current_device())
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = getattr(_lib.lib(), fn_name)(h, C.byref(prob), C.byref(args), C.byref(info), C.c_void_p(stream))
    _lib.check(rc, fn_name)
    return dict(c=c, f=f, m=m, trace=trace, key=key, info=info.asdict(), _keep=(keep, c0, grid, bound, key_index))


def kg_sweep(grid, c0, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, xcos, f_hat, IC_u1, IC_u2, BC_u,
             epochs, lr, rec_every=0, want_c=True, cfg=None, bound=None, key_index=None):
    """All-epochs Klein-Gordon sweep over `grid` [Np,3] (alpha, beta, gamma)."""
    keep = []
    n = out.shape[1]
    NR, NI, NB = out.shape[0], out_IC.shape[0], out_BC.shape[0]
    p = KGProblem()
    p.n, p.NR, p.NI, p.NB = n, NR, NI, NB
    p.P, p.Pxx, p.Ptt = _mat(out, "out", keep, n), _mat(out_xx, "out_xx", keep, n), _mat(out_tt, "out_tt", keep, n)
    p.PIC, p.PICt = _mat(out_IC, "out_IC", keep, n), _mat(out_IC_t, "out_IC_t", keep, n)
    p.PBC = _mat(out_BC, "out_BC", keep, n)
    p.xcos = _vec(xcos, "xcos_x2cos2", keep, NR)
    p.fhat = _opt_vec(f_hat, "f_hat", keep, NR)
    p.ICu1 = _vec(IC_u1, "IC_u1", keep, NI)
    p.ICu2 = _opt_vec(IC_u2, "IC_u2", keep, NI)
    p.BCu = _vec(BC_u, "BC_u", keep, NB)
    return _run("gptpinn_kg_sweep", p, keep, grid, 3, n, c0, epochs, lr, rec_every, want_c, cfg, out.device, bound, key_index)


def b_sweep(grid, c0, out, out_x, out_t, out_xx, out_IC, out_BC, f_hat, IC_u, BC_u,
            epochs, lr, rec_every=0, want_c=True, cfg=None, bound=None, key_index=None):
    """All-epochs Burgers sweep over `grid` [Np] (nu)."""
    keep = []
    n = out.shape[1]
    NR, NI, NB = out.shape[0], out_IC.shape[0], out_BC.shape[0]
    p = BProblem()
    p.n, p.NR, p.NI, p.NB = n, NR, NI, NB
    p.P, p.Px = _mat(out, "out", keep, n), _mat(out_x, "out_x", keep, n)
    p.Pt, p.Pxx = _mat(out_t, "out_t", keep, n), _mat(out_xx, "out_xx", keep, n)
    p.PIC, p.PBC = _mat(out_IC, "out_IC", keep, n), _mat(out_BC, "out_BC", keep, n)
    p.fhat = _opt_vec(f_hat, "f_hat", keep, NR)
    p.ICu = _vec(IC_u, "IC_u", keep, NI)
    p.BCu = _opt_vec(BC_u, "BC_u", keep, NB)
    return _run("gptpinn_b_sweep", p, keep, grid, 1, n, c0, epochs, lr, rec_every, want_c, cfg, out.device, bound, key_index)


def ac_sweep(grid, c0, out, out_t, out_xx, out_IC, out_BC_ub, out_BC_lb, out_BC_diff, out_BC_ub_x,
             out_BC_lb_x, out_BC_diff_x, f_hat, IC_u, epochs, lr, rec_every=0, want_c=True, cfg=None, bound=None,
             t_given=False, key_index=None):
    """All-epochs Allen-Cahn sweep over `grid` [Np,2] (lambda, eps).  With t_given, `out_t` is the materialised
    T = P_t - lambda P_xx - eps P (used bit for bit; `out_xx` ignored, only eps of each row is used)."""
    keep = []
    n = out.shape[1]
    NR, NI, NB = out.shape[0], out_IC.shape[0], out_BC_ub.shape[0]
    p = ACProblem()
    p.n, p.NR, p.NI, p.NB = n, NR, NI, NB
    p.P, p.Pt, p.Pxx = _mat(out, "out", keep, n), _mat(out_t, "out_t", keep, n), _mat(out_xx, "out_xx", keep, n)
    p.PIC = _mat(out_IC, "out_IC", keep, n)
    p.Pub, p.Plb = _mat(out_BC_ub, "out_BC_ub", keep, n), _mat(out_BC_lb, "out_BC_lb", keep, n)
    p.Pdiff = _mat(out_BC_diff, "out_BC_diff", keep, n)
    p.Pubx, p.Plbx = _mat(out_BC_ub_x, "out_BC_ub_x", keep, n), _mat(out_BC_lb_x, "out_BC_lb_x", keep, n)
    p.Pdiffx = _mat(out_BC_diff_x, "out_BC_diff_x", keep, n)
    p.fhat = _opt_vec(f_hat, "f_hat", keep, NR)
    p.ICu = _vec(IC_u, "IC_u", keep, NI)
    p.t_given = 1 if t_given else 0
    return _run("gptpinn_ac_sweep", p, keep, grid, 2, n, c0, epochs, lr, rec_every, want_c, cfg, out.device, bound, key_index)


def argmax_scan_device(f, m):
    """Enqueue the exact N1 scan; returns device tensors (L[1] float32, idx[1] int32) without syncing."""
    f = _req_cuda(f, "f").contiguous()
    m = _req_cuda(m, "m").contiguous()
    L = torch.zeros(1, dtype=torch.float32, device=f.device)
    idx = torch.full((1,), -1, dtype=torch.int32, device=f.device)
    if f.numel() == 0:
        return L, idx                      # empty grid: largest_loss stays 0, no case (KG_train.py:12)
    with torch.cuda.device(f.device):
        stream = torch.cuda.current_stream(f.device).cuda_stream
        rc = _lib.lib().gptpinn_argmax_scan(f.data_ptr(), m.data_ptr(), f.numel(), L.data_ptr(), idx.data_ptr(),
                                            C.c_void_p(stream))
    _lib.check(rc, "gptpinn_argmax_scan")
    return L, idx


def argmax_scan(f, m):
    """(largest_loss tensor (0-dim, device), winning grid index or -1).  One host sync."""
    L, idx = argmax_scan_device(f, m)
    i = int(idx.item())
    return L[0], i
