"""ctypes wrapper around oracle/_build/liboracle.so (the CPU restatement).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product path (gpt-pinn_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_fp = C.POINTER(C.c_float)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("gptpinn_oracle.c", "gptpinn_oracle_impl.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


class KGProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("ld", C.c_int), ("NR", C.c_int), ("NI", C.c_int), ("NB", C.c_int),
                ("P", _fp), ("Pxx", _fp), ("Ptt", _fp), ("PIC", _fp), ("PICt", _fp), ("PBC", _fp),
                ("xcos", _fp), ("fhat", _fp), ("ICu1", _fp), ("ICu2", _fp), ("BCu", _fp)]


class BProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("ld", C.c_int), ("NR", C.c_int), ("NI", C.c_int), ("NB", C.c_int),
                ("P", _fp), ("Px", _fp), ("Pt", _fp), ("Pxx", _fp), ("PIC", _fp), ("PBC", _fp),
                ("fhat", _fp), ("ICu", _fp), ("BCu", _fp)]


class ACProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("ld", C.c_int), ("NR", C.c_int), ("NI", C.c_int), ("NB", C.c_int),
                ("P", _fp), ("Pt", _fp), ("Pxx", _fp), ("PIC", _fp),
                ("Pub", _fp), ("Plb", _fp), ("Pdiff", _fp), ("Pubx", _fp), ("Plbx", _fp), ("Pdiffx", _fp),
                ("fhat", _fp), ("ICu", _fp)]


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return a.ctypes.data_as(_fp)


class _Holder:
    """Keeps the numpy arrays alive next to the ctypes struct."""

    def __init__(self, struct, keep):
        self.struct, self.keep = struct, keep


def _mats(n, mats):
    """mats: dict name -> 2-D array with >= n columns.  Uses the leading n columns, packed (ld = n)."""
    return {k: _f32(np.asarray(v)[:, :n]) for k, v in mats.items()}


def kg_problem(n, out, out_xx, out_tt, out_IC, out_IC_t, out_BC, xcos, fhat, IC_u1, IC_u2, BC_u):
    m = _mats(n, dict(P=out, Pxx=out_xx, Ptt=out_tt, PIC=out_IC, PICt=out_IC_t, PBC=out_BC))
    v = {k: _f32(np.asarray(a).reshape(-1)) for k, a in
         dict(xcos=xcos, fhat=fhat, ICu1=IC_u1, ICu2=IC_u2, BCu=BC_u).items()}
    s = KGProblem(n, n, m["P"].shape[0], m["PIC"].shape[0], m["PBC"].shape[0],
                  _ptr(m["P"]), _ptr(m["Pxx"]), _ptr(m["Ptt"]), _ptr(m["PIC"]), _ptr(m["PICt"]),
                  _ptr(m["PBC"]), _ptr(v["xcos"]), _ptr(v["fhat"]), _ptr(v["ICu1"]), _ptr(v["ICu2"]),
                  _ptr(v["BCu"]))
    return _Holder(s, (m, v))


def b_problem(n, out, out_x, out_t, out_xx, out_IC, out_BC, fhat, IC_u, BC_u):
    m = _mats(n, dict(P=out, Px=out_x, Pt=out_t, Pxx=out_xx, PIC=out_IC, PBC=out_BC))
    v = {k: _f32(np.asarray(a).reshape(-1)) for k, a in dict(fhat=fhat, ICu=IC_u, BCu=BC_u).items()}
    s = BProblem(n, n, m["P"].shape[0], m["PIC"].shape[0], m["PBC"].shape[0],
                 _ptr(m["P"]), _ptr(m["Px"]), _ptr(m["Pt"]), _ptr(m["Pxx"]), _ptr(m["PIC"]),
                 _ptr(m["PBC"]), _ptr(v["fhat"]), _ptr(v["ICu"]), _ptr(v["BCu"]))
    return _Holder(s, (m, v))


def ac_problem(n, out, out_t, out_xx, out_IC, out_BC_ub, out_BC_lb, out_BC_diff, out_BC_ub_x,
               out_BC_lb_x, out_BC_diff_x, fhat, IC_u):
    m = _mats(n, dict(P=out, Pt=out_t, Pxx=out_xx, PIC=out_IC, Pub=out_BC_ub, Plb=out_BC_lb,
                      Pdiff=out_BC_diff, Pubx=out_BC_ub_x, Plbx=out_BC_lb_x, Pdiffx=out_BC_diff_x))
    v = {k: _f32(np.asarray(a).reshape(-1)) for k, a in dict(fhat=fhat, ICu=IC_u).items()}
    s = ACProblem(n, n, m["P"].shape[0], m["PIC"].shape[0], m["Pub"].shape[0],
                  _ptr(m["P"]), _ptr(m["Pt"]), _ptr(m["Pxx"]), _ptr(m["PIC"]), _ptr(m["Pub"]),
                  _ptr(m["Plb"]), _ptr(m["Pdiff"]), _ptr(m["Pubx"]), _ptr(m["Plbx"]),
                  _ptr(m["Pdiffx"]), _ptr(v["fhat"]), _ptr(v["ICu"]))
    return _Holder(s, (m, v))


def sweep(kind, prob, grid, c0, epochs, lr, rec_every=0, precision="f32"):
    """Run E epochs for every grid row.  Returns dict(c, f, m, trace)."""
    n = prob.struct.n
    grid = np.ascontiguousarray(np.asarray(grid, dtype=np.float64))
    Np = grid.shape[0]
    c0 = _f32(np.broadcast_to(np.asarray(c0, dtype=np.float32), (Np, n)))
    c = np.zeros((Np, n), np.float32)
    f = np.zeros(Np, np.float32)
    m = np.zeros(Np, np.float32)
    nrec = epochs // rec_every + 1 if rec_every > 0 else 0
    tr = np.zeros((Np, max(nrec, 1)), np.float32)
    fn = getattr(lib(), f"oracle_{kind}_sweep_{precision}")
    fn.restype = C.c_int
    rc = fn(C.byref(prob.struct), grid.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(Np), _ptr(c0),
            C.c_int(epochs), C.c_double(lr), C.c_int(rec_every), _ptr(c), _ptr(f), _ptr(m),
            _ptr(tr) if nrec else None)
    if rc != 0:
        raise RuntimeError(f"oracle sweep failed rc={rc}")
    return dict(c=c, f=f, m=m, trace=tr if nrec else None)


def kg_offline_literal(prob, grid, c0, epochs, lr, precision="f32"):
    n = prob.struct.n
    grid = np.ascontiguousarray(np.asarray(grid, dtype=np.float64))
    Np = grid.shape[0]
    c0 = _f32(np.broadcast_to(np.asarray(c0, dtype=np.float32), (Np, n)))
    L = C.c_float(0)
    idx = C.c_int(-1)
    fn = getattr(lib(), f"oracle_kg_offline_{precision}")
    fn.restype = C.c_int
    rc = fn(C.byref(prob.struct), grid.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(Np), _ptr(c0),
            C.c_int(epochs), C.c_double(lr), C.byref(L), C.byref(idx))
    if rc != 0:
        raise RuntimeError(f"oracle offline failed rc={rc}")
    return L.value, idx.value


def argmax_scan(f, m):
    f = _f32(f)
    m = _f32(m)
    L = C.c_float(0)
    idx = C.c_int(-1)
    lib().oracle_argmax_scan(_ptr(f), _ptr(m), C.c_int(f.shape[0]), C.byref(L), C.byref(idx))
    return L.value, idx.value


def mlp_channels(layers, weights, biases, act, xt):
    """Forward-mode value/derivative channels.  Returns dict v,x,t,q,r of [N] float32."""
    layers = np.ascontiguousarray(np.asarray(layers, dtype=np.int32))
    nl = len(layers)
    Ws = [_f32(w) for w in weights]
    bs = [_f32(b).reshape(-1) for b in biases]
    Wp = (_fp * (nl - 1))(*[_ptr(w) for w in Ws])
    bp = (_fp * (nl - 1))(*[_ptr(b) for b in bs])
    xt = _f32(xt)
    N = xt.shape[0]
    outs = {k: np.zeros(N, np.float32) for k in "vxtqr"}
    actid = {"cos": 0, "tanh": 1}[act]
    rc = lib().oracle_mlp_channels(layers.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(nl), Wp, bp,
                                   C.c_int(actid), _ptr(xt), C.c_int(N),
                                   *[_ptr(outs[k]) for k in "vxtqr"])
    if rc != 0:
        raise RuntimeError("oracle mlp failed")
    return outs
