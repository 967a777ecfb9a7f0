"""ctypes binding of libgptpinn_b200.so (C ABI: include/gptpinn_b200.h).

There is no fallback of any kind: if the shared library is missing or the device is not a
B200-class (sm_100) GPU the import / first call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPTPINN_LIB") or os.path.join(_HERE, "lib", "libgptpinn_b200.so")   # env override: kernel-variant experiments only

_fp = C.POINTER(C.c_float)


class Mat(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_longlong)]


class KGProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("NR", C.c_int), ("NI", C.c_int), ("NB", C.c_int),
                ("P", Mat), ("Pxx", Mat), ("Ptt", Mat), ("PIC", Mat), ("PICt", Mat), ("PBC", Mat),
                ("xcos", C.c_void_p), ("fhat", C.c_void_p), ("ICu1", C.c_void_p), ("ICu2", C.c_void_p),
                ("BCu", C.c_void_p)]


class BProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("NR", C.c_int), ("NI", C.c_int), ("NB", C.c_int),
                ("P", Mat), ("Px", Mat), ("Pt", Mat), ("Pxx", Mat), ("PIC", Mat), ("PBC", Mat),
                ("fhat", C.c_void_p), ("ICu", C.c_void_p), ("BCu", C.c_void_p)]


class ACProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("NR", C.c_int), ("NI", C.c_int), ("NB", C.c_int),
                ("P", Mat), ("Pt", Mat), ("Pxx", Mat), ("PIC", Mat),
                ("Pub", Mat), ("Plb", Mat), ("Pdiff", Mat), ("Pubx", Mat), ("Plbx", Mat), ("Pdiffx", Mat),
                ("fhat", C.c_void_p), ("ICu", C.c_void_p), ("t_given", C.c_int)]


class SweepArgs(C.Structure):
    _fields_ = [("grid_host", C.POINTER(C.c_double)), ("Np", C.c_int),
                ("c0_dev", C.c_void_p), ("c0_per_param", C.c_int),
                ("epochs", C.c_int), ("lr", C.c_double), ("rec_every", C.c_int),
                ("c_out_dev", C.c_void_p), ("f_out_dev", C.c_void_p), ("m_out_dev", C.c_void_p),
                ("trace_dev", C.c_void_p),
                ("cfg_M", C.c_int), ("cfg_SL", C.c_int), ("cfg_W", C.c_int), ("cfg_mode", C.c_int),
                ("cfg_K", C.c_int), ("bound_dev", C.c_void_p), ("key_out_dev", C.c_void_p), ("key_index_dev", C.c_void_p)]


class SweepInfo(C.Structure):
    _fields_ = [("M", C.c_int), ("SL", C.c_int), ("W", C.c_int), ("items", C.c_int), ("groups", C.c_int),
                ("ctas", C.c_int), ("rounds", C.c_int), ("kernels_launched", C.c_int),
                ("smem_bytes", C.c_int), ("tiles_per_pass", C.c_int), ("priv", C.c_int), ("K", C.c_int)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Mlp(C.Structure):
    _fields_ = [("nlayers", C.c_int), ("layers", C.c_int * 8), ("W", C.c_void_p * 7), ("b", C.c_void_p * 7),
                ("act", C.c_int)]


class PinnProblem(C.Structure):
    _fields_ = [("pde", C.c_int), ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("NR", C.c_int), ("NI", C.c_int), ("NB", C.c_int),
                ("xt_resid", C.c_void_p), ("f_hat", C.c_void_p), ("xcos", C.c_void_p), ("IC_xt", C.c_void_p),
                ("IC_u1", C.c_void_p), ("IC_u2", C.c_void_p), ("BC_xt", C.c_void_p), ("BC_u", C.c_void_p)]


class TrainArgs(C.Structure):
    _fields_ = [("epochs", C.c_int), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("tol", C.c_double), ("trace_every", C.c_int),
                ("loss_trace_dev", C.c_void_p), ("grad_out_dev", C.c_void_p)]


EXPORTS = ("gptpinn_version", "gptpinn_last_error", "gptpinn_create", "gptpinn_destroy",
           "gptpinn_kg_sweep", "gptpinn_b_sweep", "gptpinn_ac_sweep", "gptpinn_argmax_scan",
           "gptpinn_mlp_channels", "gptpinn_ffma_peak", "gptpinn_kg_plan", "gptpinn_pinn_train",
           "gptpinn_pinn_train_batch",
           "gptpinn_sa_layer0", "gptpinn_sa_act_forward", "gptpinn_sa_act_backward", "gptpinn_sa_output_forward",
           "gptpinn_sa_loss_workspace_floats", "gptpinn_sa_loss", "gptpinn_sa_output_backward", "gptpinn_sa_tick", "gptpinn_sa_adam",
           "gptpinn_sa_lbfgs_max_params", "gptpinn_sa_lbfgs_direction", "gptpinn_sa_workspace_floats", "gptpinn_sa_wgrad",
           "gptpinn_sa_bias_grad", "gptpinn_sa_layer0_grad", "gptpinn_sa_map_forward", "gptpinn_sa_map_adjoint")

_lib = None


class GptPinnError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GptPinnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C gpt-pinn_b200/csrc`.  There is no CPU / PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    L.gptpinn_version.restype = C.c_int
    L.gptpinn_sa_workspace_floats.restype = C.c_longlong
    L.gptpinn_last_error.restype = C.c_char_p
    L.gptpinn_create.argtypes = [C.POINTER(C.c_void_p)]
    L.gptpinn_destroy.argtypes = [C.c_void_p]
    for name, prob in (("gptpinn_kg_sweep", KGProblem), ("gptpinn_b_sweep", BProblem),
                       ("gptpinn_ac_sweep", ACProblem)):
        fn = getattr(L, name)
        fn.argtypes = [C.c_void_p, C.POINTER(prob), C.POINTER(SweepArgs), C.POINTER(SweepInfo), C.c_void_p]
        fn.restype = C.c_int
    L.gptpinn_argmax_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gptpinn_argmax_scan.restype = C.c_int
    L.gptpinn_mlp_channels.argtypes = [C.POINTER(Mlp), C.c_void_p, C.c_longlong, C.POINTER(C.c_void_p * 5),
                                       C.POINTER(C.c_longlong * 5), C.c_void_p]
    L.gptpinn_mlp_channels.restype = C.c_int
    L.gptpinn_ffma_peak.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_void_p]
    L.gptpinn_ffma_peak.restype = C.c_int
    L.gptpinn_kg_plan.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(SweepArgs), C.POINTER(SweepInfo), C.POINTER(C.c_int * 6), C.POINTER(C.c_int)]
    L.gptpinn_kg_plan.restype = C.c_int
    L.gptpinn_pinn_train.argtypes = [C.c_void_p, C.POINTER(PinnProblem), C.POINTER(Mlp), C.POINTER(TrainArgs), C.c_void_p, C.c_void_p]
    L.gptpinn_pinn_train.restype = C.c_int
    L.gptpinn_pinn_train_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(PinnProblem), C.POINTER(Mlp), C.POINTER(TrainArgs),
                                           C.c_void_p, C.c_void_p]
    L.gptpinn_pinn_train_batch.restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().gptpinn_last_error().decode("utf-8", "replace")
        raise GptPinnError(f"{what} failed (code {rc}): {msg}")


_handles = {}


def handle(device_index):
    """Per-device workspace handle (created lazily; lives for the process)."""
    import torch
    h = _handles.get(device_index)
    if h is None:
        with torch.cuda.device(device_index):
            hp = C.c_void_p()
            check(lib().gptpinn_create(C.byref(hp)), "gptpinn_create")
        _handles[device_index] = h = hp
    return h


def kg_plan(grid, n, NR=10000, NI=512, NB=1024, n_sm=148, cfg=None):
    """Host-only view of the work-item plan (no GPU needed)."""
    import numpy as np
    grid = np.ascontiguousarray(np.asarray(grid, dtype=np.float64).reshape(-1, 3))
    a = SweepArgs()
    if cfg:
        a.cfg_M, a.cfg_SL, a.cfg_W, a.cfg_mode, a.cfg_K = (int(cfg.get(k, 0)) for k in ("M", "SL", "W", "mode", "K"))
    info = SweepInfo()
    hist = (C.c_int * 6)()
    cov = C.c_int(0)
    check(lib().gptpinn_kg_plan(grid.ctypes.data_as(C.POINTER(C.c_double)), grid.shape[0], n, NR, NI, NB, n_sm,
                                C.byref(a), C.byref(info), C.byref(hist), C.byref(cov)), "gptpinn_kg_plan")
    d = info.asdict()
    d["sl_hist"] = {1 << i: hist[i] for i in range(6) if hist[i]}
    d["covered"] = cov.value
    return d
