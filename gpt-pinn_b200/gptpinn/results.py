"""On-disk result files of the reference drivers: format check and a per-stage checkpoint.

The reference's `KG_main.py:175-205` / `B_main.py:190-224` write their results with `np.savetxt`
(+ one pickled `params.npy`); `KG_plotting.py:11-26` / `B_plotting.py:10-25` read them back with
`np.loadtxt`.  With the drop-in modules the unchanged mains still write these files; `check_results`
verifies names, shapes (as `np.loadtxt` returns them), finiteness and the few cross-file invariants
the plotting scripts rely on.  `save_stage` / `load_stage` add what the reference lacks: a
checkpoint of the greedy state after every neuron stage, so a long offline run can resume.
"""
import os

import numpy as np

# name -> shape as np.loadtxt returns it, in terms of: n (neurons), g (grid dimension), NR, NT (test points),
# C (test cases), LG (gpt loss records), LP (pinn loss records), E (stages that ran)
_KG = {
    "generation_time.dat": ("n",), "loss_list.dat": ("n",), "neurons.dat": ("n", "g"), "total_time.dat": (),
    "xt_resid.dat": ("NR", 2), "kg_test.dat": ("C", "g"), "xt_test.dat": ("NT", 2),
    "gpt_test_losses.dat": ("LG", "C"), "gpt_test_soln.dat": ("NT", "C"), "gpt_test_time.dat": ("C",),
    "pinn_test_losses.dat": ("LP", "C"), "pinn_test_soln.dat": ("NT", "C"), "pinn_test_time.dat": ("C",),
}
_B = {
    "generation_time.dat": ("E",), "max_losses.dat": ("E",), "neurons.dat": ("n",), "total_time.dat": (),
    "xt_resid.dat": ("NR", 2), "b_test.dat": ("C",), "xt_test.dat": ("NT", 2),
    "gpt_test_losses.dat": ("LG", "C"), "gpt_test_soln.dat": ("NT", "C"), "gpt_test_time.dat": ("C",),
    "pinn_test_losses.dat": ("LP", "C"), "pinn_test_soln.dat": ("NT", "C"), "pinn_test_time.dat": ("C",),
}
SPEC = {"kg": _KG, "burgers": _B}
PARAM_KEYS = {
    "kg": ("Device", "Domain", "Data sizes", "layers_pinn", "lr_pinn", "epochs_pinn", "parameter size",
           "number_of_neurons", "lr_gpt", "epochs_gpt_train", "test_cases", "epochs_gpt_test"),
    "burgers": ("Device", "Domain", "Data sizes", "tol", "layers_pinn", "lr_pinn", "epochs_pinn", "parameter size",
                "number_of_neurons", "lr_gpt", "epochs_gpt_train", "test_cases", "epochs_gpt_test", "num_largest_mag"),
}


class ResultFormatError(ValueError):
    pass


def expected_dims(pde, params):
    """Symbol table for SPEC from the `params.npy` dictionary the main wrote."""
    sizes = params["Data sizes"]
    n = int(params["number_of_neurons"])
    d = {"n": n, "C": int(params["test_cases"]), "NR": int(sizes["Nc"]) ** 2, "NT": int(sizes["N_test"]) ** 2,
         "LG": 11, "g": 3 if pde == "kg" else 1, "E": n}
    ep = int(params["epochs_pinn"])
    # KG_test.py:119 zeros((101, .)) for 75 000 epochs / 1000; B_test.py:157 zeros((61, .)) for 60 000 / 1000
    d["LP"] = 101 if pde == "kg" else 61
    d["epochs_pinn"] = ep
    return d


def check_results(directory, pde):
    """Raise ResultFormatError unless `directory` holds the complete, well-formed result set of `pde`'s main.
    Returns {file name: array} (what the plotting script would load)."""
    spec = SPEC[pde]
    ppath = os.path.join(directory, "params.npy")
    if not os.path.exists(ppath):
        raise ResultFormatError(f"{ppath} missing")
    params = np.load(ppath, allow_pickle=True).item()
    miss = [k for k in PARAM_KEYS[pde] if k not in params]
    if miss:
        raise ResultFormatError(f"params.npy lacks keys {miss}")
    dims = expected_dims(pde, params)
    out = {}
    for name, shp in spec.items():
        path = os.path.join(directory, name)
        if not os.path.exists(path):
            raise ResultFormatError(f"{name} missing")
        a = np.loadtxt(path)
        want = tuple(dims[s] if isinstance(s, str) else s for s in shp)
        if "E" in shp and a.shape != want:          # Burgers with train_final = False writes one stage fewer
            want_alt = tuple((dims[s] - 1 if s == "E" else dims[s]) if isinstance(s, str) else s for s in shp)
            if a.shape == want_alt:
                want = want_alt
        # np.loadtxt squeezes a single test case / single column
        if a.shape != want and a.shape != tuple(x for x in want if x != 1):
            raise ResultFormatError(f"{name}: shape {a.shape}, expected {want}")
        if a.dtype != np.float64:
            raise ResultFormatError(f"{name}: dtype {a.dtype}, expected float64")
        if not np.all(np.isfinite(a)):
            raise ResultFormatError(f"{name}: non-finite entries")
        out[name] = a
    # cross-file invariants the plotting scripts rely on
    tt = float(out["total_time.dat"])
    if tt <= 0:
        raise ResultFormatError("total_time.dat must be positive")
    g = out["gpt_test_time.dat"]
    if np.any(np.diff(np.atleast_1d(g)) < 0) or np.atleast_1d(g)[0] < tt:
        raise ResultFormatError("gpt_test_time.dat must be cumulative hours on top of total_time")
    if np.any(np.diff(np.atleast_1d(out["pinn_test_time.dat"])) < 0):
        raise ResultFormatError("pinn_test_time.dat must be cumulative")
    losses = out["loss_list.dat" if pde == "kg" else "max_losses.dat"]
    if np.any(losses < 0):
        raise ResultFormatError("largest losses must be non-negative")
    neurons = out["neurons.dat"]
    rows = neurons if neurons.ndim == 2 else neurons.reshape(-1, 1)
    if len({tuple(r) for r in rows}) != len(rows):
        raise ResultFormatError("neurons.dat holds a repeated parameter (a chosen neuron must leave the grid)")
    test = out["kg_test.dat" if pde == "kg" else "b_test.dat"]
    trows = test if test.ndim == 2 else test.reshape(-1, 1)
    if {tuple(r) for r in trows} & {tuple(r) for r in rows}:
        raise ResultFormatError("a test parameter coincides with a neuron (test cases are drawn from the remaining grid)")
    out["params.npy"] = params
    return out


def save_stage(directory, stage, neurons, loss_list, grid, extra=None):
    """Checkpoint after neuron stage `stage` (0-based): chosen neurons so far, their largest losses and the remaining
    grid.  Written atomically (tmp + rename) as stage_XX.npz; `extra` may carry PINN weights (name -> ndarray)."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"stage_{stage:02d}.npz")
    tmp = path + ".tmp.npz"
    payload = {"stage": np.int64(stage), "neurons": np.asarray(neurons, np.float64),
               "loss_list": np.asarray(loss_list, np.float64), "grid": np.asarray(grid, np.float64)}
    for k, v in (extra or {}).items():
        payload[f"x_{k}"] = np.asarray(v)
    np.savez(tmp, **payload)
    os.replace(tmp, path)
    return path


def load_stage(directory):
    """Latest complete checkpoint in `directory`, or None."""
    if not os.path.isdir(directory):
        return None
    names = sorted(f for f in os.listdir(directory) if f.startswith("stage_") and f.endswith(".npz") and ".tmp" not in f)
    for name in reversed(names):
        try:
            z = np.load(os.path.join(directory, name))
            d = {k: z[k] for k in z.files}
            d["stage"] = int(d["stage"])
            return d
        except Exception:
            continue
    return None
