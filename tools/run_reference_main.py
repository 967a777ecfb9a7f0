#!/usr/bin/env python
"""Run the reference's own driver script UNCHANGED on top of the drop-in modules.

  python tools/run_reference_main.py kg      [--out gpurun_out/main_kg]      # oracle/_ref/Klein-Gordon/KG_main.py
  python tools/run_reference_main.py burgers [--out gpurun_out/main_b]       # oracle/_ref/Burgers/B_main.py
  torchrun --nproc-per-node 8 ... tools/run_reference_main.py kg             # same script, grid sharded over 8 GPUs

What this does and does not do:
  * the script file is the byte-identical copy made by oracle/fetch_ref.sh (its sha256 is checked against the
    manifest, and printed); it is executed with runpy, i.e. exactly like `python KG_main.py`;
  * gpt-pinn_b200/dropin is the first entry of sys.path, so `from KG_train import ...` etc. resolve to the drop-in
    modules (equivalent to `cd Klein-Gordon; PYTHONPATH=<repo>/gpt-pinn_b200/dropin python -P KG_main.py`);
  * the working directory is --out, so ./kg_data / ./b_data land there; the log is tee'd to --out/main.log;
  * afterwards the result files are checked with gptpinn.results.check_results and a summary is written to
    --out/summary.json (wall seconds, neurons chosen, largest losses, which native library was loaded).
"""
import argparse
import hashlib
import json
import os
import runpy
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "gpt-pinn_b200", "dropin")
REF = os.path.join(ROOT, "oracle", "_ref")
MAINS = {"kg": ("Klein-Gordon", "KG_main.py", "kg_data"), "burgers": ("Burgers", "B_main.py", "b_data")}


class Tee:
    def __init__(self, *streams):
        self.streams = streams

    def write(self, s):
        for st in self.streams:
            st.write(s)
            st.flush()

    def flush(self):
        for st in self.streams:
            st.flush()


def _last_filled(losses):
    """last recorded row of a [records, cases] loss table (the reference leaves unused trailing rows at zero,
    e.g. rows 76..100 of KG's pinn_test_losses for 75 000 epochs, KG_test.py:119,137)"""
    import numpy as np
    a = np.atleast_2d(losses)
    nz = np.nonzero(np.any(a != 0, axis=1))[0]
    return a[nz[-1]] if nz.size else a[-1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pde", choices=sorted(MAINS))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    folder, script, data = MAINS[args.pde]
    path = os.path.join(REF, folder, script)
    if not os.path.exists(path):
        raise SystemExit(f"{path} missing: run `sh oracle/fetch_ref.sh` in the build container first")
    sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
    manifest = {}
    for line in open(os.path.join(REF, "MANIFEST.sha256")):
        h, name = line.split()
        manifest[os.path.normpath(name)] = h
    want = manifest.get(os.path.normpath(os.path.join(folder, script)))
    if want != sha:
        raise SystemExit(f"{path} does not match the manifest ({sha} vs {want})")

    rank = int(os.environ.get("RANK", "0"))
    out = os.path.abspath(args.out or os.path.join(ROOT, "gpurun_out", f"main_{args.pde}"))
    if rank > 0:
        out = os.path.join(out, f"rank{rank}")          # every rank writes identical files; keep them apart
    os.makedirs(out, exist_ok=True)
    os.chdir(out)
    sys.path.insert(0, DROPIN)
    log = open(os.path.join(out, "main.log"), "w")
    real = sys.stdout
    sys.stdout = Tee(real, log)
    print(f"[run_reference_main] {folder}/{script} sha256 {sha} (unchanged), drop-in modules from {DROPIN}")
    t0 = time.time()
    try:
        runpy.run_path(path, run_name="__main__")
    finally:
        wall = time.time() - t0
        sys.stdout = real
    import numpy as np
    from gptpinn import _lib, results
    res = results.check_results(os.path.join(out, data), args.pde)
    losses = res["loss_list.dat" if args.pde == "kg" else "max_losses.dat"]
    import torch.distributed as dist
    summary = {
        "script": f"{folder}/{script}", "sha256": sha, "unchanged": True, "wall_s": wall,
        "world_size": dist.get_world_size() if dist.is_initialized() else 1,
        "native_library": _lib.LIB_PATH, "native_loaded": _lib._lib is not None,
        "modules_from_dropin": sorted(m for m, mod in sys.modules.items()
                                      if getattr(mod, "__file__", None) and os.path.dirname(mod.__file__) == DROPIN),
        "neurons": np.asarray(res["neurons.dat"]).tolist(), "largest_losses": np.asarray(losses).tolist(),
        "total_training_hours": float(res["total_time.dat"]),
        "gpt_test_final_loss_median": float(np.median(_last_filled(res["gpt_test_losses.dat"]))),
        "pinn_test_final_loss_median": float(np.median(_last_filled(res["pinn_test_losses.dat"]))),
        "result_files_ok": sorted(k for k in res if k.endswith(".dat")),
    }
    # relative L2 error of the reduced model against the full PINNs on the test grid (what the paper's figures show)
    gs, ps = np.atleast_2d(res["gpt_test_soln.dat"]), np.atleast_2d(res["pinn_test_soln.dat"])
    err = np.linalg.norm(gs - ps, axis=0) / np.linalg.norm(ps, axis=0)
    summary["gpt_vs_pinn_rel_l2"] = {"median": float(np.median(err)), "max": float(np.max(err))}
    if rank == 0:
        json.dump(summary, open(os.path.join(out, "summary.json"), "w"), indent=1)
        print(json.dumps(summary))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
