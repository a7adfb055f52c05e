"""BASELINE configs[4]: the long hole-tracking run (1e5 steps, script-default parameters) on one GPU, with the
diagnostics of SURVEY.md s.8d recorded every 100 steps, next to the ORACLE's history of the first 1e4 steps from
the same initial arrays (tests/golden/longrun_c1_n*.npz, made by tests/golden/make_longrun_golden.py from
oracle/loop.py over the compiled reference).  Writes a JSON history (kept under profiles/).

    python scripts/longrun_configs4.py [--n 1e6] [--steps 1e5] [--out gpurun_out/r2_configs4_history.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import run_gpu_hole_history  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=float, default=1e6)
ap.add_argument("--steps", type=float, default=1e5)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_configs4_history.json"))
args = ap.parse_args()
n, steps = int(args.n), int(args.steps)
hist, sim = run_gpu_hole_history(n, steps, progress=lambda k, s: print(f"step {k}: {s:.1f} s", flush=True))
out = {"config": {"n_per_species": n, "n_cells": 512, "steps": steps, "record_every": 100, "seed": 384,
                  "time_steps_immobile_ions": sim.cfg.time_steps_immobile_ions,
                  "note": "walls + injection, electron dimple, hole tracking every step; ions are frozen after step 30000 "
                          "as in the reference (infMagSim_script.py:929-939)"},
       "seconds": float(hist["seconds"]), "ms_per_step": 1e3 * float(hist["seconds"]) / steps,
       "status": int(hist["status"]),
       "gpu": {k: [float(x) for x in v] for k, v in hist.items() if k not in ("seconds", "status")}}
gold = os.path.join(ROOT, "tests", "golden", f"longrun_c1_n{n:.0e}.npz".replace("+0", ""))
if os.path.exists(gold):
    g = np.load(gold)
    m = min(len(g["step"]), len(hist["step"]))
    out["oracle_first_steps"] = {k: [float(x) for x in g[k][:m]] for k in ("step", "max_phi", "hole", "ke_i", "ke_e", "fe", "n_i", "n_e")}
    late = slice(m // 2, m)
    out["comparison_first_steps"] = {
        "steps_compared": int(g["step"][m - 1]) + 1,
        "ke_e_ratio_max_dev": float(np.abs(hist["ke_e"][:m] / g["ke_e"][:m] - 1).max()),
        "ke_i_ratio_max_dev": float(np.abs(hist["ke_i"][:m] / g["ke_i"][:m] - 1).max()),
        "n_e_ratio_max_dev": float(np.abs(hist["n_e"][:m] / g["n_e"][:m] - 1).max()),
        "max_phi_mean_gpu": float(hist["max_phi"][:m][late].mean()), "max_phi_mean_oracle": float(g["max_phi"][:m][late].mean()),
        "hole_median_abs_diff": float(np.median(np.abs(hist["hole"][:m][late] - g["hole"][:m][late])))}
ke = hist["ke_e"] + hist["ke_i"] + hist["fe"]
out["summary"] = {"total_energy_first": float(ke[0]), "total_energy_last": float(ke[-1]),
                  "total_energy_max_rel_dev_from_mean": float(np.abs(ke / ke.mean() - 1).max()),
                  "n_e_min": int(hist["n_e"].min()), "n_e_max": int(hist["n_e"].max()),
                  "n_i_min": int(hist["n_i"].min()), "n_i_max": int(hist["n_i"].max()),
                  "max_phi_last_mean": float(hist["max_phi"][-100:].mean()), "hole_last": float(hist["hole"][-1])}
os.makedirs(os.path.dirname(args.out), exist_ok=True)
with open(args.out, "w") as fh:
    json.dump(out, fh)
print(json.dumps({k: out[k] for k in ("config", "seconds", "ms_per_step", "status", "summary") if k in out}))
if "comparison_first_steps" in out:
    print(json.dumps(out["comparison_first_steps"]))
