"""Whole-step cost on the 65536-cell grid with the automatic sort policy (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(os.environ.get("PROBE_N", "1e8")))
periodic = os.environ.get("PROBE_WALLS", "0") != "1"
kw = {}
if "PROBE_SORT_E" in os.environ: kw["sort_interval_electrons"] = int(os.environ["PROBE_SORT_E"])
if "PROBE_SORT_I" in os.environ: kw["sort_interval_ions"] = int(os.environ["PROBE_SORT_I"])
cfg = SimConfig(n_steps=400, n_particles=n, n_cells=65536, periodic_particles=periodic, periodic_potential=False,
                extra_storage_factor=1.0 if periodic else 1.25, track_hole=False, **kw)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
print("sort intervals", sim.sort_intervals, "key_shift", sim.sort_key_shift)
for _ in range(10): sim.step()
steps = int(os.environ.get("PROBE_STEPS", "60"))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); ev0.record()
for _ in range(steps): sim.step()
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / steps
na = sim.n_active()
tot = na["ions"] + na["electrons"]
print("periodic" if periodic else "walls", f"ms/step {ms:.4f}  particle-steps/s {tot / ms * 1e3:.4e}  active {tot}  status {sim.status_word()}")
