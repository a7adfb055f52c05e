"""Mover cost on the 65536-cell grid, step by step after a cell sort (development probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(os.environ.get("PROBE_N", "1e8")))
periodic = os.environ.get("PROBE_WALLS", "0") != "1"
cfg = SimConfig(n_steps=400, n_particles=n, n_cells=65536, periodic_particles=periodic, periodic_potential=False,
                extra_storage_factor=1.0 if periodic else 1.25, track_hole=False)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
for _ in range(3): sim.step()
def mover_ms(steps=1):
    sim.ctx.timing_enable(True)
    for _ in range(steps): sim.step()
    torch.cuda.synchronize()
    r = sim.ctx.timing_read(); sim.ctx.timing_enable(False)
    return r["mean_ms"]
print("periodic" if periodic else "walls", "mover ms/launch unsorted", round(mover_ms(3), 4), flush=True)
shift = int(os.environ.get("PROBE_SHIFT", "0"))
sim.sort_particles(key_shift=shift); torch.cuda.synchronize()
for _ in range(2): sim.step()
for name in ("ions", "electrons"):
    torch.cuda.synchronize(); t0 = time.perf_counter(); sim.sort_particles((name,), key_shift=shift); torch.cuda.synchronize()
    print("sort", name, "key_shift", shift, "ms", round((time.perf_counter() - t0) * 1e3, 3))
for k in range(int(os.environ.get("PROBE_STEPS", "24"))):
    print(f"step {k} after sort: mover ms/launch (mean of ions, electrons)", round(mover_ms(), 4), flush=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5): sim.step()
ev1.record(); torch.cuda.synchronize()
print("ms/step", ev0.elapsed_time(ev1) / 5)
