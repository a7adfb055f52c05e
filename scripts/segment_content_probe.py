"""What is in the slow segments of the windowed mover? (development probe)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(os.environ.get("PROBE_N", "1e8")))
cfg = SimConfig(n_steps=400, n_particles=n, n_cells=65536, periodic_particles=False, track_hole=True, extra_storage_factor=1.1)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
for k in range(21): sim.step()
e = sim.electrons
last = e.largest_index[0]; top = e.current_empty_slot[0]
print("largest_index", last, "stack top", top, "storage", e.storage, "active", sim.n_active())
n_slots = last + 1
x = e.particles[0, :n_slots]
tags = e.tags[0, :n_slots]
for seg in (0, 4, 60, 139, 140, 141, 142, 143, 147):
    a, b = seg * n_slots // 148, (seg + 1) * n_slots // 148
    xs = x[a:b].float().cpu().numpy(); tg = tags[a:b].cpu().numpy()
    live = (xs > -3) & (xs < 3)
    cells = ((xs[live] + 3) / 6 * 65536)
    q = np.percentile(cells, [0, 1, 25, 50, 75, 99, 100]).astype(int) if live.any() else []
    print(f"segment {seg}: slots [{a},{b}) live {live.mean():.3f} cell quantiles {list(q)} injected-since-start {np.mean(tg[live] > 0):.3f} "
          f"left-third {np.mean(cells < 21845):.3f} right-third {np.mean(cells > 43690):.3f}")
