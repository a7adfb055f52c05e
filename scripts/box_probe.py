"""Moving-box controller in the loop: prints the tracked hole position and the box velocity (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(os.environ.get("PROBE_N", "1e6")))
cfg = SimConfig(n_steps=400, n_particles=n, n_cells=512, periodic_particles=False, track_hole=True,
                moving_box_simulation=True, hole_tracking_end_step=380, v_b_0=0.0)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
sim.run(400)
pos = sim.hole_relative_positions.cpu().numpy()
print("status", sim.status_word())
for k in range(80, 400, 20):
    print(k, "pos", pos[k], "box v", sim.ctrl.box[0, k], "a", sim.ctrl.box[1, k], "hole v", sim.ctrl.hole_velocities[k])
print("max potential", float(sim.potential.max()))
