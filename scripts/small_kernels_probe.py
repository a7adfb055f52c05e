"""Small run (1e5 particles per species, 512 cells) for an ncu launch list: which kernels make up a ~60 us step.
PROBE_PERIODIC=1/0, PROBE_N, PROBE_CELLS; 50 steps, stepwise (one espic_pic_step per step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
periodic = os.environ.get("PROBE_PERIODIC", "1") == "1"
n = int(os.environ.get("PROBE_N", "100000"))
cells = int(os.environ.get("PROBE_CELLS", "512"))
cfg = SimConfig(n_steps=100, n_particles=n, n_cells=cells, periodic_particles=periodic, periodic_potential=False,
                track_hole=True, initial_transient_steps=5)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
sim.native_run = False
sim.run(50)
torch.cuda.synchronize()
print("done", sim.status_word())
