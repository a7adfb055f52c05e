"""Sorts one 1e8-particle store on the 65536-cell grid (for ncu launch lists; development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
cfg = SimConfig(n_steps=4, n_particles=int(float(os.environ.get("PROBE_N", "1e8"))), n_cells=65536,
                periodic_particles=True, periodic_potential=False, extra_storage_factor=1.0, track_hole=False)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
sim.step()
for shift in (8, 0):
    sim.sort_particles(("electrons",), key_shift=shift)
    sim.step()
torch.cuda.synchronize()
