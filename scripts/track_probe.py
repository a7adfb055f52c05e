"""Cost of the per-step hole tracking on large grids (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
for cells in (4096, 65536):
    res = {}
    for track in (False, True):
        cfg = SimConfig(n_steps=200, n_particles=2_000_000, n_cells=cells, periodic_particles=True, periodic_potential=False,
                        extra_storage_factor=1.0, track_hole=track, initial_transient_steps=2)
        sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
        for _ in range(10): sim.step()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(100): sim.step()
        ev1.record(); torch.cuda.synchronize()
        res[track] = ev0.elapsed_time(ev1) / 100
    print(cells, "ms/step without / with tracking", round(res[False], 4), round(res[True], 4), "tracking costs us", round((res[True] - res[False]) * 1e3, 1))
