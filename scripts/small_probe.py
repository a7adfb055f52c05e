"""Step rate at the reference's default scale (C1: ~1e6 particles, 512 cells, walls + injection)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
for n, cells, periodic in ((500_000, 512, False), (500_000, 512, True), (5_000_000, 512, False)):
    cfg = SimConfig(n_steps=3000, n_particles=n, n_cells=cells, periodic_particles=periodic, periodic_potential=False,
                    track_hole=True)
    sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
    for _ in range(100): sim.step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = 500
    for _ in range(K): sim.step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"n={n} cells={cells} periodic={periodic}: {1e3*(t2-t0)/K:.3f} ms/step (cpu enqueue {1e3*(t1-t0)/K:.3f}) "
          f"-> {2*n*K/(t2-t0):.3e} particle-steps/s")
