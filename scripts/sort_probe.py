"""Cost of the cell sort and its effect on the mover (development probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
for cells in (4096, 65536):
    cfg = SimConfig(n_steps=400, n_particles=100_000_000, n_cells=cells, periodic_particles=True, periodic_potential=False,
                    extra_storage_factor=1.0, track_hole=False)
    sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
    for _ in range(5): sim.step()
    def mover_ms(steps=5):
        sim.ctx.timing_enable(True)
        for _ in range(steps): sim.step()
        torch.cuda.synchronize()
        r = sim.ctx.timing_read(); sim.ctx.timing_enable(False)
        return r["mean_ms"]
    print(cells, "mover ms unsorted", round(mover_ms(), 4))
    torch.cuda.synchronize(); t0 = time.perf_counter(); sim.sort_particles(); torch.cuda.synchronize()
    print(cells, "sort both species ms", round((time.perf_counter() - t0) * 1e3, 2))
    t0 = time.perf_counter(); sim.sort_particles(); torch.cuda.synchronize()
    print(cells, "sort again (already sorted) ms", round((time.perf_counter() - t0) * 1e3, 2))
    for blk in range(4):
        print(cells, f"mover ms after sort, steps {blk*5}-{blk*5+4}:", round(mover_ms(), 4))
    del sim
    torch.cuda.empty_cache()
