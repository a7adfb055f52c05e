#!/usr/bin/env python
"""Wall-clock us/step of small runs (BASELINE configs[0]/[4] sizes) through Simulation.run():
stepwise (one espic_pic_step host call per step from Python), espic_pic_run launch by launch, and
espic_pic_run with injection-free steps replayed from a CUDA graph.  One JSON line per case."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from espic_hole_tracking_b200.simulation import SimConfig, Simulation  # noqa: E402


def case(n, n_cells, periodic, mode, steps=3000, warm=300):
    # periodic particles with the Robin-end field solve, as in bench.py (the reference's periodic
    # solve is unstable at 4097 points)
    cfg = SimConfig(n_steps=steps + warm + 10, n_particles=n, n_cells=n_cells, periodic_particles=periodic,
                    periodic_potential=False, track_hole=True, initial_transient_steps=5)
    sim = Simulation(cfg)
    sim.init_synthetic()
    sim.initialize()
    sim.native_run = mode != "stepwise"
    sim.use_step_graph = mode == "graph"
    sim.run(warm)
    torch.cuda.synchronize()
    l0 = sim.ctx.launch_count()
    t0 = time.perf_counter()
    sim.run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = dict(n_particles=n, n_cells=n_cells, periodic=periodic, mode=mode, steps=steps,
               us_per_step=round(1e6 * dt / steps, 2), kernels_per_step=(sim.ctx.launch_count() - l0) / steps,
               particle_steps_per_s=2 * n * steps / dt, status=sim.status_word())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for n, cells in ((100_000, 512), (1_000_000, 4096)):
        for periodic in (True, False):
            for mode in ("stepwise", "native", "graph"):
                case(n, cells, periodic, mode)
