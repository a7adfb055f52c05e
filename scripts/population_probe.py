"""Population drift of wall-bounded runs in a few configurations (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
for label, kw in (("512 cells, no dimple", dict(n_cells=512, create_electron_dimple=False)),
                  ("512 cells, dimple", dict(n_cells=512)),
                  ("65536 cells, no dimple, auto sort", dict(n_cells=65536, create_electron_dimple=False)),
                  ("65536 cells, no dimple, never sort", dict(n_cells=65536, create_electron_dimple=False, sort_interval_ions=0, sort_interval_electrons=0)),
                  ("65536 cells, dimple, auto sort", dict(n_cells=65536))):
    cfg = SimConfig(n_steps=800, n_particles=1_000_000, periodic_particles=False, track_hole=False, extra_storage_factor=1.1, **kw)
    sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
    out = []
    for k in range(700):
        sim.step()
        if k in (99, 399, 699):
            a = sim.n_active(); out.append((k + 1, a["ions"], a["electrons"]))
    print(label, out, "status", sim.status_word(), "max phi", round(float(sim.potential.max()), 3), flush=True)
