"""BASELINE configs[3] as the reference would run it: 65536 cells, walls + injection, electron dimple,
hole tracking every step; whole-step time, population, energies, status (development probe)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(os.environ.get("PROBE_N", "1e8")))
steps = int(os.environ.get("PROBE_STEPS", "600"))
sort_e = os.environ.get("PROBE_SORT_E")
cfg = SimConfig(n_steps=steps + 200, n_particles=n, n_cells=65536, periodic_particles=False, track_hole=True,
                extra_storage_factor=1.1, sort_interval_electrons=int(sort_e) if sort_e else None)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
print("sort intervals", sim.sort_intervals, "key_shift", sim.sort_key_shift, flush=True)
if os.environ.get("PROBE_WARM_RUN") == "1":
    sim.run(100, check_every=0)
else:
    for _ in range(100): sim.step()
e0 = sim.energies(); a0 = sim.n_active()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l_before = sim.ctx.launch_count()
torch.cuda.synchronize(); ev0.record()
if os.environ.get("PROBE_RUN") == "1":
    sim.run(steps, check_every=0)   # stretches of steps from one host call, replayed from step graphs
else:
    for _ in range(steps): sim.step()
ev1.record(); torch.cuda.synchronize()
print("launches in the timed region", sim.ctx.launch_count() - l_before, flush=True)
ms = ev0.elapsed_time(ev1) / steps
e1 = sim.energies(); a1 = sim.n_active()
tot = a1["ions"] + a1["electrons"]
pos = sim.hole_relative_positions.cpu().numpy()[100:100 + steps]
print(json.dumps({"ms_per_step": ms, "particle_steps_per_s": tot / ms * 1e3, "active_start": a0, "active_end": a1,
                  "energies_start": e0, "energies_end": e1, "status": sim.status_word(),
                  "hole_position_last_10": [float(x) for x in pos[-10:]], "max_potential": float(sim.potential.max())}))
