"""Where does the e2e loop spend its time? (development probe, not part of the product)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cfg = SimConfig(n_steps=200, n_particles=n, n_cells=4096, periodic_particles=True, periodic_potential=False, extra_storage_factor=1.0, track_hole=False)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
for _ in range(3): sim.step()
torch.cuda.synchronize()
K = 20
def timed(fn, label):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"{label}: {(time.perf_counter()-t0)/K*1e3:.3f} ms/step")
def a():
    for _ in range(K): sim.step()
timed(a, "steps only")
def a_cpu():
    t0 = time.perf_counter()
    for _ in range(K): sim.step()
    print(f"  cpu enqueue time {(time.perf_counter()-t0)/K*1e3:.3f} ms/step")
timed(a_cpu, "steps only (2)")
out = [torch.zeros((3, sim.n_points)).pin_memory() for _ in range(2)]
ev = [torch.cuda.Event(), torch.cuda.Event()]
def b():
    for i in range(K):
        sim.step(); o = out[i & 1]
        o[0].copy_(sim.potential, non_blocking=True); o[1:3].copy_(sim.densities, non_blocking=True)
timed(b, "steps + D2H")
def c():
    for i in range(K):
        bb = i & 1
        sim.step(); o = out[bb]
        o[0].copy_(sim.potential, non_blocking=True); o[1:3].copy_(sim.densities, non_blocking=True)
        ev[bb].record()
        if i > 0:
            ev[1 - bb].synchronize(); float(out[1 - bb][0].max())
timed(c, "steps + D2H + lagged consume")

# detailed CPU-side timeline of the lagged loop
import collections
T = collections.defaultdict(float)
def d():
    for i in range(K):
        bb = i & 1
        t0 = time.perf_counter(); sim.step(); t1 = time.perf_counter()
        o = out[bb]
        o[0].copy_(sim.potential, non_blocking=True); o[1:3].copy_(sim.densities, non_blocking=True)
        ev[bb].record(); t2 = time.perf_counter()
        if i > 0:
            ev[1 - bb].synchronize(); t3 = time.perf_counter(); float(out[1 - bb][0].max()); t4 = time.perf_counter()
            T["wait"] += t3 - t2; T["max"] += t4 - t3
        T["step"] += t1 - t0; T["copies"] += t2 - t1
timed(d, "lagged loop (instrumented)")
print({k: round(v / K * 1e3, 4) for k, v in T.items()})
def e():
    for i in range(K):
        bb = i & 1
        sim.step(); o = out[bb]
        o[0].copy_(sim.potential, non_blocking=True); o[1:3].copy_(sim.densities, non_blocking=True)
        ev[bb].record()
        if i > 0:
            ev[1 - bb].synchronize()
timed(e, "lagged loop without max()")
import numpy as np
outn = [o.numpy() for o in out]
def f():
    for i in range(K):
        bb = i & 1
        sim.step(); o = out[bb]
        o[0].copy_(sim.potential, non_blocking=True); o[1:3].copy_(sim.densities, non_blocking=True)
        ev[bb].record()
        if i > 0:
            ev[1 - bb].synchronize(); float(outn[1 - bb][0].max())
timed(f, "lagged loop with numpy max")
