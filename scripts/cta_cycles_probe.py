"""Per-CTA cycles of the windowed mover in the production-like 65536-cell run (ESPIC_DEBUG_CTA=1): which CTAs are slow."""
import ctypes as C, os, sys
os.environ["ESPIC_DEBUG_CTA"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(os.environ.get("PROBE_N", "1e8")))
cfg = SimConfig(n_steps=400, n_particles=n, n_cells=65536, periodic_particles=False, track_hole=True, extra_storage_factor=1.1)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
buf = (C.c_longlong * 2048)()
def cycles():
    k = sim.ctx._lib.espic_debug_cta_cycles(sim.ctx.handle, buf, 2048)
    return np.array(buf[:k], dtype=np.int64)
nseg = 148 * int(os.environ.get("ESPIC_WINDOW_SEGMENTS", "2"))
for k in range(int(os.environ.get("PROBE_STEPS", "60"))):
    sim.step()
    if k in (5, 20, 27, 33, 47, 55):
        raw = cycles()
        for name, off in (("electrons", 0), ("ions", 512)):
            e = raw[off:off + min(nseg, 512)]
            order = np.argsort(-e)
            print(f"step {k} {name}: per-segment cycles min {e.min()} median {int(np.median(e))} max {e.max()} sum/148 {int(e.sum() / 148)}  "
                  f"slowest segments {order[:8].tolist()} -> {e[order[:8]].tolist()}", flush=True)
