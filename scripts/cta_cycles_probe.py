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
for k in range(int(os.environ.get("PROBE_STEPS", "60"))):
    sim.step()
    if k in (5, 20, 27, 33, 47, 55):
        # the last windowed launch of a step is the electron mover; run the ions alone once more to see them
        raw = cycles()
        e = raw[:148]
        seg = raw[160:160 + 8 * 148].reshape(148, 8)
        slow = np.argsort(-seg[:, 0])[:4]
        for sidx in list(slow) + [60]:
            print(f"   segment {sidx}: total {seg[sidx,0]} search {seg[sidx,1]} build {seg[sidx,2]} ranges {seg[sidx,3]} (warp0 {seg[sidx,5]}) flush {seg[sidx,4]} w_lo {seg[sidx,6]} rare {seg[sidx,7] >> 40} stray-gather {(seg[sidx,7] >> 20) & 0xfffff} stray-deposit {seg[sidx,7] & 0xfffff} warp-quads {2 * 676000 // 128}")
        order = np.argsort(-e)
        print(f"step {k} electrons: min {e.min()} median {int(np.median(e))} max {e.max()}  slowest CTAs {order[:6].tolist()} -> {e[order[:6]].tolist()}", flush=True)
