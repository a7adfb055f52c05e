"""Step time and SM clock over a long run (development probe)."""
import sys, time, os, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from espic_hole_tracking_b200.simulation import SimConfig, Simulation
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cfg = SimConfig(n_steps=2000, n_particles=n, n_cells=4096, periodic_particles=True, periodic_potential=False, extra_storage_factor=1.0, track_hole=False)
sim = Simulation(cfg); sim.init_synthetic(); sim.initialize()
torch.cuda.synchronize()
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "250"], stdout=subprocess.PIPE, text=True)
lines = []
threading.Thread(target=lambda: [lines.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
for blk in range(12):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    sim.ctx.timing_enable(True)
    ev0.record()
    for _ in range(50): sim.step()
    ev1.record(); torch.cuda.synchronize()
    tm = sim.ctx.timing_read()
    st = sim.ctx.status()
    e = sim.energies() if blk % 3 == 0 else {}
    print("   mover mean ms", round(tm["mean_ms"], 4), "status", st, {k: round(v, 4) for k, v in e.items()})
    recent = [l for t, l in lines if t >= t0]
    print(f"block {blk}: {ev0.elapsed_time(ev1)/50:.3f} ms/step   smi: {recent[-1] if recent else None}")
proc.terminate()
