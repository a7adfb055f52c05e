"""Literal host-array drop-in (move_particles_c with numpy arrays) at 1e8 slots per species for a few chunk sizes."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from espic_hole_tracking_b200._lib import lib
from oracle import workloads as wl
L = lib()
n, cells = 100_000_000, 4096
grid = wl.make_grid(cells); mask = np.zeros_like(grid); phi = wl.smooth_potential(grid)
rng = np.random.default_rng(0)
p = np.empty((2, n), np.float32)
p[0] = rng.uniform(-3, 3, n).astype(np.float32); p[1] = (rng.standard_normal(n) * 40).astype(np.float32)
stack = -np.ones(n, np.int32); den = np.zeros_like(grid)
P = lambda a: C.c_void_p(a.ctypes.data)
for chunk in [int(c) for c in os.environ.get("CHUNKS", "8388608").split(",")]:
    os.environ["ESPIC_HOST_CHUNK"] = str(chunk)
    top = C.c_int(-1)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        L.move_particles_c(P(grid), P(mask), P(phi), 8.75e-4, -1836.0, 5.0, n - 1, P(p), P(den), P(stack),
                           C.cast(C.byref(top), C.c_void_p), 0.0, 1, 1, n - 1, len(grid), n)
        dt = time.perf_counter() - t0
        if rep: best = min(best, dt)
    print(f"chunk {chunk:>9d} slots: {best * 1e3:7.2f} ms per call, {n / best:.3e} particle-steps/s, {1.6e9 / best / 1e9:.1f} GB/s both ways", flush=True)
os.environ["ESPIC_HOST_MODE"] = "zerocopy"
top = C.c_int(-1)
best = 1e9
ref = p.copy() if os.environ.get("CHECK") else None
for rep in range(4):
    t0 = time.perf_counter()
    L.move_particles_c(P(grid), P(mask), P(phi), 8.75e-4, -1836.0, 5.0, n - 1, P(p), P(den), P(stack),
                       C.cast(C.byref(top), C.c_void_p), 0.0, 1, 1, n - 1, len(grid), n)
    dt = time.perf_counter() - t0
    if rep: best = min(best, dt)
print(f"zero-copy: {best * 1e3:7.2f} ms per call, {n / best:.3e} particle-steps/s, {1.6e9 / best / 1e9:.1f} GB/s both ways  density sum {den.sum():.6f}", flush=True)
L.espic_host_release()
