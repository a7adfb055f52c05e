"""Times one mover variant alone on a store nothing leaves (development probe).
PROBE_CELLS, PROBE_WALLS=1, PROBE_SORTED=1, PROBE_DEAD=<fraction of parked slots>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import espic_hole_tracking_b200 as ops
from oracle import workloads as wl
n = 100_000_000
cells = int(os.environ.get("PROBE_CELLS", "4096"))
periodic = os.environ.get("PROBE_WALLS", "0") != "1"
grid = torch.from_numpy(wl.make_grid(cells)).cuda()
g = torch.Generator(device="cuda"); g.manual_seed(1)
p = torch.empty((2, n), device="cuda")
p[0] = (torch.rand(n, device="cuda", generator=g) * 5.0 - 2.5)
p[1] = torch.randn(n, device="cuda", generator=g) * float(os.environ.get("PROBE_VTH", "40.0"))
if os.environ.get("PROBE_SORTED") == "1":
    p[0] = torch.sort(p[0]).values
dead = float(os.environ.get("PROBE_DEAD", "0"))
if dead > 0:
    m = torch.rand(n, device="cuda", generator=g) < dead
    p[0][m] = 6.0
stack = torch.full((n,), -1, dtype=torch.int32, device="cuda")
top, last = ops.DeviceInt(-1), ops.DeviceInt(n - 1)
if os.environ.get("PROBE_SORTED") == "coarse":
    tl, ll = [-1], [n - 1]
    ops.sort_particles_by_cell(grid, p, None, stack, tl, ll, periodic_particles=periodic, key_shift=int(os.environ.get("PROBE_SHIFT", "8")))
    stack.fill_(-1)
    top, last = ops.DeviceInt(-1), ops.DeviceInt(ll[0])
den = torch.zeros(cells + 1, device="cuda")
zeros = torch.zeros(cells + 1, device="cuda")
ctx = ops.default_context()
dt = float(os.environ.get("PROBE_DT", "1e-7"))
for rep in range(2):
    ctx.timing_enable(True)
    for _ in range(6):
        ops.move_particles(grid, zeros, zeros, dt, -1836.0, 5.0, last, p, den, stack, top, 0.0, periodic_particles=periodic)
    torch.cuda.synchronize()
    r = ctx.timing_read(); ctx.timing_enable(False)
print("cells", cells, "periodic" if periodic else "walls", "sorted" if os.environ.get("PROBE_SORTED") else "random", "dead", dead,
      "mover ms/launch", round(r["mean_ms"], 4), "top", top[0])
