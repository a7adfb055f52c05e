"""BASELINE.json's full single-GPU size (1e8 particles per species, 4096 cells): the oracle cannot
run this in seconds, so the checks are size-independent properties of the path:

* exact charge conservation -- the deposit is integer fixed point, so sum(density)*bg equals the
  number of live particles to float rounding of the final conversion, for any ordering;
* the deposit is independent of particle order: sorting the store changes no density bit;
* the entry fold of the periodic mover (x -> fl(fl(x - z_min) + z_min), infMagSim_c.c:148-153) can move
  a freshly wrapped position by one ulp, exactly as in the reference, and is idempotent from then on:
  a second dt=0 call leaves x bit-identical;
* every particle stays inside the domain, its cell index recomputed on the host matches the
  density's support;
* free-slot bookkeeping: live + stacked == storage, all stacked slots distinct and parked;
* the sort leaves keys non-decreasing and is a permutation (multiset of (x,v) preserved);
* a sample of 2e5 particles pushed by the oracle from the same inputs is bit-identical.
"""
import numpy as np
import pytest

from oracle import workloads as wl

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N = 100_000_000
N_CELLS = 4096


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 20e9:
        pytest.skip("needs > 20 GB of device memory")
    import espic_hole_tracking_b200 as ops
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    S = N + 4096
    p = torch.empty((2, S), dtype=torch.float32, device="cuda")
    p[0, :N] = (torch.rand(N, generator=gen, device="cuda", dtype=torch.float64) * 6 - 3).float()
    p[1, :N] = torch.randn(N, generator=gen, device="cuda") * 42.85
    p[0, N:] = 6.0
    p[1, N:] = 0.0
    grid = torch.from_numpy(wl.make_grid(N_CELLS)).cuda()
    phi = torch.from_numpy(wl.smooth_potential(wl.make_grid(N_CELLS), amplitude=1.0)).cuda()
    stack = torch.full((S,), -1, dtype=torch.int32, device="cuda")
    stack[:4096] = torch.arange(S - 1, N - 1, -1, dtype=torch.int32, device="cuda")
    return ops, dict(p=p, grid=grid, phi=phi, stack=stack, S=S)


@pytest.mark.parametrize("periodic", [True, False])
def test_full_size_properties(setup, periodic):
    ops, s = setup
    grid, phi, S = s["grid"], s["phi"], s["S"]
    p, stack = s["p"].clone(), s["stack"].clone()
    tags = torch.zeros((1, S), dtype=torch.int32, device="cuda")
    mask = torch.zeros_like(grid)
    top, last = ops.DeviceInt(4095), ops.DeviceInt(N - 1)
    bg = wl.background_density(N, 1, N_CELLS)
    den = torch.zeros(N_CELLS + 1, device="cuda")
    sample = torch.arange(0, N, 500, device="cuda")           # 2e5 slots followed through the oracle
    x0, v0 = p[0, sample].cpu().numpy(), p[1, sample].cpu().numpy()
    ops.move_particles(grid, mask, phi, wl.timestep(), -1836.0, bg, last, p, den, stack, top, 0.0,
                       periodic_particles=periodic)
    # --- oracle on the sample, bit for bit
    from oracle import cpu as ocpu
    k = ocpu.load("reference" if ocpu.have("reference") else "port")
    ps = np.stack([x0, v0]).astype(np.float32)
    n_s = ps.shape[1]
    st = -np.ones(n_s, np.int32)
    tp = [-1]
    k.move_particles(grid.cpu().numpy(), np.zeros(N_CELLS + 1, np.float32), phi.cpu().numpy(), wl.timestep(), -1836.0, bg,
                     [n_s - 1], ps, np.zeros(N_CELLS + 1, np.float32), st, tp, a_b=0.0, periodic_particles=periodic)
    got = torch.stack([p[0, sample], p[1, sample]]).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), ps.view(np.uint32))
    # --- bookkeeping and conservation
    live = p[0] < 5.9
    n_live = int(live.sum().item())
    t = top[0]
    assert n_live + (t + 1) == S
    if periodic:
        assert n_live == N
    else:
        assert N - 2_000_000 < n_live < N   # a thermal electron within v*dt of a wall leaves
    free = stack[:t + 1]
    assert int(torch.unique(free).numel()) == t + 1 and bool((p[0][free.long()] == 6.0).all())
    total = float(den.double().sum().item()) * bg
    assert abs(total - n_live) <= 2e-7 * n_live, (total, n_live)
    x = p[0][live]
    assert float(x.min()) >= -3.0 and float(x.max()) < 3.0 + (1e-6 if not periodic else 0.0)
    # --- order independence of the deposit + idempotent fold: dt = 0, no position update
    den1, den2 = torch.zeros_like(den), torch.zeros_like(den)
    ops.move_particles(grid, mask, phi, 0.0, -1836.0, bg, last, p, den1, stack, top, 0.0, update_position=False,
                       periodic_particles=periodic)
    x_before = p[0].clone()
    ops.move_particles(grid, mask, phi, 0.0, -1836.0, bg, last, p, den2, stack, top, 0.0, update_position=False,
                       periodic_particles=periodic)
    assert torch.equal(p[0], x_before)       # the fold is idempotent once applied
    assert torch.equal(den2, den1)           # same positions -> same integer sums -> same bits
    if not periodic:
        assert torch.equal(den1, den)        # no fold with walls: dt=0 changes nothing at all
    den = den1
    ops.sort_particles_by_cell(grid, p, tags, stack, top, last, periodic_particles=periodic)
    assert last[0] == n_live - 1 and top[0] == S - n_live - 1
    den3 = torch.zeros_like(den)
    ops.move_particles(grid, mask, phi, 0.0, -1836.0, bg, last, p, den3, stack, top, 0.0, update_position=False,
                       periodic_particles=periodic)
    assert torch.equal(den3, den)
    # dz = 6/4096 is exact in float32; a TENSOR divisor forces an IEEE division (torch turns division by a
    # Python scalar into a multiplication by its reciprocal, which rounds differently)
    dz_t = torch.full((1,), 6.0 / N_CELLS, dtype=torch.float32, device="cuda")
    cells = ((p[0, :n_live] - (-3.0)) / dz_t).int()
    if periodic:
        cells = torch.where(cells > N_CELLS - 1, torch.zeros_like(cells), cells)   # infMagSim_c.c:162-163
    assert bool((cells[1:] >= cells[:-1]).all())             # sorted by the mover's cell index
    assert ops.default_context().status() == 0


@pytest.mark.parametrize("periodic", [True, False])
def test_full_size_large_grid(setup, periodic):
    """The same store on the 65536-cell grid of BASELINE configs[3]: the mover keeps a window of the grid
    in shared memory per segment of the store (mover_win.cu).  On the random store most slots take
    its global fallback, on the coarse-sorted copy nearly all take the window; both must give the
    same density bit for bit (integer deposit), conserve charge exactly, leave the same multiset of
    particle states, and agree bit for bit with the oracle on a 2e5-slot sample of the SORTED store."""
    ops, s = setup
    S = s["S"]
    n_cells = 65536
    grid = torch.from_numpy(wl.make_grid(n_cells)).cuda()
    phi = torch.from_numpy(wl.smooth_potential(wl.make_grid(n_cells), amplitude=1.0)).cuda()
    mask = torch.zeros_like(grid)
    bg = wl.background_density(N, 1, n_cells)
    dt = wl.timestep()
    stores = {}
    for name in ("random", "sorted"):
        p, stack = s["p"].clone(), s["stack"].clone()
        top, last = ops.DeviceInt(4095), ops.DeviceInt(N - 1)
        if name == "sorted":
            ops.sort_particles_by_cell(grid, p, None, stack, top, last, periodic_particles=periodic, key_shift=8,
                                       lookahead=2.0 * dt)
        stores[name] = dict(p=p, stack=stack, top=top, last=last)
    st = stores["sorted"]
    sample = torch.arange(0, N, 500, device="cuda")
    x0, v0 = st["p"][0, sample].cpu().numpy(), st["p"][1, sample].cpu().numpy()
    dens = {}
    for name, d in stores.items():
        den = torch.zeros(n_cells + 1, device="cuda")
        for _ in range(2):
            ops.move_particles(grid, mask, phi, dt, -1836.0, bg, d["last"], d["p"], den, d["stack"], d["top"], 0.0,
                               periodic_particles=periodic)
        dens[name] = den
    assert torch.equal(dens["random"], dens["sorted"])
    # --- oracle on the sample of the sorted store, two steps, bit for bit
    from oracle import cpu as ocpu
    k = ocpu.load("reference" if ocpu.have("reference") else "port")
    ps = np.stack([x0, v0]).astype(np.float32)
    n_s = ps.shape[1]
    stk, tp = -np.ones(n_s, np.int32), [-1]
    for _ in range(2):
        k.move_particles(grid.cpu().numpy(), np.zeros(n_cells + 1, np.float32), phi.cpu().numpy(), dt, -1836.0, bg,
                         [n_s - 1], ps, np.zeros(n_cells + 1, np.float32), stk, tp, a_b=0.0, periodic_particles=periodic)
    got = torch.stack([st["p"][0, sample], st["p"][1, sample]]).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), ps.view(np.uint32))
    # --- conservation, bookkeeping, same multiset of states in both stores
    counts = {}
    for name, d in stores.items():
        live = d["p"][0] < 5.9
        counts[name] = int(live.sum().item())
        assert counts[name] + (d["top"][0] + 1) == S
    assert counts["random"] == counts["sorted"]
    total = float(dens["sorted"].double().sum().item()) * bg
    assert abs(total - counts["sorted"]) <= 2e-7 * counts["sorted"]
    keys = []
    for d in stores.values():
        live = d["p"][0] < 5.9
        packed = (d["p"][0][live].view(torch.int32).long() << 32) | (d["p"][1][live].view(torch.int32).long() & 0xffffffff)
        keys.append(torch.sort(packed).values)
    assert torch.equal(keys[0], keys[1])
    assert ops.default_context().status() == 0
