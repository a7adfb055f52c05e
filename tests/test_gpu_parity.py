"""GPU parity tests proper: the CUDA path (through the C ABI, via the operator layer) against
the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): particle state (x, v), particle-to-cell indexing, removal
order and free-slot stack are BIT-EXACT; density and potential agree within 1e-6 relative.
The reference's own float32 serial accumulation is itself off from the exact sum of its
weights by up to a few 1e-6 at thousands of particles per cell, so density
is checked twice: against the float64-exact sum of the reference's weights (tolerance 1e-6
of the peak density, typically 1e-7) and against the reference's float32 result (tolerance =
1e-6 plus the reference's own measured deviation from the exact sum).
"""
import math

import numpy as np
import pytest

from oracle import workloads as wl
from tests.helpers import assert_bits_equal, deposit_f64, to_dev

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REL = 1e-6  # north-star tolerance for densities and fields


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import espic_hole_tracking_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def cpu(ref_or_port):
    return ref_or_port


@pytest.fixture(scope="module")
def ref_or_port():
    from oracle import cpu as ocpu
    if ocpu.have("reference"):
        return ocpu.load("reference")
    if not ocpu.have("port"):
        ocpu.build()
    return ocpu.load("port")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def check_density(d_gpu, d_ref, grid, x_final, bg, periodic):
    exact, n_alive = deposit_f64(grid, x_final, bg, periodic)
    peak = exact.max()
    err_gpu = np.abs(d_gpu.astype(np.float64) - exact).max() / peak
    err_ref = np.abs(d_ref.astype(np.float64) - exact).max() / peak
    assert err_gpu <= REL, (err_gpu, err_ref)
    assert np.abs(d_gpu.astype(np.float64) - d_ref).max() / peak <= REL + err_ref
    return err_gpu, err_ref


@pytest.mark.parametrize("n_cells", [500, 512, 4096])
@pytest.mark.parametrize("periodic", [False, True])
def test_mover_multi_step(ops, cpu, n_cells, periodic):
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid, amplitude=2.0)
    dt = wl.timestep()
    n = 60_000
    _, sp = wl.make_plasma(n, n_cells, seed=n_cells + int(periodic), holes=0.02)
    sp.particles[0, :6] = np.array([-3.0, 3.0, -3.0000005, 2.9999995, 3.5, -7.0], np.float32)
    bg = wl.background_density(n, 1, n_cells)
    c = sp.copy()
    g = to_dev(sp, torch)
    g_grid, g_mask, g_phi = dev(grid), dev(mask), dev(phi)
    g_den = torch.full((len(grid),), 7.0, device="cuda")
    top_g, last_g = list(sp.current_empty_slot), list(sp.largest_index)
    for step in range(10):
        a_b = 0.37 * (step % 3)
        d_ref = np.full_like(grid, -7.0)
        cpu.move_particles(grid, mask, phi, dt, sp.charge_to_mass, bg, c.largest_index, c.particles, d_ref,
                           c.empty_slots, c.current_empty_slot, a_b=a_b, update_position=True,
                           periodic_particles=periodic)
        ops.move_particles(g_grid, g_mask, g_phi, dt, sp.charge_to_mass, bg, last_g, g["particles"], g_den,
                           g["empty_slots"], top_g, a_b, update_position=True, periodic_particles=periodic,
                           use_pure_c_version=True)
        assert top_g == c.current_empty_slot
        assert_bits_equal(g["particles"].cpu().numpy(), c.particles, f"particles step {step}")
        top = top_g[0]
        assert_bits_equal(g["empty_slots"].cpu().numpy()[:top + 1], c.empty_slots[:top + 1], "stack")
        check_density(g_den.cpu().numpy(), d_ref, grid, c.particles[0, :n], bg, periodic)
    assert ops.default_context().status() == 0
    if not periodic:
        assert top_g[0] > sp.current_empty_slot[0]


def test_mover_periodic_removal_order(ops, cpu):
    """Periodic runs remove a particle only when a step carries it past 0.99 * (2 z_max) (ref
    infMagSim_c.c:172): rare, but the freed slots must still reach the stack in ascending slot order.
    The periodic mover deals its 256-slot units round-robin to the warps and stages removals per
    unit; here ~2 % of the particles, spread over all units and over both the fast and the
    generic path, are flung out over three steps."""
    n_cells = 4096
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid, amplitude=1.0)
    dt = wl.timestep()
    n = 1_500_000      # 5859 units: more units than mover warps (4736), so some warps own two
    _, sp = wl.make_plasma(n, n_cells, seed=21, holes=0.01)
    rng = np.random.default_rng(5)
    flung = rng.choice(n, size=n // 50, replace=False)
    sp.particles[1, flung] = (rng.random(len(flung)) * 8000.0 + 2000.0).astype(np.float32)   # 1.7 .. 8.7 per step
    bg = wl.background_density(n, 1, n_cells)
    c = sp.copy()
    g = to_dev(sp, torch)
    top_g, last_g = list(sp.current_empty_slot), list(sp.largest_index)
    g_den = torch.zeros(len(grid), device="cuda")
    d_ref = np.zeros_like(grid)
    tops = [top_g[0]]
    for step in range(3):
        cpu.move_particles(grid, mask, phi, dt, sp.charge_to_mass, bg, c.largest_index, c.particles, d_ref,
                           c.empty_slots, c.current_empty_slot, a_b=0.0, periodic_particles=True)
        ops.move_particles(dev(grid), dev(mask), dev(phi), dt, sp.charge_to_mass, bg, last_g, g["particles"], g_den,
                           g["empty_slots"], top_g, 0.0, periodic_particles=True)
        assert top_g == c.current_empty_slot
        tops.append(top_g[0])
        assert_bits_equal(g["particles"].cpu().numpy(), c.particles, f"particles step {step}")
        assert_bits_equal(g["empty_slots"].cpu().numpy()[:top_g[0] + 1], c.empty_slots[:top_g[0] + 1], "stack")
        check_density(g_den.cpu().numpy(), d_ref, grid, c.particles[0, :n], bg, True)
    assert tops[1] > tops[0] + 1000 and tops[3] > tops[1]
    # a removal-free launch afterwards: the per-unit counters were left clean
    phi0 = np.zeros_like(grid)
    g["particles"][1].clamp_(-50.0, 50.0)
    c.particles[1] = np.clip(c.particles[1], -50.0, 50.0)
    cpu.move_particles(grid, mask, phi0, dt, sp.charge_to_mass, bg, c.largest_index, c.particles, d_ref,
                       c.empty_slots, c.current_empty_slot, a_b=0.0, periodic_particles=True)
    ops.move_particles(dev(grid), dev(mask), dev(phi0), dt, sp.charge_to_mass, bg, last_g, g["particles"], g_den,
                       g["empty_slots"], top_g, 0.0, periodic_particles=True)
    assert top_g == c.current_empty_slot and top_g[0] == tops[3]
    assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles, quiet step")


@pytest.mark.parametrize("periodic", [False, True])
def test_mover_scalar_path_and_device_scalars(ops, cpu, periodic):
    """Storage length not a multiple of 4 (no float4 path) + DeviceInt stack top / largest index."""
    n_cells = 512
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid)
    rng = np.random.default_rng(11)
    n, storage = 10_001, 12_347
    sp = wl.make_species(n, storage, rng, math.sqrt(1836.0), -1836.0, holes=0.01)
    c = sp.copy()
    g = to_dev(sp, torch)
    top = ops.DeviceInt(sp.current_empty_slot[0])
    last = ops.DeviceInt(sp.largest_index[0])
    g_den = torch.zeros(len(grid), device="cuda")
    d_ref = np.zeros_like(grid)
    for _ in range(3):
        cpu.move_particles(grid, mask, phi, 1e-3, -1836.0, 5.0, c.largest_index, c.particles, d_ref, c.empty_slots,
                           c.current_empty_slot, a_b=0.1, periodic_particles=periodic)
        ops.move_particles(dev(grid), dev(mask), dev(phi), 1e-3, -1836.0, 5.0, last, g["particles"], g_den,
                           g["empty_slots"], top, 0.1, periodic_particles=periodic)
    assert top[0] == c.current_empty_slot[0]
    assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
    assert_bits_equal(g["empty_slots"].cpu().numpy()[:top[0] + 1], c.empty_slots[:top[0] + 1], "stack")
    check_density(g_den.cpu().numpy(), d_ref, grid, c.particles[0, :n], 5.0, periodic)


def test_mover_object_and_no_position_update(ops, cpu):
    n_cells = 512
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    mask[250:262] = 1.0
    mask[249], mask[262] = 0.4, 0.7
    phi = wl.smooth_potential(grid)
    ions, _ = wl.make_plasma(40_000, n_cells, seed=5)
    c = ions.copy()
    g = to_dev(ions, torch)
    top_g, last_g = list(ions.current_empty_slot), list(ions.largest_index)
    g_den = torch.zeros(len(grid), device="cuda")
    d_ref = np.zeros_like(grid)
    for upd in (False, True):
        cpu.move_particles(grid, mask, phi, 1e-2, 1.0, 3.0, c.largest_index, c.particles, d_ref, c.empty_slots,
                           c.current_empty_slot, a_b=0.0, update_position=upd)
        ops.move_particles(dev(grid), dev(mask), dev(phi), 1e-2, 1.0, 3.0, last_g, g["particles"], g_den,
                           g["empty_slots"], top_g, 0.0, update_position=upd)
        assert top_g == c.current_empty_slot
        assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
        assert_bits_equal(g["empty_slots"].cpu().numpy()[:top_g[0] + 1], c.empty_slots[:top_g[0] + 1], "stack")
        np.testing.assert_allclose(g_den.cpu().numpy(), d_ref, rtol=0, atol=3e-6 * d_ref.max())
    assert top_g[0] > ions.current_empty_slot[0]


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("n_cells", [65536, 24576, 20000])
def test_mover_large_grid_path(ops, cpu, periodic, n_cells):
    """65537 grid points: the tables do not fit in shared memory (C4 of BASELINE.json).  24576 / 20000 cells:
    too large for the whole-grid tables (~19k points) but not larger than the window of the windowed mover,
    which then covers the grid completely (no global-memory path, no sorting needed)."""
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid, amplitude=1.0)
    n = 300_000
    _, sp = wl.make_plasma(n, n_cells, seed=3, holes=0.01)
    bg = wl.background_density(n, 1, n_cells)
    c = sp.copy()
    g = to_dev(sp, torch)
    top_g, last_g = list(sp.current_empty_slot), list(sp.largest_index)
    g_den = torch.zeros(len(grid), device="cuda")
    d_ref = np.zeros_like(grid)
    for _ in range(3):
        cpu.move_particles(grid, mask, phi, wl.timestep(), sp.charge_to_mass, bg, c.largest_index, c.particles,
                           d_ref, c.empty_slots, c.current_empty_slot, a_b=0.0, periodic_particles=periodic)
        ops.move_particles(dev(grid), dev(mask), dev(phi), wl.timestep(), sp.charge_to_mass, bg, last_g,
                           g["particles"], g_den, g["empty_slots"], top_g, 0.0, periodic_particles=periodic)
    assert top_g == c.current_empty_slot
    assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
    assert_bits_equal(g["empty_slots"].cpu().numpy()[:top_g[0] + 1], c.empty_slots[:top_g[0] + 1], "stack")
    check_density(g_den.cpu().numpy(), d_ref, grid, c.particles[0, :n], bg, periodic)


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("domain", [(-7.3, 11.9, 1000), (0.25, 50.25, 777), (-40.0, -1.5, 4096)])
def test_mover_other_domains(ops, cpu, domain, periodic):
    """Nothing in the kernels assumes the script's [-3, 3] box: z_min not a multiple of dz (the
    reference's fmodf(x, dz) weights are then offset against the cells), one-signed domains."""
    z_lo, z_hi, n_cells = domain
    grid = np.linspace(z_lo, z_hi, n_cells + 1).astype(np.float32)
    mask = np.zeros_like(grid)
    rng = np.random.default_rng(n_cells)
    phi = (0.7 * np.sin(2 * np.pi * (grid - grid[0]) / (grid[-1] - grid[0]) * 3)).astype(np.float32)
    n, S = 50_000, 60_000
    L = float(grid[-1]) - float(grid[0])
    p = np.empty((2, S), np.float32)
    p[0, :n] = (float(grid[0]) + rng.random(n) * L).astype(np.float32)
    p[1, :n] = (rng.standard_normal(n) * 30.0).astype(np.float32)
    p[0, n:], p[1, n:] = np.float32(2 * grid[-1]), 0.0
    if grid[-1] <= 0:       # the reference's parked marker 2*z_max lies inside a negative domain: no dead slots then
        S = n
        p = np.ascontiguousarray(p[:, :n])
    stack = -np.ones(S, np.int32)
    stack[:S - n] = np.arange(S - 1, n - 1, -1, dtype=np.int32)
    sp = wl.Species(p, np.zeros((1, S), np.int32), stack, [S - n - 1], [n - 1], -1836.0, 30.0)
    bg, dt = 5.0, np.float32(2e-3)
    c = sp.copy()
    g = to_dev(sp, torch)
    top_g, last_g = list(sp.current_empty_slot), list(sp.largest_index)
    g_den = torch.zeros(len(grid), device="cuda")
    d_ref = np.zeros_like(grid)
    for _ in range(3):
        cpu.move_particles(grid, mask, phi, dt, sp.charge_to_mass, bg, c.largest_index, c.particles, d_ref,
                           c.empty_slots, c.current_empty_slot, a_b=0.0, periodic_particles=periodic)
        ops.move_particles(dev(grid), dev(mask), dev(phi), dt, sp.charge_to_mass, bg, last_g, g["particles"], g_den,
                           g["empty_slots"], top_g, 0.0, periodic_particles=periodic)
        assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
        assert top_g == c.current_empty_slot
        assert_bits_equal(g["empty_slots"].cpu().numpy()[:top_g[0] + 1], c.empty_slots[:top_g[0] + 1], "stack")
        np.testing.assert_allclose(g_den.cpu().numpy(), d_ref, rtol=0, atol=5e-6 * d_ref.max())


@pytest.mark.parametrize("n_cells", [19, 37, 137])
def test_mover_cell_index_overflow_just_below_zmax(ops, cpu, n_cells):
    """dz = fl(span/n_cells) is rounded, so on some grids the last float below z_max already divides
    to n_cells; the reference's periodic branch then takes cell 0 (ref infMagSim_c.c:161-165,
    185-189).  Particles parked on those floats, before and after the push, must follow it."""
    grid = wl.make_grid(n_cells)
    assert int(np.float32(np.nextafter(np.float32(6), np.float32(0))) / np.float32(np.float32(6) / np.float32(n_cells))) == n_cells
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid, amplitude=0.5)
    n, S = 4096, 4096
    rng = np.random.default_rng(n_cells)
    sp = wl.make_species(n, S, rng, 1.0, 1.0)
    top_ulps = np.float32(3.0) - np.arange(1, 9, dtype=np.float32) * np.float32(2.0 ** -22)   # the 8 floats below z_max
    sp.particles[0, :1024] = np.resize(top_ulps, 1024)
    sp.particles[1, :1024] = np.resize(np.array([0.0, 1e-9, -1e-9, 1e-4, -1e-4], np.float32), 1024)
    sp.particles[0, 1024:2048] = np.float32(-3.0) + rng.random(1024).astype(np.float32) * np.float32(1e-3)
    sp.particles[1, 1024:2048] = np.float32(-1.0)            # cross the seam downwards: land just below z_max
    bg, dt = 3.0, np.float32(1e-3)
    c = sp.copy()
    g = to_dev(sp, torch)
    top_g, last_g = list(sp.current_empty_slot), list(sp.largest_index)
    g_den = torch.zeros(len(grid), device="cuda")
    d_ref = np.zeros_like(grid)
    for _ in range(3):
        cpu.move_particles(grid, mask, phi, dt, 0.01, bg, c.largest_index, c.particles, d_ref, c.empty_slots,
                           c.current_empty_slot, a_b=0.0, periodic_particles=True)
        ops.move_particles(dev(grid), dev(mask), dev(phi), dt, 0.01, bg, last_g, g["particles"], g_den,
                           g["empty_slots"], top_g, 0.0, periodic_particles=True)
        assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
        np.testing.assert_allclose(g_den.cpu().numpy(), d_ref, rtol=0, atol=3e-6 * d_ref.max())
    assert top_g == c.current_empty_slot


@pytest.mark.parametrize("periodic", [False, True])
def test_mover_window_on_sorted_store(ops, cpu, periodic):
    """65537 grid points, store coarse-sorted first: the large-grid mover then serves nearly every
    slot from its per-segment shared-memory window (mover_win.cu), the fast electrons that outrun
    the window and -- periodic -- the segments straddling the seam included.  Same bars."""
    n_cells = 65536
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid, amplitude=1.0)
    n = 400_000
    _, sp = wl.make_plasma(n, n_cells, seed=11, holes=0.01)
    bg = wl.background_density(n, 1, n_cells)
    g = to_dev(sp, torch)
    top_g, last_g = list(sp.current_empty_slot), list(sp.largest_index)
    ops.sort_particles_by_cell(dev(grid), g["particles"], g["tags"], g["empty_slots"], top_g, last_g,
                               periodic_particles=periodic, key_shift=8)
    c = sp.copy()   # the oracle starts from the sorted store
    c.particles[:] = g["particles"].cpu().numpy()
    c.empty_slots[:] = g["empty_slots"].cpu().numpy()
    c.current_empty_slot[0], c.largest_index[0] = top_g[0], last_g[0]
    n_live = last_g[0] + 1
    assert np.all(np.diff((c.particles[0, :n_live] - grid[0]) // ((grid[-1] - grid[0]) / n_cells * 256)) >= 0)
    g_den = torch.zeros(len(grid), device="cuda")
    d_ref = np.zeros_like(grid)
    dt = 4 * wl.timestep()   # fast electrons cover thousands of cells in the 4 steps
    for _ in range(4):
        cpu.move_particles(grid, mask, phi, dt, sp.charge_to_mass, bg, c.largest_index, c.particles,
                           d_ref, c.empty_slots, c.current_empty_slot, a_b=0.0, periodic_particles=periodic)
        ops.move_particles(dev(grid), dev(mask), dev(phi), dt, sp.charge_to_mass, bg, last_g,
                           g["particles"], g_den, g["empty_slots"], top_g, 0.0, periodic_particles=periodic)
    assert top_g == c.current_empty_slot
    assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
    assert_bits_equal(g["empty_slots"].cpu().numpy()[:top_g[0] + 1], c.empty_slots[:top_g[0] + 1], "stack")
    check_density(g_den.cpu().numpy(), d_ref, grid, c.particles[0, :n_live], bg, periodic)


@pytest.mark.parametrize("n_cells", [500, 512, 1000, 4096, 65536])
def test_fast_quotient_is_ieee_division_exhaustively(ops, n_cells):
    """The mover replaces the division in (x - z_min)/dz by a 3-operation sequence; this checks
    it against the division instruction for EVERY float in [0, z_max - z_min] on the device."""
    import ctypes as C
    ctx = ops.default_context()
    grid = dev(wl.make_grid(n_cells))
    counts = (C.c_ulonglong * 2)()
    ctx.check(ctx._lib.espic_selftest_quotient(ctx.handle, C.c_void_p(grid.data_ptr()), len(grid), counts))
    assert counts[0] == 0 and counts[1] == 0, list(counts)


def test_accumulate_density_and_initialize_mover(ops, cpu):
    n_cells = 512
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid)
    ions, _ = wl.make_plasma(30_000, n_cells, seed=8)
    bg = wl.background_density(30_000, 1, n_cells)
    c = ions.copy()
    g = to_dev(ions, torch)
    top_g, last_g = list(ions.current_empty_slot), list(ions.largest_index)
    d_ref = np.zeros_like(grid)
    g_den = torch.zeros(len(grid), device="cuda")
    cpu.accumulate_density(grid, mask, bg, c.largest_index, c.particles, d_ref, c.empty_slots, c.current_empty_slot)
    ops.accumulate_density(dev(grid), dev(mask), bg, last_g, g["particles"], g_den, g["empty_slots"], top_g,
                           use_pure_c_version=True)
    check_density(g_den.cpu().numpy(), d_ref, grid, c.particles[0, :30_000], bg, False)
    assert abs(float(g_den.sum().item()) * bg - 30_000) < 0.05  # charge conservation
    cpu.initialize_mover(grid, mask, phi, wl.timestep(), 1.0, c.largest_index, c.particles, c.empty_slots,
                         c.current_empty_slot, v_b=0.123456789, a_b=0.01)
    ops.initialize_mover(dev(grid), dev(mask), dev(phi), wl.timestep(), 1.0, last_g, g["particles"],
                         g["empty_slots"], top_g, v_b=0.123456789, a_b=0.01, use_pure_c_version=True)
    assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles after initialize_mover")


# 9 .. 1024 points: one CTA; 1500: cluster of 2; 3000: 4; 4097 and up: 8 (1, 3 and 9 rows per thread)
@pytest.mark.parametrize("n_points", [9, 513, 1024, 1025, 1501, 3000, 3001, 4097, 8193, 20001, 65537])
@pytest.mark.parametrize("periodic", [False, True])
def test_poisson(ops, cpu, n_points, periodic):
    if periodic and n_points % 2 == 0:
        # the reference reads v[n_points] past its malloc'd arrays in this branch;
        # with an even point count that lands in allocator slack holding stale heap bytes, not in the
        # next chunk's size field (a denormal, i.e. 0 in the product): its output is not reproducible
        pytest.skip("reference result depends on uninitialised heap bytes for even n_points")
    grid = wl.make_grid(n_points - 1)
    rng = np.random.default_rng(n_points)
    charge = (0.05 * rng.standard_normal(n_points)).astype(np.float32)
    mask = np.zeros_like(grid)
    p_ref = np.zeros_like(grid)
    cpu.poisson_solve(grid, mask, charge, wl.DEBYE_LENGTH, p_ref, periodic_potential=periodic)
    p_gpu = torch.zeros(n_points, device="cuda")
    ops.poisson_solve(dev(grid), dev(mask), dev(charge), wl.DEBYE_LENGTH, p_gpu, periodic_potential=periodic,
                      use_pure_c_version=True)
    scale = np.abs(p_ref).max()
    assert scale > 0
    first = p_gpu.cpu().numpy().copy()
    err = np.abs(first.astype(np.float64) - p_ref).max() / scale
    assert err <= REL, err
    # the second call finds the factorisation in the workspace: same bits
    p_gpu.zero_()
    ops.poisson_solve(dev(grid), dev(mask), dev(charge), wl.DEBYE_LENGTH, p_gpu, periodic_potential=periodic)
    assert_bits_equal(p_gpu.cpu().numpy(), first, "cached vs uncached solve")
    # and a different right-hand side on the cached factorisation is still right
    charge2 = (0.05 * rng.standard_normal(n_points)).astype(np.float32)
    cpu.poisson_solve(grid, mask, charge2, wl.DEBYE_LENGTH, p_ref, periodic_potential=periodic)
    ops.poisson_solve(dev(grid), dev(mask), dev(charge2), wl.DEBYE_LENGTH, p_gpu, periodic_potential=periodic)
    err = np.abs(p_gpu.cpu().numpy().astype(np.float64) - p_ref).max() / np.abs(p_ref).max()
    assert err <= REL, err


@pytest.mark.parametrize("n_points", [513, 3001, 9001])
def test_poisson_object_and_boltzmann(ops, cpu, n_points):
    grid = wl.make_grid(n_points - 1)
    rng = np.random.default_rng(3)
    charge = (1.0 + 0.05 * rng.standard_normal(n_points)).astype(np.float32)
    mask = np.zeros_like(grid)
    k = n_points // 513      # the object straddles run and CTA boundaries on the larger grids
    mask[200 * k:240 * k] = 1.0
    mask[200 * k - 1], mask[240 * k] = 0.3, 0.6
    mask[300 * k] = 0.5
    for transparency in (0.0, 0.35, 1.0):
        for boltz in (False, True):
            p_ref = (0.1 * np.sin(grid)).astype(np.float32)
            p_gpu = dev(p_ref)
            cpu.poisson_solve(grid, mask, charge, 0.25, p_ref, object_potential=-3.0,
                              object_transparency=transparency, boltzmann_electrons=boltz)
            ops.poisson_solve(dev(grid), dev(mask), dev(charge), 0.25, p_gpu, object_potential=-3.0,
                              object_transparency=transparency, boltzmann_electrons=boltz)
            err = np.abs(p_gpu.cpu().numpy().astype(np.float64) - p_ref).max() / np.abs(p_ref).max()
            assert err <= REL, (transparency, boltz, err)


def test_prepare_charge(ops):
    rng = np.random.default_rng(0)
    for periodic in (False, True):
        ni = rng.random(513).astype(np.float32)
        ne = rng.random(513).astype(np.float32)
        gi, ge, gc = dev(ni), dev(ne), torch.zeros(513, device="cuda")
        ops.prepare_charge(gi, ge, gc, periodic_particles=periodic)
        for d in (ni, ne):  # infMagSim_script.py:764-769
            if periodic:
                d[0] += d[-1]
                d[-1] = d[0]
            else:
                d[0] *= 2.0
                d[-1] *= 2.0
        assert_bits_equal(gi.cpu().numpy(), ni)
        assert_bits_equal(ge.cpu().numpy(), ne)
        assert_bits_equal(gc.cpu().numpy(), ni - ne)


@pytest.mark.parametrize("periodic,periodic_potential", [(False, False), (True, True), (True, False)])
@pytest.mark.parametrize("n_points", [513, 4097, 65537])
def test_field_solve_fused_equals_separate_calls(ops, periodic, periodic_potential, n_points):
    rng = np.random.default_rng(5)
    grid = dev(wl.make_grid(n_points - 1))
    mask = torch.zeros_like(grid)
    ni = (1 + 0.05 * rng.standard_normal(n_points)).astype(np.float32)
    ne = (1 + 0.05 * rng.standard_normal(n_points)).astype(np.float32)
    a_i, a_e, a_c, a_p = dev(ni), dev(ne), torch.zeros_like(grid), torch.zeros_like(grid)
    b_i, b_e, b_c, b_p = dev(ni), dev(ne), torch.zeros_like(grid), torch.zeros_like(grid)
    ops.prepare_charge(a_i, a_e, a_c, periodic_particles=periodic, fix_ions=True, fix_electrons=True)
    ops.poisson_solve(grid, mask, a_c, wl.DEBYE_LENGTH, a_p, object_potential=-3.,
                      periodic_potential=periodic_potential)
    ops.field_solve(grid, mask, b_i, b_e, b_c, wl.DEBYE_LENGTH, b_p, object_potential=-3.,
                    periodic_particles=periodic, periodic_potential=periodic_potential)
    for x, y in ((a_i, b_i), (a_e, b_e), (a_c, b_c), (a_p, b_p)):
        assert_bits_equal(x.cpu().numpy(), y.cpu().numpy())


def test_injection_given_inputs(ops, cpu):
    """Injection body with both random inputs supplied: slot assignment, tags, x, v bit-exact."""
    from oracle import cpu as ocpu
    port = ocpu.load("port") if ocpu.have("port") else None
    if port is None:
        ocpu.build()
        port = ocpu.load("port")
    for n_cells in (500, 4096):
        grid = wl.make_grid(n_cells)
        ions, _ = wl.make_plasma(20_000, n_cells, seed=9)
        c = ions.copy()
        g = to_dev(ions, torch)
        d0 = np.random.default_rng(1).random(len(grid)).astype(np.float32)
        d_ref, g_den = d0.copy(), dev(d0)
        cpu.seedgenrand_c(384)
        n_inj = 3_000
        vel = cpu.draw_velocities(n_inj, 1.0, 0.3)
        uni = np.random.default_rng(2).random(n_inj).astype(np.float32)
        bg = 40.0
        port.inject_given(n_inj, grid, 0.01, bg, vel, uni, 0.2, 17, c.tags, c.particles, c.empty_slots,
                          c.current_empty_slot, c.largest_index, d_ref)
        top_g, last_g = list(ions.current_empty_slot), list(ions.largest_index)
        ops.inject_particles_given(n_inj, dev(grid), 0.01, bg, dev(vel), dev(uni), 0.2, 17, g["tags"],
                                   g["particles"], g["empty_slots"], top_g, last_g, g_den)
        assert top_g == c.current_empty_slot and last_g == c.largest_index
        assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
        assert_bits_equal(g["empty_slots"].cpu().numpy(), c.empty_slots, "stack")
        assert_bits_equal(g["tags"].cpu().numpy(), c.tags, "tags")
        np.testing.assert_allclose(g_den.cpu().numpy(), d_ref, rtol=0, atol=2e-6 * d_ref.max())
    assert ops.default_context().status() == 0


def test_injection_bad_index_replays_reference_order(ops):
    """A particle whose partial step carries it past the far wall keeps its slot on the stack
    (infMagSim_cython.py:272-274); the kernel falls back to the sequential replay."""
    from oracle import cpu as ocpu
    if not ocpu.have("port"):
        ocpu.build()
    port = ocpu.load("port")
    grid = wl.make_grid(64)
    ions, _ = wl.make_plasma(100, 64, seed=1)
    c = ions.copy()
    g = to_dev(ions, torch)
    vel = np.array([1.0, 900.0, 2.0, -3.0, 800.0], np.float32)  # 900*0.01*0.9 > 6: out of the domain
    uni = np.array([0.5, 0.9, 0.1, 0.7, 0.95], np.float32)
    d_ref = np.zeros_like(grid)
    g_den = torch.zeros(len(grid), device="cuda")
    port.inject_given(5, grid, 0.01, 1.0, vel, uni, 0.0, 3, c.tags, c.particles, c.empty_slots,
                      c.current_empty_slot, c.largest_index, d_ref)
    top_g, last_g = list(ions.current_empty_slot), list(ions.largest_index)
    ops.inject_particles_given(5, dev(grid), 0.01, 1.0, dev(vel), dev(uni), 0.0, 3, g["tags"], g["particles"],
                               g["empty_slots"], top_g, last_g, g_den)
    assert top_g == c.current_empty_slot == [ions.current_empty_slot[0] - 3]
    assert_bits_equal(g["particles"].cpu().numpy(), c.particles, "particles")
    assert_bits_equal(g["empty_slots"].cpu().numpy(), c.empty_slots, "stack")
    from espic_hole_tracking_b200._lib import lib  # noqa: F401
    assert ops.default_context().status() & 8  # ESPIC_ST_BAD_INJECT_INDEX


@pytest.mark.parametrize("v_th,v_d", [(42.85, 0.0), (1.0, 0.8), (42.85, -30.0)])
def test_device_sampler_statistics(ops, cpu, v_th, v_d):
    """Same distribution as draw_velocities_c (different stream): moments and a two-sample KS
    distance against the oracle's MT19937-driven sampler."""
    n = 400_000
    ops.seedgenrand_c(384)
    v_gpu = ops.draw_velocities(n, v_th, v_d).cpu().numpy()
    v_gpu2 = ops.draw_velocities(n, v_th, v_d).cpu().numpy()
    assert not np.array_equal(v_gpu, v_gpu2)  # the stream advances between calls
    ops.seedgenrand_c(384)
    assert np.array_equal(ops.draw_velocities(n, v_th, v_d).cpu().numpy(), v_gpu)  # reseeding replays it
    cpu.seedgenrand_c(384)
    v_cpu = cpu.draw_velocities(n, v_th, v_d)
    assert abs((v_gpu > 0).mean() - (v_cpu > 0).mean()) < 4e-3
    for sign in (1, -1):
        a, b = np.sort(v_gpu[sign * v_gpu > 0]), np.sort(v_cpu[sign * v_cpu > 0])
        assert abs(a.mean() - b.mean()) < 0.01 * v_th
        assert abs(a.std() - b.std()) < 0.01 * v_th
        both = np.concatenate([a, b])
        ks = np.abs(np.searchsorted(a, both, side="right") / len(a)
                    - np.searchsorted(b, both, side="right") / len(b)).max()
        assert ks < 1.95 * math.sqrt((len(a) + len(b)) / (len(a) * len(b)))  # alpha ~ 1e-3
    u = ops.operators.device_uniforms(1_000_000).cpu().numpy()
    assert 0.0 <= u.min() and u.max() <= 1.0 and abs(u.mean() - 0.5) < 1e-3 and abs(u.var() - 1 / 12) < 1e-3


def test_hole_tracking_filter_and_histogram(ops, cpu):
    n_cells = 512
    grid = wl.make_grid(n_cells)
    dz = 6.0 / n_cells
    rng = np.random.default_rng(4)
    z = grid.astype(np.float64)
    phi = (0.3 * np.exp(-((z - 0.7) / 0.5) ** 2) + 0.002 * rng.standard_normal(len(grid))).astype(np.float32)
    fine = np.linspace(-3, 3, 1001).astype(np.float32)
    sig = np.gradient(phi, dz).astype(np.float32)
    r_ref = cpu.electric_field_filter(grid, sig, 0.5, fine)
    r_gpu = ops.electric_field_filter(dev(grid), dev(sig), 0.5, dev(fine)).cpu().numpy()
    np.testing.assert_allclose(r_gpu, r_ref, rtol=0, atol=2e-5 * np.abs(r_ref).max())
    pos = ops.hole_position_tracking(dev(phi), dz, 0.5, dev(grid), dev(fine), 6.0 / 1000).cpu().numpy()[0]
    assert pos == fine[np.argmax(r_ref)]
    x = rng.uniform(-3.2, 3.2, 200_000).astype(np.float32)
    v = (4 * rng.standard_normal(200_000)).astype(np.float32)
    tags = (rng.random(200_000) < 0.3).astype(np.int32)
    h_ref = cpu.histogram2d_uniform_grid(x, v, -3.0, 3.0, -8.0, 8.0, 100, 100)[0]
    h_gpu, xc, yc = ops.histogram2d_uniform_grid(dev(x), dev(v), -3.0, 3.0, -8.0, 8.0, 100, 100)
    np.testing.assert_array_equal(h_gpu.cpu().numpy(), h_ref)
    h_ref0 = cpu.histogram2d_uniform_grid(x[tags == 0], v[tags == 0], -3.0, 3.0, -8.0, 8.0, 100, 100)[0]
    h_gpu0 = ops.histogram2d_uniform_grid(dev(x), dev(v), -3.0, 3.0, -8.0, 8.0, 100, 100, tags=dev(tags))[0]
    np.testing.assert_array_equal(h_gpu0.cpu().numpy(), h_ref0)


def _tracking_potential(kind, grid, rng):
    z = grid.astype(np.float64)
    if kind == "hole":
        phi = 0.3 * np.exp(-((z - 0.7) / 0.5) ** 2) + 0.002 * rng.standard_normal(len(z))
    elif kind == "near_wall":
        phi = 0.5 * np.exp(-((z + 2.6) / 0.3) ** 2) + 0.001 * rng.standard_normal(len(z))
    elif kind == "two_holes":   # nearly equal depths: the maximum of the response is contested
        phi = 0.30 * np.exp(-((z + 1.1) / 0.45) ** 2) + 0.3003 * np.exp(-((z - 1.4) / 0.45) ** 2)
    elif kind == "noisy":       # particle-noise level of a coarse run
        phi = 0.2 * np.exp(-(z / 0.6) ** 2) + 0.03 * rng.standard_normal(len(z))
    else:                       # broad hole on a sinusoidal background
        phi = 0.4 * np.exp(-((z - 0.2) / 1.2) ** 2) + 0.05 * np.sin(2 * np.pi * z / 1.7) + 0.002 * rng.standard_normal(len(z))
    return phi.astype(np.float32)


@pytest.mark.parametrize("n_points", [513, 4097, 65537])
@pytest.mark.parametrize("kind", ["hole", "near_wall", "two_holes", "noisy", "broad"])
def test_hole_tracking_argmax_at_baseline_grids(ops, cpu, n_points, kind):
    """Hole position (ref infMagSim_c.c:668-683 + infMagSim_script.py:746-753) at the BASELINE grids: the same
    fine-mesh INDEX as the compiled reference.  The reference accumulates 65537 float additions per mesh point;
    where that noise makes two neighbouring points tie, +-1 index is accepted only if the float64-exact response
    gap between the two points is below the reference's own accumulation error (measured here in float64)."""
    grid = wl.make_grid(n_points - 1)
    dz = 6.0 / (n_points - 1)
    rng = np.random.default_rng(n_points + len(kind))
    phi = _tracking_potential(kind, grid, rng)
    fine = np.linspace(-3, 3, 1001).astype(np.float32)
    sig = np.gradient(phi, dz).astype(np.float32)
    r_ref = cpu.electric_field_filter(grid, sig, 0.5, fine)
    r_gpu = ops.electric_field_filter(dev(grid), dev(sig), 0.5, dev(fine)).cpu().numpy()
    pos = ops.hole_position_tracking(dev(phi), dz, 0.5, dev(grid), dev(fine), 6.0 / 1000).cpu().numpy()[0]
    # float64-exact response on the reference's own float32 arguments s = fl(fl(x_i - g_j)/L) (ref :679)
    exact = np.empty(len(fine))
    for i0 in range(0, len(fine), 64):
        s = ((fine[i0:i0 + 64, None] - grid[None, :]) / np.float32(0.5)).astype(np.float64)
        exact[i0:i0 + 64] = (np.tanh(s) / np.cosh(s)) @ sig.astype(np.float64)
    noise_ref = np.abs(r_ref - exact).max()
    assert np.abs(r_gpu - exact).max() <= max(noise_ref, 2e-6 * np.abs(exact).max())   # at least as accurate as the reference
    i_ref, i_gpu = int(np.argmax(r_ref)), int(np.argmax(r_gpu))
    assert pos == fine[i_gpu]                                    # the fused gradient+filter+argmax kernel agrees with the filter
    if i_gpu != i_ref:
        assert abs(i_gpu - i_ref) <= 1, (i_gpu, i_ref)
        assert abs(exact[i_gpu] - exact[i_ref]) <= noise_ref, (exact[i_gpu], exact[i_ref], noise_ref)


@pytest.mark.parametrize("chunk_slots", [0, 4096, -1])
def test_host_pointer_dropins(ops, cpu, chunk_slots, monkeypatch):
    """The reference-named symbols with HOST arrays (what a relinked Cython extension calls).  With
    ESPIC_HOST_CHUNK=4096 the 20 000-slot store crosses the device in 5 pipelined chunks (upload / push /
    download on three streams), each with removals: particle state, stack order and top must not change.
    chunk_slots = -1: the zero-copy mode (the kernel works on the caller's page-locked rows over PCIe)."""
    import ctypes as C
    from espic_hole_tracking_b200._lib import lib
    L = lib()
    if chunk_slots > 0:
        monkeypatch.setenv("ESPIC_HOST_CHUNK", str(chunk_slots))
    if chunk_slots < 0:
        monkeypatch.setenv("ESPIC_HOST_MODE", "zerocopy")
        monkeypatch.setenv("ESPIC_PIN_MIN_BYTES", "1024")
    n_cells = 512
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid)
    _, sp = wl.make_plasma(20_000, n_cells, seed=21, holes=0.01)
    c, h = sp.copy(), sp.copy()
    d_ref, d_h = np.zeros_like(grid), np.zeros_like(grid)
    cpu.move_particles(grid, mask, phi, 1e-3, -1836.0, 7.0, c.largest_index, c.particles, d_ref, c.empty_slots,
                       c.current_empty_slot, a_b=0.0)
    top = C.c_int(h.current_empty_slot[0])
    p = lambda a: C.c_void_p(a.ctypes.data)
    L.move_particles_c(p(grid), p(mask), p(phi), 1e-3, -1836.0, 7.0, h.largest_index[0], p(h.particles), p(d_h),
                       p(h.empty_slots), C.cast(C.byref(top), C.c_void_p), 0.0, 1, 0, len(h.empty_slots) - 1,
                       len(grid), h.particles.shape[1])
    assert top.value == c.current_empty_slot[0]
    assert_bits_equal(h.particles, c.particles, "particles")
    assert_bits_equal(h.empty_slots[:top.value + 1], c.empty_slots[:top.value + 1], "stack")
    np.testing.assert_allclose(d_h, d_ref, rtol=0, atol=2e-6 * d_ref.max())
    removed = c.empty_slots[sp.current_empty_slot[0] + 1:top.value + 1]
    assert len(removed) > 50 and np.all(np.diff(removed) > 0)          # ascending slot order, as the reference pushes
    if chunk_slots > 0:
        assert len(np.unique(removed // chunk_slots)) >= 3            # removals in at least three chunks
    L.espic_host_release()
    charge = (d_ref - d_ref.mean()).astype(np.float32)
    p_ref, p_h = np.zeros_like(grid), np.zeros_like(grid)
    cpu.poisson_solve(grid, mask, charge, 0.125, p_ref)
    L.poisson_solve_c(p(grid), p(mask), p(charge), 0.125, p(p_h), -4.0, 1.0, 0, 0, len(grid))
    assert np.abs(p_h - p_ref).max() <= REL * np.abs(p_ref).max()


@pytest.mark.parametrize("periodic", [False, True])
def test_edge_cases_empty_tiny_and_dead_stores(ops, cpu, periodic):
    """Empty store (largest_index = -1), a single particle, a store of dead slots only, and a
    ragged store whose live particles end in the middle of a quad."""
    n_cells = 64
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid)
    for n_live, storage, last in ((0, 8, -1), (1, 5, 0), (0, 16, 11), (7, 16, 9), (259, 300, 258)):
        p = np.full((2, storage), 6.0, np.float32)
        p[1] = 0.0
        rng = np.random.default_rng(storage + n_live)
        p[0, :n_live] = rng.uniform(-2.9, 2.9, n_live).astype(np.float32)
        p[1, :n_live] = rng.standard_normal(n_live).astype(np.float32) * 40
        stack = -np.ones(storage, np.int32)
        free = np.arange(storage - 1, n_live - 1, -1, dtype=np.int32)
        stack[:len(free)] = free
        c_p, c_stack, c_top, d_ref = p.copy(), stack.copy(), [len(free) - 1], np.zeros_like(grid)
        g_p, g_stack, g_top, g_den = dev(p), dev(stack), [len(free) - 1], torch.full((len(grid),), 3.0, device="cuda")
        for _ in range(2):
            cpu.move_particles(grid, mask, phi, 1e-3, -1836.0, 2.0, [last], c_p, d_ref, c_stack, c_top, a_b=0.0,
                               periodic_particles=periodic)
            ops.move_particles(dev(grid), dev(mask), dev(phi), 1e-3, -1836.0, 2.0, [last], g_p, g_den, g_stack, g_top, 0.0,
                               periodic_particles=periodic)
        assert g_top == c_top
        assert_bits_equal(g_p.cpu().numpy(), c_p, f"particles n_live={n_live}")
        assert_bits_equal(g_stack.cpu().numpy()[:g_top[0] + 1], c_stack[:c_top[0] + 1], "stack")
        np.testing.assert_allclose(g_den.cpu().numpy(), d_ref, rtol=0, atol=1e-6 * max(d_ref.max(), 1e-30))
    assert ops.default_context().status() == 0


def test_status_word_reports_what_the_reference_prints(ops):
    """Stack exhaustion on injection and stack overflow on removal only print in the reference
    (and then index out of bounds); here they raise status bits and stay in bounds."""
    grid = wl.make_grid(64)
    ctx = ops.default_context()
    ctx.status()
    # injection with more particles than free slots
    p = torch.full((2, 8), 6.0, device="cuda")
    tags = torch.zeros((1, 8), dtype=torch.int32, device="cuda")
    stack = torch.tensor([7, 6, -1, -1, -1, -1, -1, -1], dtype=torch.int32, device="cuda")
    top, last = [1], [5]
    den = torch.zeros(65, device="cuda")
    vel = torch.tensor([1.0, -1.0, 2.0, 3.0], device="cuda")
    uni = torch.full((4,), 0.5, device="cuda")
    ops.inject_particles_given(4, dev(grid), 0.01, 1.0, vel, uni, 0.0, 1, tags, p, stack, top, last, den)
    assert ctx.status() & 4          # ESPIC_ST_STACK_EMPTY
    assert top == [-1] and last == [7]
    assert abs(float(den.sum()) - 2.0) < 1e-5   # the two particles that found a slot were deposited
    # removal of more particles than the stack can hold
    p = dev(np.array([[9.0, 9.5, 10.0, 8.0], [0, 0, 0, 0]], np.float32) * 0 + np.array([[4.0], [0.0]], np.float32))
    stack = torch.tensor([-1, -1], dtype=torch.int32, device="cuda")
    top = [-1]
    ops.move_particles(dev(grid), torch.zeros(65, device="cuda"), torch.zeros(65, device="cuda"), 1e-3, 1.0, 1.0, [3], p,
                       den, stack, top, 0.0)
    assert ctx.status() & 2          # ESPIC_ST_STACK_OVERFLOW
    assert top == [3]                # the reference's counter also runs past the end
    assert stack.cpu().tolist() == [0, 1]
