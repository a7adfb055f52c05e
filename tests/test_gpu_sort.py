"""K6, the stable cell sort: permutation bit-exact against numpy's stable argsort of the mover's
own cell index, x / v / tag moved together, dead slots dropped, stack rebuilt; and sorting is
physically a no-op for the next timestep (same density, same multiset of particle states)."""
import numpy as np
import pytest

from oracle import workloads as wl
from tests.helpers import assert_bits_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import espic_hole_tracking_b200 as pkg
    return pkg


def reference_sort(grid, p, tags, n_slots, periodic, key_shift=0, lookahead=0.0):
    """numpy statement of the sort: keys with the mover's float32 arithmetic, stable argsort."""
    z_lo, z_hi = np.float32(grid[0]), np.float32(grid[-1])
    dz = np.float32((z_hi - z_lo) / np.float32(len(grid) - 1))
    x = p[0, :n_slots].copy()
    live = x.astype(np.float64) < 0.99 * float(np.float32(2 * z_hi))
    xs = x.copy()
    if lookahead:
        xs = (xs + (p[1, :n_slots] * np.float32(lookahead)).astype(np.float32)).astype(np.float32)
    if periodic:
        off = np.fmod(xs - z_lo, z_hi - z_lo)
        xs = np.where(off >= 0, off + z_lo, off + z_hi).astype(np.float32)
    with np.errstate(invalid="ignore"):
        cell = np.trunc(np.clip((xs - z_lo) / dz, -1e9, 1e9)).astype(np.int64)
    if periodic:
        cell[cell > len(grid) - 2] = 0
    cell = np.clip(cell, 0, len(grid) - 2)
    key = np.where(live, cell >> key_shift, len(grid) - 1)
    order = np.argsort(key, kind="stable")
    n_live = int(live.sum())
    order = order[:n_live]
    S = p.shape[1]
    out = np.empty_like(p)
    out[0, :], out[1, :] = np.float32(2 * z_hi), 0.0
    out[:, n_slots:] = p[:, n_slots:]
    out[:, :n_live] = p[:, order]
    out_tags = tags.copy()
    out_tags[0, :n_slots] = 0
    out_tags[0, :n_live] = tags[0, order]
    stack = -np.ones(S, np.int32)
    stack[:S - n_live] = np.arange(S - 1, n_live - 1, -1)
    return out, out_tags, stack, n_live, key[order]


@pytest.mark.parametrize("n_cells,key_shift", [(500, 0), (4096, 0), (65536, 0), (65536, 8), (65536, 5), (4096, 3),
                                               (500, 20)])
@pytest.mark.parametrize("periodic", [False, True])
def test_sort_is_stable_argsort(ops, n_cells, key_shift, periodic):
    grid = wl.make_grid(n_cells)
    rng = np.random.default_rng(n_cells)
    n, S = 150_001, 180_004
    sp = wl.make_species(n, S, rng, 40.0, -1836.0, holes=0.05)
    sp.tags[0, :n] = rng.integers(0, 1000, n)
    sp.particles[0, :4] = np.array([-3.0, 2.9999998, 0.0, -2.9999998], np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g = dict(p=dev(sp.particles), tags=dev(sp.tags), stack=dev(sp.empty_slots))
    top, last = list(sp.current_empty_slot), list(sp.largest_index)
    ops.sort_particles_by_cell(dev(grid), g["p"], g["tags"], g["stack"], top, last, periodic_particles=periodic,
                               key_shift=key_shift)
    want_p, want_tags, want_stack, n_live, keys = reference_sort(grid, sp.particles, sp.tags, n, periodic, key_shift)
    assert last == [n_live - 1] and top == [S - n_live - 1]
    assert np.all(np.diff(keys) >= 0)
    assert_bits_equal(g["p"].cpu().numpy(), want_p, "sorted particles")
    assert_bits_equal(g["tags"].cpu().numpy(), want_tags, "sorted tags")
    assert_bits_equal(g["stack"].cpu().numpy(), want_stack, "rebuilt stack")


@pytest.mark.parametrize("periodic", [False, True])
def test_sort_with_lookahead_key(ops, periodic):
    """Key = cell of fl(x + fl(v*lookahead)): fast particles are filed where they are going (clamped
    to the end cells at walls, folded when periodic)."""
    n_cells, key_shift, lookahead = 65536, 8, 4.5 * wl.timestep()
    grid = wl.make_grid(n_cells)
    rng = np.random.default_rng(9)
    n, S = 120_000, 150_000
    sp = wl.make_species(n, S, rng, 43.0, -1836.0, holes=0.05)
    sp.tags[0, :n] = rng.integers(0, 1000, n)
    sp.particles[:, :4] = np.array([[-2.9999, 2.9999, 0.0, 2.5], [-400.0, 400.0, 90.0, 3000.0]], np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g = dict(p=dev(sp.particles), tags=dev(sp.tags), stack=dev(sp.empty_slots))
    top, last = list(sp.current_empty_slot), list(sp.largest_index)
    ops.sort_particles_by_cell(dev(grid), g["p"], g["tags"], g["stack"], top, last, periodic_particles=periodic,
                               key_shift=key_shift, lookahead=lookahead)
    want_p, want_tags, want_stack, n_live, keys = reference_sort(grid, sp.particles, sp.tags, n, periodic, key_shift,
                                                                 lookahead)
    assert last == [n_live - 1] and top == [S - n_live - 1]
    assert_bits_equal(g["p"].cpu().numpy(), want_p, "sorted particles")
    assert_bits_equal(g["tags"].cpu().numpy(), want_tags, "sorted tags")
    assert_bits_equal(g["stack"].cpu().numpy(), want_stack, "rebuilt stack")


def test_coarse_sort_without_tags(ops):
    grid = wl.make_grid(65536)
    rng = np.random.default_rng(5)
    n, S = 70_000, 80_000
    sp = wl.make_species(n, S, rng, 40.0, -1836.0, holes=0.1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g = dict(p=dev(sp.particles), stack=dev(sp.empty_slots))
    top, last = list(sp.current_empty_slot), list(sp.largest_index)
    ops.sort_particles_by_cell(dev(grid), g["p"], None, g["stack"], top, last, periodic_particles=True, key_shift=8)
    want_p, _, want_stack, n_live, _ = reference_sort(grid, sp.particles, sp.tags, n, True, 8)
    assert last == [n_live - 1] and top == [S - n_live - 1]
    assert_bits_equal(g["p"].cpu().numpy(), want_p, "sorted particles")
    assert_bits_equal(g["stack"].cpu().numpy(), want_stack, "rebuilt stack")


def test_sort_is_physically_a_noop(ops):
    """One timestep after sorting gives the same density (bit for bit: integer deposit) and the
    same multiset of (x, v) as the unsorted store."""
    n_cells, n = 4096, 300_000
    grid = wl.make_grid(n_cells)
    phi = wl.smooth_potential(grid, amplitude=1.0)
    _, sp = wl.make_plasma(n, n_cells, seed=4, holes=0.03)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    results = []
    for do_sort in (False, True):
        g = dict(p=dev(sp.particles), tags=dev(sp.tags), stack=dev(sp.empty_slots))
        top, last = list(sp.current_empty_slot), list(sp.largest_index)
        if do_sort:
            ops.sort_particles_by_cell(dev(grid), g["p"], g["tags"], g["stack"], top, last, periodic_particles=True)
        den = torch.zeros(len(grid), device="cuda")
        ops.move_particles(dev(grid), torch.zeros(len(grid), device="cuda"), dev(phi), wl.timestep(), -1836.0, 7.0, last,
                           g["p"], den, g["stack"], top, 0.0, periodic_particles=True)
        p = g["p"].cpu().numpy()
        live = p[0] < 5.9
        state = np.stack([p[0][live], p[1][live]]).view(np.uint32).astype(np.uint64)
        packed = np.sort(state[0] << 32 | state[1])
        results.append((den.cpu().numpy(), packed))
    assert_bits_equal(results[0][0], results[1][0], "density after sort")
    np.testing.assert_array_equal(results[0][1], results[1][1])
