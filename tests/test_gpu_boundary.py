"""Boundary and robustness tests on the GPU (round 2): the two remaining reference C symbols
(tridiagonal_solve_c, move_particles_c_minimal) through the C ABI with host arrays, contexts on two
devices of one process, the injection body against the reference's own compiled Cython routine, and
checkpoint/resume of a wall-bounded (injecting) run."""
import ctypes as C

import numpy as np
import pytest

from oracle import cpu as ocpu, workloads as wl
from tests.helpers import assert_bits_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import espic_hole_tracking_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def cpu():
    if ocpu.have("reference"):
        return ocpu.load("reference")
    if not ocpu.have("port"):
        ocpu.build()
    return ocpu.load("port")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("n", [1, 2, 7, 513, 1024, 1025, 4097, 65537])
def test_tridiagonal_solve_c_host_arrays(ops, cpu, n):
    """tridiagonal_solve_c (ref infMagSim_c.c:343-354): x, and the in-place modified b and d, bit for bit."""
    from espic_hole_tracking_b200._lib import lib
    rng = np.random.default_rng(n)
    a = rng.uniform(0.5, 1.5, n)
    c = rng.uniform(0.5, 1.5, n)
    b = -(a + c) - rng.uniform(0.1, 1.0, n)      # diagonally dominant, like the solver's rows
    d = rng.standard_normal(n)
    want = [v.copy() for v in (a, b, c, d)] + [np.zeros(n)]
    got = [v.copy() for v in (a, b, c, d)] + [np.full(n, 7.0)]
    cpu.tridiagonal_solve(*want)
    lib().tridiagonal_solve_c(*[_p(v) for v in got], n)
    for name, w, g in zip("abcdx", want, got):
        np.testing.assert_array_equal(w.view(np.uint64), g.view(np.uint64), err_msg=name)


@pytest.mark.parametrize("update_position", [1, 0])
def test_move_particles_c_minimal_host_arrays(ops, update_position):
    """move_particles_c_minimal (ref infMagSim_c.c:231-341) against the compiled reference symbol: no mask, no
    frame acceleration, and out-of-bounds slots that do not carry the parked marker are left alone."""
    if not ocpu.have("reference"):
        pytest.skip("needs oracle/_ref/libref.so (the restatement has no minimal variant)")
    from espic_hole_tracking_b200._lib import lib
    ref = C.CDLL(ocpu.REF_SO)
    n_cells = 500
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    mask[100:120] = 1.0                                  # must be ignored
    phi = wl.smooth_potential(grid, amplitude=2.0)
    _, sp = wl.make_plasma(30_000, n_cells, seed=31, holes=0.02)
    sp.particles[0, :6] = np.array([-3.0, 3.0, -3.0000005, 2.9999995, 3.5, -7.0], np.float32)   # 3.5 / -7: dead but not parked
    a, b = sp.copy(), sp.copy()
    for step in range(3):
        outs = []
        for fn, s in ((ref.move_particles_c_minimal, a), (lib().move_particles_c_minimal, b)):
            den = np.full_like(grid, 5.0)
            top = C.c_int(s.current_empty_slot[0])
            fn.restype = None
            fn(_p(grid), _p(mask), _p(phi), C.c_float(wl.timestep()), C.c_float(-1836.0), C.c_float(9.0),
               C.c_int(s.largest_index[0]), _p(s.particles), _p(den), _p(s.empty_slots),
               C.cast(C.byref(top), C.c_void_p),
               C.c_int(update_position), C.c_int(1), C.c_int(len(s.empty_slots) - 1), C.c_int(len(grid)),
               C.c_int(s.particles.shape[1]))
            s.current_empty_slot[0] = top.value
            outs.append(den)
        assert a.current_empty_slot == b.current_empty_slot
        assert_bits_equal(b.particles, a.particles, "particles")
        assert_bits_equal(b.empty_slots[:b.current_empty_slot[0] + 1], a.empty_slots[:a.current_empty_slot[0] + 1], "stack")
        np.testing.assert_allclose(outs[1], outs[0], rtol=0, atol=3e-6 * outs[0].max())
    assert a.particles[0, 4] == np.float32(3.5) and a.particles[0, 5] == np.float32(-7.0)   # left alone (ref :333)
    if update_position:
        assert a.current_empty_slot[0] > sp.current_empty_slot[0]


def test_contexts_on_two_devices_in_one_process(ops, cpu):
    """Launch configuration (>48 KB shared-memory opt-ins, occupancy choices) is per context: a second
    context on another device works, whatever device is current, and leaves the caller's device alone."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from espic_hole_tracking_b200._lib import Context
    n_cells = 4096
    grid = wl.make_grid(n_cells)
    phi = wl.smooth_potential(grid)
    _, sp = wl.make_plasma(50_000, n_cells, seed=3)
    want = sp.copy()
    d_ref = np.zeros_like(grid)
    cpu.move_particles(grid, np.zeros_like(grid), phi, wl.timestep(), -1836.0, 5.0, want.largest_index, want.particles,
                       d_ref, want.empty_slots, want.current_empty_slot, a_b=0.0, periodic_particles=True)
    torch.cuda.set_device(0)
    for d in (0, 1, 0):
        ctx = Context(d)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{d}")
        p, stack, den = to(sp.particles), to(sp.empty_slots), torch.zeros(len(grid), device=f"cuda:{d}")
        top = [sp.current_empty_slot[0]]
        ops.move_particles(to(grid), torch.zeros(len(grid), device=f"cuda:{d}"), to(phi), wl.timestep(), -1836.0, 5.0,
                           [sp.largest_index[0]], p, den, stack, top, 0.0, periodic_particles=True, ctx=ctx)
        assert torch.cuda.current_device() == 0
        assert_bits_equal(p.cpu().numpy(), want.particles, f"device {d}")
        # and the large-grid kernels (window mover, sort, cluster solve) on the same context
        g2 = to(wl.make_grid(65536))
        rho = torch.zeros_like(g2)
        pot = torch.zeros_like(g2)
        ops.poisson_solve(g2, torch.zeros_like(g2), rho, 0.125, pot, ctx=ctx)
        assert ctx.status() == 0
        ctx.close()


@pytest.mark.parametrize("n_cells,dt,a_b", [(500, 0.01, 0.2), (4096, wl.timestep(), 0.0), (500, 0.2, 0.2)])
def test_injection_against_the_compiled_reference_routine(ops, capfd, n_cells, dt, a_b):
    """espic_inject_particles against THE reference's inject_particles (infMagSim_cython.py:230-285, compiled
    unchanged by oracle/build_inject_ref.py) on the same velocities and fractions: state, tags, stack, top and
    largest index bit for bit, density within 1e-6; the third case has bad left_index particles."""
    if not ocpu.have_inject_reference():
        pytest.skip("oracle/_ref/inject_ref*.so not built")
    grid = wl.make_grid(n_cells)
    ions, _ = wl.make_plasma(5000, n_cells, seed=9)
    c = ions.copy()
    rng = np.random.default_rng(n_cells)
    n, bg = 700, 4.0
    v_scale = 42.85 if dt > 0.1 or n_cells == 4096 else 1.0
    vel = (rng.standard_normal(n) * v_scale).astype(np.float32)
    uni = rng.random(n).astype(np.float32)
    den0 = rng.random(len(grid)).astype(np.float32)
    d_ref = den0.copy()
    ocpu.reference_inject_given(n, grid, np.float32(dt), np.float32(bg), vel, uni, np.float32(a_b), 23, c.tags,
                                c.particles, c.empty_slots, c.current_empty_slot, c.largest_index, d_ref)
    capfd.readouterr()
    g = dict(p=dev(ions.particles), tags=dev(ions.tags), stack=dev(ions.empty_slots))
    top, last, den = [ions.current_empty_slot[0]], [ions.largest_index[0]], dev(den0)
    ops.inject_particles_given(n, dev(grid), dt, bg, dev(vel), dev(uni), a_b, 23, g["tags"], g["p"], g["stack"], top,
                               last, den)
    assert top == c.current_empty_slot and last == c.largest_index
    assert_bits_equal(g["p"].cpu().numpy(), c.particles, "particles")
    assert_bits_equal(g["tags"].cpu().numpy(), c.tags, "tags")
    np.testing.assert_array_equal(g["stack"].cpu().numpy(), c.empty_slots)
    np.testing.assert_allclose(den.cpu().numpy(), d_ref, rtol=0, atol=2e-6 * d_ref.max())
    ops.default_context().status()   # the bad-index case raises ESPIC_ST_BAD_INJECT_INDEX by design: clear it


def test_walls_checkpoint_resume_is_bit_identical(ops, tmp_path):
    """A wall-bounded run injects from the device sampler every step: the checkpoint carries the sampler's
    stream position (espic_rng_get_state), so the resumed run draws what the uninterrupted one drew."""
    from espic_hole_tracking_b200 import io as eio
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    cfg = SimConfig(n_steps=40, n_particles=60_000, n_cells=512, periodic_particles=False, track_hole=True,
                    initial_transient_steps=3)
    sim = Simulation(cfg)
    sim.init_synthetic()
    sim.initialize()
    sim.run(12)
    eio.save_checkpoint(sim, str(tmp_path / "walls.pt"))
    sim.run(9)
    resumed = Simulation(cfg)
    resumed.init_synthetic()
    eio.load_checkpoint(resumed, str(tmp_path / "walls.pt"))
    resumed.run(9)
    assert resumed.n_active() == sim.n_active()
    for a, b, what in ((resumed.potential, sim.potential, "potential"),
                       (resumed.electrons.particles, sim.electrons.particles, "electrons"),
                       (resumed.ions.particles, sim.ions.particles, "ions"),
                       (resumed.electrons.tags, sim.electrons.tags, "tags"),
                       (resumed.electrons.empty_slots, sim.electrons.empty_slots, "stack"),
                       (resumed.hole_relative_positions, sim.hole_relative_positions, "hole positions")):
        assert_bits_equal(a.cpu().numpy(), b.cpu().numpy(), what)
    # a checkpoint of another configuration (other seed / flags) is refused
    other = Simulation(SimConfig(n_steps=40, n_particles=60_000, n_cells=512, periodic_particles=True))
    with pytest.raises(ValueError):
        eio.load_checkpoint(other, str(tmp_path / "walls.pt"))
