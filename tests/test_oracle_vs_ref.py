"""Pins the oracle: our C restatement (oracle/espic_oracle.c) against the compiled, unmodified
reference (oracle/_ref/libref.so) -- bit for bit -- and both against the known-answer vectors
recorded while surveying the reference (tests/golden/known_answers.json)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import cpu, workloads as wl

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _species_pair(n, n_cells, seed, holes=0.0):
    ions, electrons = wl.make_plasma(n, n_cells, seed=seed, holes=holes)
    return ions, electrons


@pytest.mark.parametrize("n_cells", [500, 512, 4096])
@pytest.mark.parametrize("periodic", [False, True])
def test_mover_bitwise(ref, port, n_cells, periodic):
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    phi = wl.smooth_potential(grid, amplitude=2.0)
    dt = np.float32(wl.timestep())
    _, sp = _species_pair(60_000, n_cells, seed=n_cells + int(periodic), holes=0.02)
    # put a few particles exactly on the walls and outside
    sp.particles[0, :6] = np.array([-3.0, 3.0, -3.0000005, 2.9999995, 3.5, -7.0], np.float32)
    a, b = sp.copy(), sp.copy()
    bg = np.float32(wl.background_density(60_000, 1, n_cells))
    for step in range(12):
        da, db = np.full_like(grid, 7.0), np.full_like(grid, -7.0)
        for k, s, d in ((ref, a, da), (port, b, db)):
            k.move_particles(grid, mask, phi, dt, sp.charge_to_mass, bg, s.largest_index,
                             s.particles, d, s.empty_slots, s.current_empty_slot,
                             a_b=np.float32(0.37 * (step % 3)), update_position=True,
                             periodic_particles=periodic)
        assert a.current_empty_slot == b.current_empty_slot
        np.testing.assert_array_equal(a.particles.view(np.uint32), b.particles.view(np.uint32))
        np.testing.assert_array_equal(a.empty_slots, b.empty_slots)
        np.testing.assert_array_equal(da.view(np.uint32), db.view(np.uint32))
    if not periodic:
        assert a.current_empty_slot[0] > sp.current_empty_slot[0]  # some particles did leave


def test_mover_no_position_update_and_object(ref, port):
    n_cells = 512
    grid = wl.make_grid(n_cells)
    mask = np.zeros_like(grid)
    mask[250:262] = 1.0
    mask[249], mask[262] = 0.4, 0.7  # partially covered end cells
    phi = wl.smooth_potential(grid)
    ions, _ = _species_pair(40_000, n_cells, seed=5)
    a, b = ions.copy(), ions.copy()
    da, db = np.zeros_like(grid), np.zeros_like(grid)
    for upd in (False, True):
        for k, s, d in ((ref, a, da), (port, b, db)):
            k.move_particles(grid, mask, phi, np.float32(1e-2), 1.0, np.float32(3.0),
                             s.largest_index, s.particles, d, s.empty_slots,
                             s.current_empty_slot, a_b=0.0, update_position=upd)
        np.testing.assert_array_equal(a.particles.view(np.uint32), b.particles.view(np.uint32))
        np.testing.assert_array_equal(a.empty_slots, b.empty_slots)
        np.testing.assert_array_equal(da.view(np.uint32), db.view(np.uint32))
    assert a.current_empty_slot[0] > ions.current_empty_slot[0]  # the object absorbed some


@pytest.mark.parametrize("n_points", [9, 513, 4097, 65537])
@pytest.mark.parametrize("periodic", [False, True])
def test_poisson_bitwise(ref, port, n_points, periodic):
    grid = wl.make_grid(n_points - 1)
    rng = np.random.default_rng(n_points)
    charge = (0.05 * rng.standard_normal(n_points)).astype(np.float32)
    mask = np.zeros_like(grid)
    pa, pb = np.zeros_like(grid), np.zeros_like(grid)
    for k, p in ((ref, pa), (port, pb)):
        k.poisson_solve(grid, mask, charge, np.float32(wl.DEBYE_LENGTH), p,
                        periodic_potential=periodic)
    np.testing.assert_array_equal(pa.view(np.uint32), pb.view(np.uint32))
    assert np.abs(pa).max() > 0


def test_poisson_object_and_boltzmann_bitwise(ref, port):
    n_points = 513
    grid = wl.make_grid(n_points - 1)
    rng = np.random.default_rng(3)
    charge = (1.0 + 0.05 * rng.standard_normal(n_points)).astype(np.float32)
    mask = np.zeros_like(grid)
    mask[200:240] = 1.0
    mask[199], mask[240] = 0.3, 0.6
    mask[300] = 0.5  # single partially covered cell: the "only point inside" branch
    for transparency in (0.0, 0.35):
        for boltz in (False, True):
            pa = (0.1 * np.sin(grid)).astype(np.float32)
            pb = pa.copy()
            for k, p in ((ref, pa), (port, pb)):
                k.poisson_solve(grid, mask, charge, np.float32(0.25), p, object_potential=-3.0,
                                object_transparency=transparency, boltzmann_electrons=boltz)
            np.testing.assert_array_equal(pa.view(np.uint32), pb.view(np.uint32))


def test_tridiagonal_bitwise(ref, port):
    rng = np.random.default_rng(0)
    n = 257
    a, c = rng.standard_normal(n), rng.standard_normal(n)
    b = 4.0 + rng.standard_normal(n)
    d = rng.standard_normal(n)
    outs = []
    for k in (ref, port):
        bb, dd, x = b.copy(), d.copy(), np.zeros(n)
        k.tridiagonal_solve(a.copy(), bb, c.copy(), dd, x)
        outs.append((bb, dd, x))
    for u, v in zip(*outs):
        np.testing.assert_array_equal(u, v)


@pytest.mark.parametrize("seed", [4357, 384, 385, 1])
def test_mt19937_stream_bitwise(ref, port, seed):
    ref.seedgenrand_c(seed)
    port.seedgenrand_c(seed)
    a = np.array([ref.genrand() for _ in range(3000)], np.float32)
    b = np.array([port.genrand() for _ in range(3000)], np.float32)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("v_th,v_d", [(42.85, 0.0), (1.0, 0.0), (1.0, 0.8), (42.85, -30.0)])
def test_draw_velocities_bitwise(ref, port, v_th, v_d):
    ref.seedgenrand_c(384)
    port.seedgenrand_c(384)
    a = ref.draw_velocities(20_000, np.float32(v_th), np.float32(v_d))
    b = port.draw_velocities(20_000, np.float32(v_th), np.float32(v_d))
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    # the streams stay aligned after a variable number of rejections
    assert ref.genrand() == port.genrand()


def test_known_answers(ref, port):
    """Vectors recorded from the compiled reference while surveying it."""
    with open(os.path.join(GOLDEN, "known_answers.json")) as fh:
        ka = json.load(fh)
    for k in (ref, port):
        for seed, expect in ka["genrand"].items():
            k.seedgenrand_c(int(seed))
            got = [k.genrand() for _ in expect]
            np.testing.assert_allclose(got, expect, rtol=0, atol=5e-11)
        k.seedgenrand_c(384)
        v = k.draw_velocities(1_000_000, np.float32(42.85), np.float32(0.0))
        np.testing.assert_allclose(v[:5], ka["draw_velocities_384_42.85_0"], rtol=2e-7)
        assert abs(np.mean(np.abs(v)) / 42.85 - math.sqrt(math.pi / 2)) < 2e-3
        assert abs(np.mean(v > 0) - 0.4995) < 1e-3
        # mirrored deposit orientation (infMagSim_c.c:219-225), dz = 0.5
        grid = np.linspace(-3, 3, 13).astype(np.float32)
        for x0, nodes in ka["deposit_orientation"]:
            p = np.array([[x0], [0.0]], np.float32)
            d = np.zeros_like(grid)
            stack = -np.ones(1, np.int32)
            k.accumulate_density(grid, np.zeros_like(grid), np.float32(1.0), [0], p, d, stack, [-1])
            for node, w in nodes:
                j = int(round((node + 3.0) / 0.5))
                assert abs(d[j] - w) < 1e-6, (x0, node, d[j], w)
            assert abs(d.sum() - 1.0) < 1e-6


@pytest.fixture(scope="module")
def inject_ref():
    """The reference's own Cython inject_particles (infMagSim_cython.py:230-285), cut unchanged out of the
    reference file and compiled by oracle/build_inject_ref.py."""
    if not cpu.have_inject_reference():
        if os.path.exists("/root/reference/infMagSim_cython.py"):
            cpu.build()
        if not cpu.have_inject_reference():
            pytest.skip("oracle/_ref/inject_ref*.so not built and /root/reference absent")
    return cpu


@pytest.mark.parametrize("n_cells,dt,v_th,v_d,a_b", [
    (500, 0.01, 1.0, 0.3, 0.2),          # dz not exactly representable: fmod vs index disagreement
    (4096, wl.timestep(), 42.85, 0.0, 0.0),   # the bench grid, electron parameters
    (4096, wl.timestep(), 1.0, 0.8, 0.37),
    (500, 0.2, 42.85, 0.0, 0.2),         # "bad left_index": |partial_dt*v| exceeds the domain for a third of the particles
])
def test_injection_port_vs_compiled_reference(port, inject_ref, capfd, n_cells, dt, v_th, v_d, a_b):
    """oracle_inject_particles == the reference's inject_particles bit for bit (state, tags, stack,
    density), the reference drawing its own velocities from its own MT19937 (same seed on both sides)."""
    grid = wl.make_grid(n_cells)
    ions, _ = wl.make_plasma(2000, n_cells, seed=9)
    a, b, c = ions.copy(), ions.copy(), ions.copy()
    da = np.random.default_rng(1).random(len(grid)).astype(np.float32)
    db, dc = da.copy(), da.copy()
    n, bg, k = 300, np.float32(4.0), 17
    dt, v_th, v_d, a_b = np.float32(dt), np.float32(v_th), np.float32(v_d), np.float32(a_b)
    inject_ref.reference_inject_particles(n, grid, dt, v_th, bg, wl.HostUniformSampler(5), v_d, a_b, k, a.tags,
                                          a.particles, a.empty_slots, a.current_empty_slot, a.largest_index, da,
                                          seed=384)
    port.seedgenrand_c(384)
    vel = port.draw_velocities(n, v_th, v_d)
    uni = np.ascontiguousarray(wl.HostUniformSampler(5).get(n).astype(np.float32).T[0])
    port.inject_given(n, grid, dt, bg, vel, uni, a_b, k, b.tags, b.particles, b.empty_slots, b.current_empty_slot,
                      b.largest_index, db)
    # and the reference body fed the same velocities through the shim build
    inject_ref.reference_inject_given(n, grid, dt, bg, vel, uni, a_b, k, c.tags, c.particles, c.empty_slots,
                                      c.current_empty_slot, c.largest_index, dc)
    capfd.readouterr()  # the reference prints one line per bad particle
    bad = dt > 0.1
    for r, dr in ((a, da), (c, dc)):
        assert r.current_empty_slot == b.current_empty_slot
        assert r.largest_index == b.largest_index
        np.testing.assert_array_equal(r.particles.view(np.uint32), b.particles.view(np.uint32))
        np.testing.assert_array_equal(r.empty_slots, b.empty_slots)
        np.testing.assert_array_equal(r.tags, b.tags)
        np.testing.assert_array_equal(dr.view(np.uint32), db.view(np.uint32))
    if bad:   # bad particles keep their slot on the stack (ref :272-274)
        assert ions.current_empty_slot[0] - n < b.current_empty_slot[0] < ions.current_empty_slot[0]
    else:
        assert b.current_empty_slot == [ions.current_empty_slot[0] - n]


def test_injection_python_restatement_vs_compiled_reference(inject_ref):
    """The scalar Python restatement (the fallback of the "reference" arm where the compiled extension is
    absent) against the compiled reference function."""
    n_cells = 500
    grid = wl.make_grid(n_cells)
    ions, _ = wl.make_plasma(2000, n_cells, seed=9)
    a, b = ions.copy(), ions.copy()
    da = np.random.default_rng(1).random(len(grid)).astype(np.float32)
    db = da.copy()
    ref_k = cpu.load("reference")
    ref_k.seedgenrand_c(384)
    vel = ref_k.draw_velocities(300, np.float32(1.0), np.float32(0.3))
    uni = np.random.default_rng(2).random(300).astype(np.float32)
    inject_ref.reference_inject_given(300, grid, np.float32(0.01), np.float32(4.0), vel, uni, np.float32(0.2), 17,
                                      a.tags, a.particles, a.empty_slots, a.current_empty_slot, a.largest_index, da)
    cpu.inject_python(300, grid, np.float32(0.01), np.float32(4.0), vel, uni, np.float32(0.2), 17,
                      b.tags, b.particles, b.empty_slots, b.current_empty_slot, b.largest_index, db)
    assert a.current_empty_slot == b.current_empty_slot == [ions.current_empty_slot[0] - 300]
    assert a.largest_index == b.largest_index
    np.testing.assert_array_equal(a.particles.view(np.uint32), b.particles.view(np.uint32))
    np.testing.assert_array_equal(a.empty_slots, b.empty_slots)
    np.testing.assert_array_equal(a.tags, b.tags)
    np.testing.assert_array_equal(da.view(np.uint32), db.view(np.uint32))


def test_diagnostics_bitwise(ref, port):
    grid = wl.make_grid(512)
    rng = np.random.default_rng(4)
    sig = rng.standard_normal(len(grid)).astype(np.float32)
    fine = np.linspace(-3, 3, 101).astype(np.float32)
    ra = ref.electric_field_filter(grid, sig, np.float32(0.5), fine)
    rb = port.electric_field_filter(grid, sig, np.float32(0.5), fine)
    np.testing.assert_array_equal(ra.view(np.uint32), rb.view(np.uint32))
    x = rng.uniform(-3.2, 3.2, 50_000).astype(np.float32)
    v = (4 * rng.standard_normal(50_000)).astype(np.float32)
    ha = ref.histogram2d_uniform_grid(x, v, -3.0, 3.0, -8.0, 8.0, 100, 100)[0]
    hb = port.histogram2d_uniform_grid(x, v, -3.0, 3.0, -8.0, 8.0, 100, 100)[0]
    np.testing.assert_array_equal(ha, hb)
    assert ha.sum() > 40_000
