"""Integration parity on the GPU: N steps of the full driver loop against the oracle loop
(oracle/loop.py, the Python-3 restatement of infMagSim_script.py:756-1023 over the compiled
reference kernels), periodic and non-periodic with wall injection."""
import os

import numpy as np
import pytest

from oracle import loop as oloop, workloads as wl
from tests.helpers import assert_bits_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import espic_hole_tracking_b200 as p
    return p


def kind():
    from oracle import cpu
    return "reference" if cpu.have("reference") else "port"


def particles(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for v_th in (1.0, np.sqrt(wl.MASS_RATIO)):
        x = rng.uniform(wl.Z_MIN, wl.Z_MAX, n)
        x = x + 0.05 * np.sin(2 * np.pi * x / 6.0)
        out.append(np.stack([x.astype(np.float32), (rng.standard_normal(n) * v_th).astype(np.float32)]))
    return out


@pytest.mark.parametrize("n_cells", [512, 4096])
def test_periodic_loop_lockstep(pkg, n_cells):
    """Periodic driver loop in lockstep with the oracle: each step the GPU solves the field from the
    ORACLE's densities (so the solve is compared on identical inputs: <=1e-6) and pushes with the
    oracle's potential (so particle state must stay bit-identical step after step)."""
    ops = pkg
    n, steps = 200_000, 8
    ions, electrons = particles(n, 3)
    cpu = oloop.CpuSimulation(n, n_cells, True, kind=kind())
    cpu.load_particles(ions, electrons)
    cpu.initialize()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    grid, mask = dev(cpu.grid), dev(cpu.object_mask)
    g = [dict(p=dev(s.particles), stack=dev(s.empty_slots), top=ops.DeviceInt(s.current_empty_slot[0]),
              last=ops.DeviceInt(s.largest_index[0]), qm=s.charge_to_mass) for s in (cpu.ions, cpu.electrons)]
    rho, phi = torch.zeros_like(grid), torch.zeros_like(grid)
    for k in range(steps):
        dens = torch.stack([dev(cpu.ion_density), dev(cpu.electron_density)])
        ops.field_solve(grid, mask, dens[0], dens[1], rho, wl.DEBYE_LENGTH, phi, object_potential=-3.,
                        periodic_particles=True)
        cpu.field_step()
        assert_bits_equal(dens[0].cpu().numpy(), cpu.ion_density, "end-node fix")
        assert_bits_equal(rho.cpu().numpy(), cpu.charge_density, "charge density")
        scale = np.abs(cpu.potential).max()
        assert np.abs(phi.cpu().numpy().astype(np.float64) - cpu.potential).max() <= 1e-6 * scale, k
        phi_same = dev(cpu.potential)
        cpu.push_step()
        cpu.k += 1
        for sp, den, ref, dref in zip(g, dens, (cpu.ions, cpu.electrons), (cpu.ion_density, cpu.electron_density)):
            ops.move_particles(grid, mask, phi_same, cpu.dt, sp["qm"], cpu.bg, sp["last"], sp["p"], den, sp["stack"],
                               sp["top"], 0.0, periodic_particles=True)
            assert_bits_equal(sp["p"].cpu().numpy(), ref.particles, f"particles step {k}")
            assert np.abs(den.cpu().numpy() - dref).max() <= 3e-6 * dref.max()
    assert ops.default_context().status() == 0


def test_periodic_simulation_free_running(pkg):
    """The device-resident driver (one espic_pic_step host call per step) free-running next to the
    oracle loop from identical particle arrays.  Bitwise: fused call == operator-by-operator.
    Against the oracle only statistics can be compared: the two sides' potentials differ at the
    1e-3 level after one step (float32 summation order of the density, amplified by the solve's
    net-charge mode ~(L/lambda_D)^2/8), and the reference's mirrored weights make the deposit
    discontinuous at a node, so single particles crossing a node on one side only change two nodes
    by 1/particles-per-cell."""
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    n_cells, n, steps = 512, 200_000, 10
    ions, electrons = particles(n, 3)
    cpu = oloop.CpuSimulation(n, n_cells, True, kind=kind())
    cpu.load_particles(ions, electrons)
    cpu.initialize()
    cfg = SimConfig(n_steps=steps, n_particles=n, n_cells=n_cells, periodic_particles=True, track_hole=False)
    sims = []
    for _ in range(2):
        sim = Simulation(cfg)
        sim.load_particles(ions, electrons)
        sim.initialize()
        sims.append(sim)
    sims[1].cfg = SimConfig(**{**cfg.__dict__, "phase_space_histograms": True})  # forces the operator-by-operator path
    assert_bits_equal(sims[0].ions.particles.cpu().numpy()[:, :n], cpu.ions.particles[:, :n], "after initialize_mover")
    for k in range(steps):
        cpu.step()
        for sim in sims:
            sim.step()
        assert_bits_equal(sims[0].potential.cpu().numpy(), sims[1].potential.cpu().numpy(), "fused vs operators")
        assert_bits_equal(sims[0].electrons.particles.cpu().numpy(), sims[1].electrons.particles.cpu().numpy())
        scale = np.abs(cpu.potential).max()
        assert np.abs(sims[0].potential.cpu().numpy() - cpu.potential).max() <= (3e-3 if k == 0 else 0.1) * scale, k
    assert sims[0].n_active() == cpu.n_active()
    e = sims[0].energies()
    v = cpu.electrons.particles[1, :n].astype(np.float64)
    ke_cpu = 0.5 / wl.MASS_RATIO * (v * v).sum() / cpu.bg
    assert abs(e["ke_electrons"] - ke_cpu) <= 5e-3 * ke_cpu, (e["ke_electrons"], ke_cpu)
    assert sims[0].ctx.status() == 0


@pytest.mark.parametrize("fast_fraction", [0.0002, 0.3])
def test_periodic_mass_ejection_keeps_the_reference_stack_order(pkg, fast_fraction):
    """A periodic box only loses a particle that crosses more than a box length in one step (ref
    infMagSim_c.c:172); the step path pushes such slots on the stack as they come and the next field solve
    restores the reference's ascending order.  A handful is sorted by one thread; a run that blows up can
    eject a third of the store at once -- the whole team sorts, in milliseconds, instead of one thread
    for minutes.  Checked against the operator-by-operator path, whose removal epilogue builds the stack in
    slot order directly."""
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    n_cells, n, steps = 512, 200_000, 3
    ions, electrons = particles(n, 5)
    rng = np.random.default_rng(11)
    fast = rng.random(n) < fast_fraction
    electrons[1][fast] = np.float32(4.0e4)          # 35 box lengths per step: out of the box for good
    cfg = SimConfig(n_steps=steps, n_particles=n, n_cells=n_cells, periodic_particles=True, track_hole=False)
    sims = []
    for _ in range(2):
        sim = Simulation(cfg)
        sim.load_particles(ions, electrons)
        sim.initialize()
        sims.append(sim)
    sims[1].cfg = SimConfig(**{**cfg.__dict__, "phase_space_histograms": True})  # operator-by-operator path
    for k in range(steps):
        for sim in sims:
            sim.step()
    torch.cuda.synchronize()
    a, b = sims
    assert a.n_active() == b.n_active()
    assert a.n_active()["electrons"] <= n - int(0.9 * fast.sum())
    top = int(b.electrons.current_empty_slot[0])
    assert int(a.electrons.current_empty_slot[0]) == top
    # the stretch pushed by the last step is ordered by the NEXT solve: take one more step on both before comparing
    for sim in sims:
        sim.step()
    np.testing.assert_array_equal(a.electrons.empty_slots.cpu().numpy()[:top + 1], b.electrons.empty_slots.cpu().numpy()[:top + 1])
    assert_bits_equal(a.electrons.particles.cpu().numpy(), b.electrons.particles.cpu().numpy(), "electrons")
    assert_bits_equal(a.potential.cpu().numpy(), b.potential.cpu().numpy(), "potential")


def test_nonperiodic_injection_loop_bit_exact_particles(pkg):
    """Walls + injection over several steps with identical random inputs on both sides.  The GPU
    side is pushed with the oracle's potential each step, so particle state, free-slot stack,
    injection tags and stack tops must match bit for bit through removal and re-injection; the
    GPU's own field solve is checked against the oracle's at every step."""
    ops = pkg
    n_cells, n, steps = 500, 60_000, 10
    ions, electrons = particles(n, 11)
    cpu = oloop.CpuSimulation(n, n_cells, False, kind=kind())
    cpu.load_particles(ions, electrons)
    cpu.initialize()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    grid, mask = dev(cpu.grid), dev(cpu.object_mask)
    g = []
    for s in (cpu.ions, cpu.electrons):   # the oracle's post-initialize state (already compared elsewhere)
        g.append(dict(p=dev(s.particles), tags=dev(s.tags), stack=dev(s.empty_slots),
                      top=ops.DeviceInt(s.current_empty_slot[0]), last=ops.DeviceInt(s.largest_index[0]),
                      qm=s.charge_to_mass))
    dens = torch.stack([dev(cpu.ion_density), dev(cpu.electron_density)])
    rho, phi = torch.zeros_like(grid), torch.zeros_like(grid)
    rng = np.random.default_rng(5)
    from oracle import cpu as ocpu
    sampler = ocpu.load(kind())
    sampler.seedgenrand_c(384)
    for k in range(steps):
        # --- field: both sides solve; compare; then the GPU pushes with the oracle's potential
        ops.field_solve(grid, mask, dens[0], dens[1], rho, wl.DEBYE_LENGTH, phi, object_potential=-3.,
                        periodic_particles=False)
        cpu.field_step()
        scale = np.abs(cpu.potential).max()
        assert np.abs(phi.cpu().numpy() - cpu.potential).max() <= 3e-3 * scale
        phi_same = dev(cpu.potential)
        cpu.push_step()
        for sp, den in zip(g, dens):
            ops.move_particles(grid, mask, phi_same, cpu.dt, sp["qm"], cpu.bg, sp["last"], sp["p"], den, sp["stack"],
                               sp["top"], 0.0, periodic_particles=False)
        # --- injection with the same counts, velocities and fractions
        n_i, n_e = cpu.injection_counts()
        vel = {"ions": sampler.draw_velocities(n_i, 1.0, 0.0), "electrons": sampler.draw_velocities(n_e, cpu.v_th_e, 0.0)}
        uni = {"ions": rng.random(n_i).astype(np.float32), "electrons": rng.random(n_e).astype(np.float32)}
        cpu.inject_step(counts=(n_i, n_e), velocities=vel, uniforms=uni)
        for name, sp, den, cnt in (("ions", g[0], dens[0], n_i), ("electrons", g[1], dens[1], n_e)):
            if cnt > 0:
                ops.inject_particles_given(cnt, grid, cpu.dt, cpu.bg, dev(vel[name]), dev(uni[name]), 0.0, k, sp["tags"],
                                           sp["p"], sp["stack"], sp["top"], sp["last"], den)
        cpu.t += cpu.dt
        cpu.k += 1
        for sp, ref in zip(g, (cpu.ions, cpu.electrons)):
            assert sp["top"][0] == ref.current_empty_slot[0] and sp["last"][0] == ref.largest_index[0], k
            assert_bits_equal(sp["p"].cpu().numpy(), ref.particles, f"particles step {k}")
            assert_bits_equal(sp["stack"].cpu().numpy(), ref.empty_slots, f"stack step {k}")
            assert_bits_equal(sp["tags"].cpu().numpy(), ref.tags, f"tags step {k}")
        for den, ref in zip(dens, (cpu.ion_density, cpu.electron_density)):
            assert np.abs(den.cpu().numpy() - ref).max() <= 3e-6 * ref.max()
    assert cpu.electrons.current_empty_slot[0] != electrons.shape[1]  # electrons did leave and were re-injected
    assert int(cpu.electrons.tags.max()) == steps - 1
    assert ops.default_context().status() == 0


def test_nonperiodic_simulation_runs_and_conserves(pkg):
    """The device-resident driver with its own (Philox) injection: population stays near n, the
    free-slot stack stays consistent, no status bits."""
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    cfg = SimConfig(n_steps=40, n_particles=100_000, n_cells=512, periodic_particles=False, track_hole=True,
                    initial_transient_steps=5)
    sim = Simulation(cfg)
    sim.init_synthetic()
    sim.initialize()
    sim.run(40)
    act = sim.n_active()
    assert abs(act["electrons"] - 100_000) < 3_000 and abs(act["ions"] - 100_000) < 500
    for s, name in ((sim.ions, "ions"), (sim.electrons, "electrons")):
        top = s.current_empty_slot[0]
        free = s.empty_slots.cpu().numpy()[:top + 1]
        assert len(np.unique(free)) == len(free) and free.min() >= 0
        x = s.particles[0].cpu().numpy()
        assert np.all(x[free] == np.float32(6.0))                       # every listed slot is parked
        assert (x < 5.9).sum() == act[name] == s.storage - (top + 1)    # every unlisted slot is live
    assert sim.status_word() == 0
    e = sim.energies()
    assert e["fe"] > 0 and e["ke_electrons"] > 0
    pos = sim.hole_relative_positions.cpu().numpy()
    assert np.all(np.abs(pos) <= 3.0)


def test_nonperiodic_fused_step_equals_operator_path(pkg):
    """Wall-bounded step: the fused host call (deposit left in the accumulators, removal epilogue in one launch)
    against the same step through the individual operators (k_finish + k_scatter, float densities).  The two
    paths hand the device sampler's uniforms out differently (one per particle / the first of a pair), so the
    injected particles themselves are not comparable; everything else of the step is, bit for bit: potential,
    the free-slot stack in the reference's ascending order, which slots were refilled, and every particle that
    was not injected in this step."""
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    cfg = SimConfig(n_steps=30, n_particles=150_000, n_cells=512, periodic_particles=False, track_hole=True,
                    initial_transient_steps=3)
    sims = []
    for _ in range(2):
        sim = Simulation(cfg)
        sim.init_synthetic()
        sim.initialize()
        sims.append(sim)
    sims[1].cfg = SimConfig(**{**cfg.__dict__, "phase_space_histograms": True})  # operator-by-operator path
    a, b = sims
    for sim in sims:   # injection writes the step index (0) into the tag of the slot it fills
        sim.ions.tags.fill_(-1)
        sim.electrons.tags.fill_(-1)
    a.step()
    b.step()
    assert_bits_equal(a.potential.cpu().numpy(), b.potential.cpu().numpy(), "potential")
    for name in ("ions", "electrons"):
        sa, sb = getattr(a, name), getattr(b, name)
        top = int(sb.current_empty_slot[0])
        assert int(sa.current_empty_slot[0]) == top, name
        np.testing.assert_array_equal(sa.empty_slots.cpu().numpy()[:top + 1], sb.empty_slots.cpu().numpy()[:top + 1],
                                      err_msg=f"{name} stack")
        ta, tb = sa.tags.cpu().numpy(), sb.tags.cpu().numpy()
        np.testing.assert_array_equal(ta, tb, err_msg=f"{name}: slots refilled")
        pa, pb = sa.particles.cpu().numpy(), sb.particles.cpu().numpy()
        injected = tb[0] == 0
        if name == "electrons":
            assert injected.sum() > 100   # particles did leave and enter
        keep = ~injected
        assert_bits_equal(pa[:, keep], pb[:, keep], f"{name}: particles that were not injected")
        assert np.all(pa[0, injected] < 5.9)
    assert a.n_active() == b.n_active()
    assert a.status_word() == 0 and b.status_word() == 0


def test_npz_output_and_checkpoint_resume(pkg, tmp_path):
    """The reference's .npz key set (consumed by its analysis notebook) and a checkpoint/resume
    round trip that continues a periodic run bit-identically."""
    from espic_hole_tracking_b200 import io as eio
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    cfg = SimConfig(n_steps=30, n_particles=50_000, n_cells=256, periodic_particles=True, periodic_potential=False,
                    track_hole=True, initial_transient_steps=3, phase_space_histograms=True)
    sim = Simulation(cfg)
    sim.init_synthetic()
    sim.initialize()
    rec = eio.Recorder(sim)
    for _ in range(12):
        sim.step()
        rec.record()
    eio.save_checkpoint(sim, str(tmp_path / "ck.pt"))
    for _ in range(8):
        sim.step()
        rec.record()
    out = rec.save(str(tmp_path))
    data = np.load(out)
    assert set(data.files) == set(eio.NPZ_KEYS)
    assert data["potentials"].shape == (20, 257) and data["times"].shape == (20,)
    assert data["electron_distribution_functions"].shape == (20, 100, 100)
    assert 49_900 <= data["electron_distribution_functions"][0].sum() <= 50_000   # |v| < 4 v_th window
    assert os.path.basename(out).startswith("l0.1250_Vd0.000_hole_depth")
    resumed = Simulation(cfg)
    resumed.init_synthetic()
    eio.load_checkpoint(resumed, str(tmp_path / "ck.pt"))
    for _ in range(8):
        resumed.step()
    assert_bits_equal(resumed.potential.cpu().numpy(), sim.potential.cpu().numpy(), "resumed potential")
    assert_bits_equal(resumed.electrons.particles.cpu().numpy(), sim.electrons.particles.cpu().numpy(), "resumed electrons")


def _live_state(sim):
    out = []
    for s in (sim.ions, sim.electrons):
        p = s.particles.cpu().numpy()
        live = p[0] < 5.9
        st = np.stack([p[0][live], p[1][live]]).view(np.uint32).astype(np.uint64)
        out.append(np.sort(st[0] << 32 | st[1]))
    return out


@pytest.mark.parametrize("periodic", [True, False])
def test_large_grid_auto_sort_is_physically_invisible(pkg, periodic):
    """C4 grid (65536 cells): the driver's automatic coarse re-sorting for the windowed mover
    changes slot numbers only -- after 25 steps the densities and the potential are bit-identical
    to a run that never sorts, and the live (x, v) multisets are equal (integer deposit: order
    independent; with walls the injected particles take other slots but the same states)."""
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    runs = []
    for never in (True, False):
        kw = dict(sort_interval_ions=0, sort_interval_electrons=0) if never else dict(sort_interval_ions=7)
        cfg = SimConfig(n_steps=25, n_particles=300_000, n_cells=65536, periodic_particles=periodic,
                        periodic_potential=False, track_hole=False, **kw)
        sim = Simulation(cfg)
        sim.init_synthetic()
        sim.initialize()
        if not never:
            assert sim.sort_intervals["electrons"] == (36 if periodic else 32) and sim.sort_key_shift == 8
        sim.run(25)
        assert sim.status_word() == 0
        runs.append((sim.ion_density.cpu().numpy().copy(), sim.electron_density.cpu().numpy().copy(),
                     sim.potential.cpu().numpy().copy(), _live_state(sim), sim.n_active()))
    a, b = runs
    assert a[4] == b[4]
    for i, what in enumerate(("ion density", "electron density", "potential")):
        assert_bits_equal(a[i], b[i], what)
    for x, y in zip(a[3], b[3]):
        np.testing.assert_array_equal(x, y)


def test_moving_box_controller_in_the_loop(pkg):
    """moving_box_simulation=True (off in the reference's defaults, infMagSim_script.py:141): the
    control law runs on the host between the field solve and the pushes of every step.  Fed the hole
    positions the run itself tracked, the reference's own loop (:917-928, scipy filtfilt over the
    whole history each step) must reproduce the box velocities and accelerations the run used."""
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    from tests.test_host_logic import _reference_box_loop
    n_steps, end = 160, 150
    cfg = SimConfig(n_steps=n_steps, n_particles=100_000, n_cells=512, periodic_particles=False, track_hole=True,
                    moving_box_simulation=True, hole_tracking_end_step=end, v_b_0=0.25)
    sim = Simulation(cfg)
    sim.init_synthetic()
    sim.initialize()
    sim.run(n_steps)
    # a synthetic start has no settled hole yet: the tracked position jumps about, the law answers with
    # large accelerations, and a few injected particles can land outside the box ("bad left_index",
    # infMagSim_cython.py:272-273) exactly as they would in the reference; nothing else may be flagged
    assert sim.status_word() & ~8 == 0
    pos = sim.hole_relative_positions.cpu().numpy()
    assert np.all(pos[:81] == 0) and np.abs(pos[81:]).max() > 0
    want_box, want_hv = _reference_box_loop(pos, sim.dt, 0.25, cfg.step_size, sim.v_th_e, cfg.debye_length, end)
    np.testing.assert_allclose(sim.ctrl.hole_velocities[:n_steps], want_hv, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(sim.ctrl.box[:, :n_steps + 1], want_box, rtol=2e-6, atol=1e-5 * np.abs(want_box).max())
    assert np.abs(sim.ctrl.box[1, 101:end]).max() > 0          # the frame did accelerate ...
    assert np.all(sim.ctrl.box[1, :101] == 0) and np.all(sim.ctrl.box[1, end:] == 0)   # ... only while tracking
    act = sim.n_active()
    assert abs(act["electrons"] - 100_000) < 20_000 and abs(act["ions"] - 100_000) < 20_000
    # same configuration with the controller off: frame at rest, and no host read-back was needed
    cfg2 = SimConfig(n_steps=40, n_particles=100_000, n_cells=512, periodic_particles=False, track_hole=True, v_b_0=0.25)
    sim2 = Simulation(cfg2)
    sim2.init_synthetic()
    sim2.initialize()
    sim2.run(40)
    assert np.all(sim2.ctrl.box[0] == np.float32(0.25)) and not sim2.ctrl.box[1].any()
