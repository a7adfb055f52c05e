"""Long-run statistical agreement (BASELINE configs[4], scaled down): an electron-hole run with
walls and injection, started from identical particle arrays (the reference's phase-space dimple
included), free-running on both sides with their own injection random numbers.  Trajectories
decorrelate, so the comparison is statistical: energy histories, population, peak potential and
the hole position found by the reference's matched filter."""
import numpy as np
import pytest

from oracle import loop as oloop, workloads as wl

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_hole_run_statistics_agree():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    from oracle import cpu as ocpu
    kind = "reference" if ocpu.have("reference") else "port"
    n, n_cells, steps = 100_000, 512, 400
    cfg = SimConfig(n_steps=steps, n_particles=n, n_cells=n_cells, periodic_particles=False, track_hole=True,
                    dimple_height=0.95, dimple_velocity_width=0.3, dimple_spatial_width=4.0, seed=384)
    sim = Simulation(cfg)
    sim.init_synthetic(perturb=0.0)
    ions = sim.ions.particles[:, :n].cpu().numpy()
    electrons = sim.electrons.particles[:, :n].cpu().numpy()
    cpu = oloop.CpuSimulation(n, n_cells, False, kind=kind)
    cpu.load_particles(ions, electrons)
    sim.initialize()
    cpu.initialize()
    k_ = ocpu.load(kind)
    fine = np.linspace(-3, 3, 1001).astype(np.float32)
    hist = {"g_phi": [], "c_phi": [], "g_ke": [], "c_ke": [], "g_hole": [], "c_hole": [], "g_n": [], "c_n": []}
    for k in range(steps):
        sim.step()
        cpu.step()
        if k % 20 == 19:
            hist["g_phi"].append(float(sim.potential.max()))
            hist["c_phi"].append(float(cpu.potential.max()))
            e = sim.energies()
            hist["g_ke"].append(e["ke_electrons"])
            alive = cpu.electrons.particles[0] < 5.9
            v = cpu.electrons.particles[1][alive].astype(np.float64)
            hist["c_ke"].append(0.5 / wl.MASS_RATIO * float((v * v).sum()) / cpu.bg)
            hist["g_n"].append(sim.n_active()["electrons"])
            hist["c_n"].append(int(alive.sum()))
            sig = np.gradient(cpu.potential, cpu.dz).astype(np.float32)   # infMagSim_script.py:746-753
            hist["c_hole"].append(float(fine[np.argmax(k_.electric_field_filter(cpu.grid, sig, np.float32(0.5), fine))]))
            hist["g_hole"].append(float(sim.hole_relative_positions[k].item()))
    H = {a: np.array(b) for a, b in hist.items()}
    # electron kinetic energy (within 3 %: wall losses and injection are stochastic) and population (1 %)
    assert np.all(np.abs(H["g_ke"] / H["c_ke"] - 1) < 0.03), (H["g_ke"], H["c_ke"])
    assert np.all(np.abs(H["g_n"] / H["c_n"] - 1) < 0.01)
    # the seeded hole is a positive potential bump near the centre: same depth (10 %) ...
    late = slice(4, None)
    assert H["c_phi"][late].mean() > 0.05
    assert abs(H["g_phi"][late].mean() / H["c_phi"][late].mean() - 1) < 0.10, (H["g_phi"], H["c_phi"])
    # ... and the matched filter finds it at the same place (within 0.15 = a bit more than a Debye length)
    assert np.median(np.abs(H["g_hole"][late] - H["c_hole"][late])) < 0.15, (H["g_hole"], H["c_hole"])
    assert sim.ctx.status() == 0
