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
    assert sim.status_word() == 0


def test_configs4_hundred_thousand_steps_against_oracle_history():
    """BASELINE configs[4] at its stated length: 1e5 steps of the script-default hole run on the GPU (walls,
    injection, dimple, hole tracking every step, ions frozen after step 30000 as in the reference,
    infMagSim_script.py:929-939), diagnostics every 100 steps.
    (1) the whole run: no status bit, populations and energies bounded, the hole stays tracked inside the box;
    (2) the first 1e4 steps against the ORACLE's history from the same initial arrays
    (tests/golden/longrun_c1_n1e5.npz: oracle/loop.py over the compiled reference,
    infMagSim_script.py:756-1027) -- the two sides draw their injections from different generators, so the
    agreement is statistical: kinetic energies, populations, hole depth and hole position."""
    import os
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests.helpers import run_gpu_hole_history
    n, steps = 100_000, 100_000
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "longrun_c1_n1e5.npz"))
    assert int(gold["n"]) == n and int(gold["record_every"]) == 100
    hist, sim = run_gpu_hole_history(n, steps)
    assert int(hist["status"]) == 0
    assert len(hist["step"]) == steps // 100
    # ---- (1) the whole 1e5-step run
    for key in ("max_phi", "hole", "ke_i", "ke_e", "fe"):
        assert np.all(np.isfinite(hist[key])), key
    # the violent start expels ~5 % of both species (the oracle's populations dip to 0.948 / 0.951 of n around
    # step 1800 and recover as the walls refill the box)
    assert np.all(np.abs(hist["n_e"] / n - 1) < 0.08), (hist["n_e"].min(), hist["n_e"].max())
    assert np.all(np.abs(hist["n_i"] / n - 1) < 0.08), (hist["n_i"].min(), hist["n_i"].max())
    assert np.all(np.abs(hist["n_e"][hist["step"] > 20000] / n - 1) < 0.01)
    frozen = hist["step"] > sim.cfg.time_steps_immobile_ions + 100
    assert np.all(hist["n_i"][frozen] == hist["n_i"][frozen][0])          # frozen ions are neither pushed nor injected
    assert np.all(hist["ke_i"][frozen] == hist["ke_i"][frozen][0])
    # the dimple start is violent (max potential 1.4, electron KE 33 % above thermal: the rejection removed slow
    # electrons); it relaxes within ~3000 steps as the walls replace the electron population.  From then on: no
    # secular heating -- KE per species stays at the thermal value n/(2 bg) = 256 within 4 %, to the end of the run
    ke_e = hist["ke_e"]
    settled = hist["step"] > 5000
    thermal = n / (2 * sim.background_density)
    assert np.all(np.abs(ke_e[settled] / thermal - 1) < 0.04), (ke_e[settled].min(), ke_e[settled].max(), thermal)
    assert np.all(np.abs(hist["ke_i"][settled] / thermal - 1) < 0.10)
    assert np.all(hist["fe"] < 0.05 * ke_e)                                # field energy stays a small fraction
    assert np.all(np.abs(hist["hole"]) <= 3.0)
    # ---- (2) the first 1e4 steps against the oracle history.  Tolerances: ~1.5x the reference's OWN spread between
    # two runs that differ only in the injection seed (measured with oracle/loop.py, n = 1e5: electron KE up to 7.5 %
    # apart around step 2000, populations 0.7 %, ion KE 2 %, mean hole depth of the first 1000 steps 6 %; the hole
    # trajectories coincide for ~400 steps, then decorrelate as the hole leaves the box)
    m = len(gold["step"])
    np.testing.assert_array_equal(gold["step"], hist["step"][:m])
    assert np.all(np.abs(hist["ke_e"][:m] / gold["ke_e"] - 1) < 0.11)
    assert np.all(np.abs(hist["ke_i"][:m] / gold["ke_i"] - 1) < 0.04)
    assert np.all(np.abs(hist["n_e"][:m] / gold["n_e"] - 1) < 0.015)
    assert np.all(np.abs(hist["n_i"][:m] / gold["n_i"] - 1) < 0.01)
    assert abs(hist["max_phi"][:10].mean() / gold["max_phi"][:10].mean() - 1) < 0.15      # hole depth while it lives
    # same hole, same early trajectory: the first 200 steps coincide; at steps 300 / 400 eleven REFERENCE runs that
    # differ only in the injection seed scatter over [-0.53, -0.23] / [-0.62, -0.01], and from step 500 on the hole
    # runs off to either side (their mean positions over the first 1000 steps range from -1.5 to +0.55)
    assert np.all(np.abs(hist["hole"][:2] - gold["hole"][:2]) < 0.15)
    assert np.all(np.abs(hist["hole"][2:4] - gold["hole"][2:4]) < 0.7)
    late = slice(m // 2, m)                                                               # after the hole has left
    assert abs(hist["ke_e"][:m][late].mean() / gold["ke_e"][late].mean() - 1) < 0.02
    assert abs(hist["max_phi"][:m][late].mean()) < 0.25 and abs(gold["max_phi"][late].mean()) < 0.25
