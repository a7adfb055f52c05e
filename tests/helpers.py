"""Shared helpers for the parity tests (numpy side)."""
import numpy as np

from oracle import workloads as wl


def geom(grid):
    z_lo, z_hi = np.float32(grid[0]), np.float32(grid[-1])
    dz = np.float32((z_hi - z_lo) / np.float32(len(grid) - 1))
    return z_lo, z_hi, dz


def deposit_f64(grid, x, bg, periodic):
    """Exact (float64-accumulated) sum of the reference's own float32 deposit weights
    (infMagSim_c.c:219-225) for final positions x: what the density would be without any
    float32 accumulation error.  Parked / out-of-bounds slots contribute nothing."""
    z_lo, z_hi, dz = geom(grid)
    x = x.astype(np.float32)
    if periodic:
        alive = x.astype(np.float64) < 0.99 * float(np.float32(2 * z_hi))
    else:
        alive = (x > np.float32(z_lo + np.float32(1e-6))) & (x < np.float32(z_hi - np.float32(1e-6)))
    xa = x[alive]
    left = ((xa - z_lo) / dz).astype(np.int32)
    if periodic:
        left[left > len(grid) - 2] = 0
    frac = np.fmod(xa, dz) / dz
    share = np.where(xa > 0, frac, (1.0 + frac.astype(np.float64)).astype(np.float32)).astype(np.float32)
    w0 = (share / np.float32(bg)).astype(np.float64)
    w1 = ((np.float32(1) - share) / np.float32(bg)).astype(np.float64)
    d = np.zeros(len(grid), np.float64)
    np.add.at(d, left, w0)
    np.add.at(d, left + 1, w1)
    return d, int(alive.sum())


def to_dev(sp, torch, device="cuda"):
    """Species (numpy) -> dict of device tensors."""
    return dict(particles=torch.from_numpy(sp.particles.copy()).to(device),
                tags=torch.from_numpy(sp.tags.copy()).to(device),
                empty_slots=torch.from_numpy(sp.empty_slots.copy()).to(device))


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.dtype == np.float32:
        a, b = a.view(np.uint32), b.view(np.uint32)
    bad = np.flatnonzero(a.reshape(-1) != b.reshape(-1))
    assert bad.size == 0, f"{what}: {bad.size} mismatches, first at {bad[:5]}"


def run_gpu_hole_history(n, steps, record_every=100, n_cells=512, seed=384, device=None, progress=None):
    """The long hole-tracking run (BASELINE configs[4]) on the GPU: script-default parameters from the seeded
    arrays of oracle.workloads.make_hole_plasma (the same arrays tests/golden/make_longrun_golden.py gives the
    oracle), recording every ``record_every`` steps what SURVEY.md s.8d lists: max potential, tracked hole
    position, kinetic energies per species, field energy, live populations.  Returns (history dict of numpy
    arrays, Simulation)."""
    import time

    import torch

    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    cfg = SimConfig(n_steps=steps, n_particles=n, n_cells=n_cells, periodic_particles=False, track_hole=True, seed=seed)
    sim = Simulation(cfg, device=device)
    ions, electrons = wl.make_hole_plasma(n, seed=seed)
    sim.load_particles(ions, electrons)
    sim.initialize()
    rec = {k: [] for k in ("step", "max_phi", "hole", "ke_i", "ke_e", "fe", "n_i", "n_e")}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(record_every - 1, steps, record_every):
        sim.run(record_every)     # espic_pic_run: the whole stretch from one host call
        if True:
            e = sim.energies()
            act = sim.n_active()
            rec["step"].append(k)
            rec["max_phi"].append(float(sim.potential.max()))
            rec["hole"].append(float(sim.hole_relative_positions[k]))
            rec["ke_i"].append(e["ke_ions"])
            rec["ke_e"].append(e["ke_electrons"])
            rec["fe"].append(e["fe"])
            rec["n_i"].append(act["ions"])
            rec["n_e"].append(act["electrons"])
            if progress and (k + 1) % (100 * record_every) == 0:
                progress(k + 1, time.perf_counter() - t0)
    torch.cuda.synchronize()
    hist = {k: np.array(v) for k, v in rec.items()}
    hist["seconds"] = np.array(time.perf_counter() - t0)
    hist["status"] = np.array(sim.status_word())
    return hist, sim
