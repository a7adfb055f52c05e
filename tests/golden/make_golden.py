"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref.so, i.e.
/root/reference/infMagSim_c.c compiled by oracle/Makefile with the reference's flags).

    python tests/golden/make_golden.py        # needs /root/reference (build container only)

The fixtures pin the oracle (tests/test_golden.py, CPU) and the CUDA path (-m gpu) to outputs of
the reference itself; they travel with the repository, /root/reference does not.
The injection fixture comes from the reference's own inject_particles
(infMagSim_cython.py:230-285), cut unchanged out of the reference file and cythonized by
oracle/build_inject_ref.py (oracle/_ref/inject_ref*.so), drawing its velocities with the
reference's own draw_velocities_c.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cpu, workloads as wl  # noqa: E402


def main():
    ref = cpu.load("reference")
    # ---- mover: 3 steps, walls and periodic, n_cells=500 (dz inexact) and 64
    for periodic in (False, True):
        for n_cells in (64, 500):
            grid = wl.make_grid(n_cells)
            phi = wl.smooth_potential(grid, amplitude=2.0, seed=n_cells)
            _, sp = wl.make_plasma(3000, n_cells, seed=100 + n_cells + int(periodic), holes=0.03)
            sp.particles[0, :6] = np.array([-3.0, 3.0, -3.0000005, 2.9999995, 3.5, -7.0], np.float32)
            before = sp.copy()
            den = np.zeros_like(grid)
            dens = []
            for step in range(3):
                ref.move_particles(grid, np.zeros_like(grid), phi, np.float32(wl.timestep()), sp.charge_to_mass,
                                   np.float32(25.0), sp.largest_index, sp.particles, den, sp.empty_slots,
                                   sp.current_empty_slot, a_b=np.float32(0.25 * step), periodic_particles=periodic)
                dens.append(den.copy())
            np.savez_compressed(os.path.join(HERE, f"mover_{'periodic' if periodic else 'walls'}_{n_cells}.npz"),
                                grid=grid, potential=phi, dt=np.float32(wl.timestep()), qm=np.float32(sp.charge_to_mass),
                                bg=np.float32(25.0), particles_in=before.particles, stack_in=before.empty_slots,
                                top_in=before.current_empty_slot[0], last=before.largest_index[0],
                                particles_out=sp.particles, stack_out=sp.empty_slots, top_out=sp.current_empty_slot[0],
                                densities=np.array(dens))
    # ---- Poisson: Robin and periodic, with and without object / Boltzmann
    rng = np.random.default_rng(7)
    for n_points in (65, 513):
        grid = wl.make_grid(n_points - 1)
        charge = (0.05 * rng.standard_normal(n_points)).astype(np.float32)
        mask = np.zeros_like(grid)
        mask[n_points // 3:n_points // 3 + 9] = 1.0
        mask[n_points // 3 - 1], mask[n_points // 3 + 9] = 0.3, 0.6
        out = {}
        for name, kw in (("robin", dict()), ("periodic", dict(periodic_potential=True)),
                         ("object", dict(object_potential=-3.0, object_transparency=0.35, use_mask=True)),
                         ("boltzmann", dict(boltzmann_electrons=True))):
            pot = (0.1 * np.sin(grid)).astype(np.float32)
            use_mask = kw.pop("use_mask", False)
            ref.poisson_solve(grid, mask if use_mask else np.zeros_like(grid), charge, np.float32(0.125), pot, **kw)
            out[name] = pot
        np.savez_compressed(os.path.join(HERE, f"poisson_{n_points}.npz"), grid=grid, charge=charge, mask=mask,
                            potential_in=(0.1 * np.sin(grid)).astype(np.float32), **out)
    # ---- sampler stream + injection
    ref.seedgenrand_c(384)
    vel = ref.draw_velocities(400, np.float32(42.85), np.float32(0.0))
    ref.seedgenrand_c(385)
    vel_drift = ref.draw_velocities(400, np.float32(1.0), np.float32(0.8))
    grid = wl.make_grid(500)
    ions, _ = wl.make_plasma(2000, 500, seed=9)
    before = ions.copy()
    den0 = np.random.default_rng(1).random(len(grid)).astype(np.float32)
    den = den0.copy()
    uni = np.random.default_rng(2).random(400).astype(np.float32)
    # the reference's own routine: re-seeded with 385 it draws vel_drift again (v_th 1, v_d 0.8) before placing
    cpu.reference_inject_particles(400, grid, np.float32(0.01), np.float32(1.0), np.float32(4.0), cpu._GivenUniforms(uni),
                                   np.float32(0.8), np.float32(0.2), 17, ions.tags, ions.particles, ions.empty_slots,
                                   ions.current_empty_slot, ions.largest_index, den, seed=385)
    np.savez_compressed(os.path.join(HERE, "sampler_and_injection.npz"), vel_384_4285_0=vel, vel_385_1_08=vel_drift,
                        grid=grid, uniforms=uni, density_in=den0, density_out=den, particles_in=before.particles,
                        stack_in=before.empty_slots, top_in=before.current_empty_slot[0], last_in=before.largest_index[0],
                        particles_out=ions.particles, stack_out=ions.empty_slots, tags_out=ions.tags,
                        top_out=ions.current_empty_slot[0], last_out=ions.largest_index[0])
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
