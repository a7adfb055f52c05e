"""Python-3 restatement of the reference driver's step loop on the CPU arms -- TEST
INFRASTRUCTURE and CPU-baseline runner, never part of the product path.

Follows infMagSim_script.py:256-278 (grid, seeds), :395-416 / :505-530 (stores, stacks,
background density), :649-662 (dt), :696-716 (start-up) and :756-1023 (the loop), with the
reference's default flags (pure-C mover and solver, no object, no Boltzmann electrons, moving
box off) and periodic / non-periodic selectable.  The compute calls go to oracle.cpu
("reference" = the compiled infMagSim_c.c, "port" = our C restatement).
"""
from __future__ import annotations

import math
import os
import time

import numpy as np

from . import cpu as ocpu
from . import workloads as wl


def expected_particle_injection(n_infinity, v_th, v_d, dt):  # infMagSim_script.py:414-416
    beta = v_d / (np.sqrt(2.) * v_th)
    return 2 * n_infinity * v_th / math.sqrt(2. * math.pi) * np.exp(-np.power(beta, 2.)) * dt \
        + n_infinity * v_d * math.erf(beta) * dt


class CpuSimulation:
    def __init__(self, n_particles, n_cells, periodic, kind="reference", n_ranks=1, rank=0, seed=384,
                 mass_ratio=wl.MASS_RATIO, step_size=wl.STEP_SIZE, sigma=1.0, storage_factor=1.25,
                 time_steps_immobile_ions=30000, v_b=0.0, periodic_potential=None):
        self.k_ = ocpu.load(kind)
        self.periodic = periodic
        self.periodic_potential = periodic if periodic_potential is None else periodic_potential  # ref :146
        self.n_particles, self.n_cells, self.n_points = n_particles, n_cells, n_cells + 1
        grid64, dz = np.linspace(wl.Z_MIN, wl.Z_MAX, num=n_cells + 1, endpoint=True, retstep=True)
        self.grid, self.dz = grid64.astype(np.float32), float(dz)
        self.mass_ratio, self.sigma = mass_ratio, sigma
        self.v_th_i, self.v_th_e = 1.0, math.sqrt(mass_ratio)
        self.dt = step_size * wl.DEBYE_LENGTH / self.v_th_e
        self.bg = n_particles * n_ranks * self.dz / (wl.Z_MAX - wl.Z_MIN)
        self.seed = seed + rank ** 3
        self.k_.seedgenrand_c(self.seed)
        self.host_rng = np.random.RandomState(self.seed % (2 ** 32))
        self.sampler = wl.HostUniformSampler(self.seed)
        self.object_mask = np.zeros_like(self.grid)
        self.potential = np.zeros_like(self.grid)
        self.ion_density = np.zeros_like(self.grid)
        self.electron_density = np.zeros_like(self.grid)
        self.charge_density = np.zeros_like(self.grid)
        self.time_steps_immobile_ions = time_steps_immobile_ions
        self.storage_factor = storage_factor
        self.v_b, self.a_b = v_b, 0.0
        self.k, self.t = 0, 0.0
        self.ions = self.electrons = None

    def load_particles(self, ions_xv, electrons_xv):
        storage = int(math.ceil(self.storage_factor * self.n_particles / 4.0)) * 4
        out = []
        for xv, qm, v_th in ((ions_xv, 1.0, self.v_th_i / math.sqrt(self.sigma)),
                             (electrons_xv, -self.mass_ratio, self.v_th_e)):
            n = xv.shape[1]
            p = np.empty((2, storage), np.float32)
            p[:, :n] = xv
            p[0, n:] = np.float32(2 * wl.Z_MAX)
            p[1, n:] = 0
            stack = -np.ones(storage, np.int32)
            stack[:storage - n] = np.arange(storage - 1, n - 1, -1, dtype=np.int32)
            out.append(wl.Species(p, np.zeros((1, storage), np.int32), stack, [storage - n - 1], [n - 1], qm, v_th))
        self.ions, self.electrons = out

    def init_synthetic(self, seed=0, perturb=0.01):
        rng = np.random.default_rng(seed)
        n = self.n_particles
        xv = []
        for v_th in (self.v_th_i / math.sqrt(self.sigma), self.v_th_e):
            x = rng.uniform(wl.Z_MIN, wl.Z_MAX, size=n)
            x = x + perturb * np.sin(2.0 * np.pi * x / (wl.Z_MAX - wl.Z_MIN))
            xv.append(np.stack([x.astype(np.float32), (rng.standard_normal(n) * v_th).astype(np.float32)]))
        self.load_particles(*xv)

    def initialize(self):  # infMagSim_script.py:696-716
        for s, d in ((self.ions, self.ion_density), (self.electrons, self.electron_density)):
            self.k_.accumulate_density(self.grid, self.object_mask, self.bg, s.largest_index, s.particles, d,
                                       s.empty_slots, s.current_empty_slot, periodic_particles=self.periodic)
        for s in (self.ions, self.electrons):
            self.k_.initialize_mover(self.grid, self.object_mask, self.potential, self.dt, s.charge_to_mass,
                                     s.largest_index, s.particles, s.empty_slots, s.current_empty_slot,
                                     v_b=self.v_b, a_b=0.0, periodic_particles=self.periodic)

    def field_step(self):
        """:763-779, :807, :831-836 (single rank: the all-reduce is the identity)."""
        k = self.k
        for d, fix in ((self.ion_density, self.periodic or k <= self.time_steps_immobile_ions),
                       (self.electron_density, True)):
            if not fix:
                continue
            if self.periodic:
                d[0] += d[-1]
                d[-1] = d[0]
            else:
                d[0] *= 2.
                d[-1] *= 2.
        np.subtract(self.ion_density, self.electron_density, out=self.charge_density)
        self.k_.poisson_solve(self.grid, self.object_mask, self.charge_density, wl.DEBYE_LENGTH, self.potential,
                              object_potential=-3., object_transparency=1., boltzmann_electrons=False,
                              periodic_potential=self.periodic_potential)

    def push_step(self):
        """:929-950."""
        k = self.k
        s = self.ions
        mobile = k <= self.time_steps_immobile_ions
        self.k_.move_particles(self.grid, self.object_mask, self.potential, self.dt if mobile else 0.,
                               s.charge_to_mass, self.bg, s.largest_index, s.particles, self.ion_density,
                               s.empty_slots, s.current_empty_slot, a_b=self.a_b, periodic_particles=self.periodic)
        if not mobile:
            self.ion_density[:] = 1.0
        s = self.electrons
        self.k_.move_particles(self.grid, self.object_mask, self.potential, self.dt, s.charge_to_mass, self.bg,
                               s.largest_index, s.particles, self.electron_density, s.empty_slots,
                               s.current_empty_slot, a_b=self.a_b, periodic_particles=self.periodic)

    def injection_counts(self):
        """:965-975, :1005-1009."""
        L = wl.Z_MAX - wl.Z_MIN
        e_i = expected_particle_injection(self.n_particles / L, self.v_th_i / np.sqrt(self.sigma), self.v_b, self.dt)
        n_i = int(e_i)
        if (e_i - n_i) > self.host_rng.rand():
            n_i += 1
        e_e = expected_particle_injection(self.n_particles / L, self.v_th_e, self.v_b, self.dt)
        n_e = int(e_e)
        if (e_e - n_e) > self.host_rng.rand():
            n_e += 1
        return n_i, n_e

    def inject_step(self, counts=None, velocities=None, uniforms=None):
        """:986-1019.  With velocities/uniforms given (dict per species) the random inputs are the
        caller's (parity tests); otherwise the arm's own sampler + the host sampler are used."""
        if self.periodic:
            return (0, 0)
        n_i, n_e = counts if counts is not None else self.injection_counts()
        k = self.k
        for name, s, n, v_th, den in (("ions", self.ions, n_i, self.v_th_i / math.sqrt(self.sigma), self.ion_density),
                                      ("electrons", self.electrons, n_e, self.v_th_e, self.electron_density)):
            if n <= 0 or (name == "ions" and k > self.time_steps_immobile_ions):
                continue
            if velocities is not None:
                self.k_.inject_given(n, self.grid, self.dt, self.bg, velocities[name], uniforms[name], self.a_b, k,
                                     s.tags, s.particles, s.empty_slots, s.current_empty_slot, s.largest_index, den)
            else:
                self.k_.inject_particles(n, self.grid, self.dt, v_th, self.bg, self.sampler, self.v_b, self.a_b, k,
                                         s.tags, s.particles, s.empty_slots, s.current_empty_slot, s.largest_index,
                                         den)
        return n_i, n_e

    def step(self):
        self.field_step()
        self.push_step()
        self.inject_step()
        self.t += self.dt
        self.k += 1

    def n_active(self):
        return {n: int((s.particles[0] < 0.99 * 2 * wl.Z_MAX).sum()) for n, s in
                (("ions", self.ions), ("electrons", self.electrons))}


# ---- CPU baseline: the reference parallelises as one MPI rank per core; mpirun
# is absent, so P independent worker processes each run one shard (the 16 KB density reduction is
# negligible and omitted), all started together and timed by the slowest. -------------------------

def _worker(args):
    kind, n_particles, n_cells, periodic, steps, warmup, rank, n_ranks, periodic_potential = args
    sim = CpuSimulation(n_particles, n_cells, periodic, kind=kind, n_ranks=n_ranks, rank=rank,
                        periodic_potential=periodic_potential)
    sim.init_synthetic(seed=rank)
    sim.initialize()
    for _ in range(warmup):
        sim.step()
    t0 = time.perf_counter()
    n_steps_done = 0
    for _ in range(steps):
        sim.step()
        n_steps_done += 1
    dt = time.perf_counter() - t0
    act = sim.n_active()
    return dt, (act["ions"] + act["electrons"]) * n_steps_done


def timed_cpu_sample(kind, n_particles, n_cells, periodic, steps, warmup=1, processes=None, periodic_potential=None):
    """Returns dict(value=particle-steps/s over all processes, cores=P, seconds=max over workers)."""
    import multiprocessing as mp
    P = processes or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    jobs = [(kind, n_particles, n_cells, periodic, steps, warmup, r, P, periodic_potential) for r in range(P)]
    if P == 1:
        res = [_worker(jobs[0])]
    else:
        with ctx.Pool(P) as pool:
            res = pool.map(_worker, jobs)
    seconds = max(r[0] for r in res)
    work = sum(r[1] for r in res)
    return dict(value=work / seconds, cores=P, seconds=seconds, particle_steps=work)
