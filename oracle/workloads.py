"""Seeded synthetic inputs for tests and bench.  numpy only, no reference import.

Layout follows the reference driver (infMagSim_script.py:256-278, 395-416, 485-489, 505-530,
614-618, 649-662): float32 grid on [-3, 3]; particle store float32 [2, S] (row 0 positions,
row 1 velocities) with unused slots parked at 2*z_max; int32 LIFO stack of free slots filled
in reverse so that the lowest free slot is popped first; one-element lists for the stack top
and the largest used index.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

Z_MIN, Z_MAX = -3.0, 3.0
DEBYE_LENGTH = 0.125  # infMagSim_script.py:268
MASS_RATIO = 1836.0
STEP_SIZE = 0.3       # infMagSim_script.py:24


def make_grid(n_cells: int) -> np.ndarray:
    """infMagSim_script.py:256-266."""
    return np.linspace(Z_MIN, Z_MAX, num=n_cells + 1, endpoint=True).astype(np.float32)


def timestep(step_size: float = STEP_SIZE, mass_ratio: float = MASS_RATIO) -> float:
    """dt = step_size*debye_length/v_th_e with v_th_e = sqrt(mass_ratio) (:649-662)."""
    return step_size * DEBYE_LENGTH / math.sqrt(mass_ratio)


def background_density(n_per_rank: int, n_ranks: int, n_cells: int) -> float:
    """n*P*dz/(z_max-z_min) (:402, :510)."""
    dz = (Z_MAX - Z_MIN) / n_cells
    return n_per_rank * n_ranks * dz / (Z_MAX - Z_MIN)


@dataclass
class Species:
    particles: np.ndarray          # float32 [2, S]
    tags: np.ndarray               # int32 [1, S]  (particles_hist / injection history)
    empty_slots: np.ndarray        # int32 [S]
    current_empty_slot: list = field(default_factory=lambda: [-1])
    largest_index: list = field(default_factory=lambda: [-1])
    charge_to_mass: float = 1.0
    v_th: float = 1.0

    def copy(self) -> "Species":
        return Species(self.particles.copy(), self.tags.copy(), self.empty_slots.copy(),
                       list(self.current_empty_slot), list(self.largest_index),
                       self.charge_to_mass, self.v_th)


def make_species(n: int, storage: int, rng: np.random.Generator, v_th: float,
                 charge_to_mass: float, perturb: float = 0.01, holes: float = 0.0) -> Species:
    """n active particles in slots [0, n), uniform in x with a small sinusoidal displacement so
    the fields are non-trivial, Maxwellian in v.  ``holes`` parks that fraction of the first n
    slots (and lists them on the stack) to exercise the scan over dead slots."""
    assert storage >= n
    p = np.empty((2, storage), dtype=np.float32)
    x = rng.uniform(Z_MIN, Z_MAX, size=n)
    x = x + perturb * np.sin(2.0 * np.pi * x / (Z_MAX - Z_MIN))
    p[0, :n] = x.astype(np.float32)
    p[1, :n] = (rng.standard_normal(n) * v_th).astype(np.float32)
    p[0, n:] = np.float32(2 * Z_MAX)
    p[1, n:] = 0.0
    stack = -np.ones(storage, dtype=np.int32)
    free = np.arange(storage - 1, n - 1, -1, dtype=np.int32)  # :487-489
    if holes > 0.0:
        dead = np.sort(rng.choice(n, size=int(holes * n), replace=False)).astype(np.int32)
        p[0, dead] = np.float32(2 * Z_MAX)
        free = np.concatenate([free, dead])
    stack[:len(free)] = free
    return Species(p, np.zeros((1, storage), dtype=np.int32), stack, [len(free) - 1], [n - 1],
                   charge_to_mass, v_th)


def make_plasma(n_per_species: int, n_cells: int, seed: int = 0, storage_factor: float = 1.25,
                perturb: float = 0.01, mass_ratio: float = MASS_RATIO, holes: float = 0.0):
    """Ions (q/m=+1, v_th=1) and electrons (q/m=-mass_ratio, v_th=sqrt(mass_ratio))."""
    rng = np.random.default_rng(seed)
    storage = int(math.ceil(n_per_species * storage_factor / 4.0)) * 4
    ions = make_species(n_per_species, storage, rng, 1.0, 1.0, perturb, holes)
    electrons = make_species(n_per_species, storage, rng, math.sqrt(mass_ratio), -mass_ratio,
                             perturb, holes)
    return ions, electrons


def smooth_potential(grid: np.ndarray, amplitude: float = 0.5, seed: int = 1) -> np.ndarray:
    """A non-trivial potential for single-call mover tests."""
    rng = np.random.default_rng(seed)
    z = grid.astype(np.float64)
    phi = amplitude * np.sin(2 * np.pi * z / 6.0) + 0.2 * amplitude * np.cos(5 * np.pi * z / 6.0)
    phi += 0.01 * amplitude * rng.standard_normal(len(grid))
    return phi.astype(np.float32)


class HostUniformSampler:
    """Stands in for the driver's uniform_2d_sampler (infMagSim_script.py:296-335) in its
    default configuration: ``get(n)`` returns an [n, 2] float64 array of U(0,1)."""
    shared_seed = False

    def __init__(self, seed: int):
        self.rng = np.random.RandomState(seed)

    def get(self, n: int) -> np.ndarray:
        return self.rng.rand(int(n), 2)


def make_hole_plasma(n: int, seed: int = 384, dimple_height: float = 0.95, dimple_velocity: float = 0.0,
                     dimple_velocity_width: float = 0.3, dimple_spatial_width: float = 4.0,
                     mass_ratio: float = MASS_RATIO):
    """Initial particle arrays of the reference's electron-hole run (script defaults, C1/C5 of SURVEY.md
    s.8d): uniform positions, Maxwellian velocities, and the phase-space "dimple" carved out of the
    electrons by rejection -- a sample is redrawn (velocity only) with probability
    height*exp(-(v-mu)^2/2sig^2)/(1+0.0002*exp(|x|/Lambda)) (infMagSim_script.py:518-520, 602-606).
    Returns (ions_xv, electrons_xv), float32 [2, n] each: identical initial arrays for both sides of a
    parity run."""
    rng = np.random.default_rng(seed)
    v_th_e = math.sqrt(mass_ratio)
    out = []
    for v_th, dimple in ((1.0, False), (v_th_e, True)):
        x = rng.uniform(Z_MIN, Z_MAX, n)
        v = rng.standard_normal(n) * v_th
        if dimple:
            sig = v_th_e * dimple_velocity_width
            lam = dimple_spatial_width * DEBYE_LENGTH
            todo = np.arange(n)
            while todo.size:
                d = dimple_height * np.exp(-(v[todo] - dimple_velocity) ** 2 / (2 * sig ** 2)) \
                    / (1. + 0.0002 * np.exp(np.abs(x[todo]) / lam))
                todo = todo[rng.random(todo.size) < d]
                v[todo] = rng.standard_normal(todo.size) * v_th
        out.append(np.stack([x.astype(np.float32), v.astype(np.float32)]))
    return out
