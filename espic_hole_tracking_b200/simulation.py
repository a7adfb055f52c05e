"""Device-resident mirror of the reference's simulation driver loop (infMagSim_script.py).

Keeps the reference driver's structure -- config keys of ``input.txt`` (:22-79), the particle
stores / free-slot stacks / one-element scalars (:395-416, 485-489, 505-530, 614-618), the
start-up sequence accumulate_density -> initialize_mover (:696-716) and the per-step order of
the loop (:756-1023): density all-reduce, end-node fix, charge density, field solve, hole
tracking, ion push, electron push, wall injection -- and calls the operator layer for every one
of them.  All state lives in HBM; a step enqueues kernels and never synchronises with the host.

Particles are sharded over ranks exactly as the reference shards them over MPI ranks: each
rank owns private particle stores and a full copy of the grids, deposits with
background_density = n*P*dz/(z_max-z_min) (:402, :510) and the only per-step exchange is the
sum all-reduce of the density grids (:763, :773), here ONE collective over the concatenated
[2, n_points] ion+electron buffer.

The moving-box controller (:917-928, off by default as in the reference) and the ion velocity
offset of the background acceleration (:852-862) are host scalar bookkeeping, kept in
``controller.BoxController``; with the controller on, a step reads the tracked hole position back
(4 bytes) between the field solve and the pushes.

Out of scope: the absorbing-object geometry helper (:192-246; object_mask stays zero as in every config),
counter-streaming beams, Boltzmann/quasineutral/prescribed-potential modes, plotting and .npz
output.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import distributed as dd
from . import operators as ops
from .controller import BoxController, PHASE2_POTENTIAL
from ._lib import Context, Species, StepArgs


@dataclass
class SimConfig:
    # ---- input.txt keys (infMagSim_script.py:22-79); names as in the file ----
    n_steps: int = 100
    n_particles: int = 100_000          # per rank, per species
    n_cells: int = 512
    step_size: float = 0.3
    v_d_i: float = 0.0
    v_b_0: float = 0.0
    dimple_velocity: float = 0.0
    dimple_height: float = 0.95
    dimple_velocity_width: float = 0.3
    dimple_spatial_width: float = 4.0
    mass_ratio: float = 1836.0
    sigma: float = 1.0                   # "Te/Ti"
    hole_tracking_end_step: int = 0
    w_ion_perturbation: float = 0.0
    amplitude_perturbation: float = 0.0
    background_acceleration: float = 0.0
    hole_pushing_start_step: int = 0
    hole_pushing_end_step: int = 0
    density_growth_start_step: int = 0   # "background_density_growth_start_step"
    density_growth_rate: float = 0.0
    # ---- module-level flags the reference hard-codes (:133-189, 256-273, 396-406, 720-744) ----
    periodic_particles: bool = False
    # None = tied to periodic_particles as in the reference (:146).  NB the reference's periodic solve
    # (Sherman-Morrison with two bugs, reproduced bit-compatibly) returns a sawtooth potential on grids of
    # ~1000+ points that heats the plasma explosively in the reference itself; periodic
    # particles with periodic_potential=False (Robin ends, ref :407-409, 426-428) is stable.
    periodic_potential: bool | None = None
    time_steps_immobile_ions: int = 30000
    time_steps_immobile_electrons: int = 0
    create_electron_dimple: bool = True
    moving_box_simulation: bool = False   # :141: a frame that follows the hole (controller.BoxController)
    debye_length: float = 0.125
    z_min: float = -3.0
    z_max: float = 3.0
    extra_storage_factor: float = 1.25   # the reference uses 10 (:401); 1.25 is ample
    n_tracking_mesh_points: int = 1001
    initial_transient_steps: int = 80
    track_hole: bool = True
    phase_space_histograms: bool = False
    seed: int = 384
    # Coarse cell sort of the stores (no reference counterpart; physically a no-op).  Steps between
    # sorts per species: None = automatic (never when the whole grid fits the mover's shared-memory
    # tables; otherwise often enough that thermal drift keeps a segment inside its window), 0 = never.
    sort_interval_ions: int | None = None
    sort_interval_electrons: int | None = None

    _KEYS = {"n_steps": int, "n_particles": int, "step_size": float, "v_d_i": float,
             "dimple_velocity": float, "dimple_height": float, "dimple_velocity_width": float,
             "dimple_spatial_width": float, "mass_ratio": float, "Te/Ti": float, "v_b_0": float,
             "n_cells": int, "hole_tracking_end_step": int, "w_ion_perturbation": float,
             "amplitude_perturbation": float, "background_acceleration": float,
             "hole_pushing_start_step": int, "hole_pushing_end_step": int,
             "background_density_growth_start_step": int, "density_growth_rate": float}

    @classmethod
    def from_input_file(cls, path: str, **overrides) -> "SimConfig":
        """Whitespace-separated ``key value`` lines, the reference's format (:33-79)."""
        rename = {"Te/Ti": "sigma", "background_density_growth_start_step": "density_growth_start_step"}
        kw = {}
        with open(path) as fh:
            for line in fh:
                parts = line.split()
                if len(parts) >= 2 and parts[0] in cls._KEYS:
                    kw[rename.get(parts[0], parts[0])] = cls._KEYS[parts[0]](parts[1])
        kw.update(overrides)
        return cls(**kw)


def expected_particle_injection(n_infinity, v_th, v_d, dt):
    """Wall flux of a drifting Maxwellian through both ends (infMagSim_script.py:414-416)."""
    beta = v_d / (math.sqrt(2.) * v_th)
    return (2 * n_infinity * v_th / math.sqrt(2. * math.pi) * math.exp(-beta ** 2) * dt
            + n_infinity * v_d * math.erf(beta) * dt)


class DeviceUniformSampler:
    """Partial-step fractions drawn on the device (same role as uniform_2d_sampler.get(n)[:,0])."""
    shared_seed = False

    def __init__(self, ctx: Context, device):
        self.ctx, self.device = ctx, device

    def get_device(self, n):
        return ops.device_uniforms(n, device=self.device, ctx=self.ctx)


class SpeciesStore:
    """One species: float32 [2,S] store, int32 [1,S] tags, int32 [S] LIFO stack, device scalars."""

    def __init__(self, n_active: int, storage: int, charge_to_mass: float, v_th: float, parked: float, device):
        self.n0 = n_active
        self.storage = storage
        self.charge_to_mass = charge_to_mass
        self.v_th = v_th
        self.particles = torch.empty((2, storage), dtype=torch.float32, device=device)
        self.particles[0, n_active:] = parked            # :484-485
        self.particles[1, n_active:] = 0.0
        self.tags = torch.zeros((1, storage), dtype=torch.int32, device=device)
        # free slots listed in reverse so the lowest is popped first (:486-489)
        self.empty_slots = torch.full((storage,), -1, dtype=torch.int32, device=device)
        n_free = storage - n_active
        if n_free > 0:
            self.empty_slots[:n_free] = torch.arange(storage - 1, n_active - 1, -1, dtype=torch.int32, device=device)
        self.current_empty_slot = ops.DeviceInt(n_free - 1, device=device)
        self.largest_index = ops.DeviceInt(n_active - 1, device=device)


class Simulation:
    def __init__(self, cfg: SimConfig, rank: int = 0, world_size: int = 1, device=None, process_group=None,
                 injection_sampler=None, exchange: str = "auto"):
        """exchange: how the per-step density all-reduce is done for world_size > 1 --
        "p2p" (default via "auto"): fused into the field-solve kernel over NVLink peer memory
        (distributed.DensityExchange); "nccl": one torch.distributed all-reduce per step."""
        self.cfg = cfg
        self.rank, self.world = rank, world_size
        self.pg = process_group
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.ctx = ops.default_context(self.device)
        c = cfg
        n_points = c.n_cells + 1
        grid64, dz = np.linspace(c.z_min, c.z_max, num=n_points, endpoint=True, retstep=True)  # :263
        self.dz = float(dz)
        self.n_points = n_points
        self.grid = torch.from_numpy(grid64.astype(np.float32)).to(self.device)
        fine = np.linspace(c.z_min, c.z_max, num=c.n_tracking_mesh_points, endpoint=True)
        self.fine_mesh = torch.from_numpy(fine.astype(np.float32)).to(self.device)
        self.seed = dd.rank_seed(c.seed, rank)                             # :277
        ops.seedgenrand_c(self.seed, ctx=self.ctx)                         # :278
        self.host_rng = np.random.RandomState(self.seed % (2 ** 32))       # np.random.seed(seed), :313
        self.v_th_i = 1.0
        self.v_th_e = math.sqrt(c.mass_ratio)                              # :509
        self.dt = c.step_size * c.debye_length / self.v_th_e               # :660
        self.background_density = dd.background_density(c.n_particles, world_size, self.dz, c.z_min, c.z_max)
        self.parked = 2 * c.z_max                                          # :406
        storage = int(math.ceil(c.extra_storage_factor * c.n_particles / 4.0)) * 4
        self.ions = SpeciesStore(c.n_particles, storage, 1.0, self.v_th_i / math.sqrt(c.sigma), self.parked,
                                 self.device)
        self.electrons = SpeciesStore(c.n_particles, storage, -c.mass_ratio, self.v_th_e, self.parked,
                                      self.device)
        z = lambda: torch.zeros(n_points, dtype=torch.float32, device=self.device)
        self.object_mask = z()
        self.potential = z()
        self.densities = torch.zeros((2, n_points), dtype=torch.float32, device=self.device)
        self.ion_density, self.electron_density = self.densities[0], self.densities[1]
        self.charge_density = z()
        self.background_potential = torch.from_numpy(
            np.linspace(c.background_acceleration, 0., num=n_points, endpoint=True).astype(np.float32)).to(self.device)
        self.set_background_acceleration = 0 < c.hole_pushing_start_step < c.n_steps  # :149-152
        self.hole_relative_positions = torch.zeros(max(c.n_steps, 1), dtype=torch.float32, device=self.device)
        self.sampler = injection_sampler or DeviceUniformSampler(self.ctx, self.device)
        if exchange == "auto":
            exchange = "p2p" if world_size > 1 else "none"
        self.exchange_mode = exchange if world_size > 1 or exchange == "p2p" else "none"
        self.xch = None
        if self.exchange_mode == "p2p":
            self.ctx = Context(self.device.index)   # the exchange buffers belong to a context of their own
            ops.seedgenrand_c(self.seed, ctx=self.ctx)
            self.sampler = injection_sampler or DeviceUniformSampler(self.ctx, self.device)
            self.xch = dd.DensityExchange(self.ctx, n_points, world_size, rank, self.device, process_group)
        self.periodic_potential = c.periodic_particles if c.periodic_potential is None else c.periodic_potential
        # Fused deposit (espic_species.acc): the pushes and injections of a step leave their deposit in two
        # caller-owned planes of fixed-point accumulators; the next field solve (or, multi-GPU, the publication
        # of the exchange slot) converts it.  No conversion pass and no epilogue launch on the step path.
        # Not with exchange="nccl": torch's all-reduce needs the float rows.
        self.fused_deposit = self.exchange_mode in ("none", "p2p")
        self.acc = torch.zeros((2, (n_points + 64) & ~63), dtype=torch.int64, device=self.device)   # >= n_points + 1
        self._acc_pending = [False, False]
        # what the reference driver stores per step (:906-910) are the reduced, end-fixed densities the solve used.
        # With the fused deposit ion_density / electron_density hold exactly that after a step; without it
        # (exchange="nccl") the pushes overwrite them, and a Recorder asks for a snapshot (keep_solved_densities)
        self.keep_solved_densities = False
        self._solved = None
        self.k = 0
        self.t = 0.0
        self.sort_intervals = {"ions": 0, "electrons": 0}   # set by initialize() (_plan_sorting)
        self.sort_key_shift = 0
        self.v_b = c.v_b_0     # initial box velocity (:736); per step: self.ctrl.box
        self.a_b = 0.0
        self.ctrl = BoxController(c.n_steps, self.dt, self.dz, c.step_size, self.v_th_e, c.debye_length,
                                  v_b_0=c.v_b_0, moving_box=c.moving_box_simulation,
                                  hole_tracking_end_step=c.hole_tracking_end_step,
                                  background_acceleration=c.background_acceleration,
                                  hole_pushing_start_step=c.hole_pushing_start_step,
                                  hole_pushing_end_step=c.hole_pushing_end_step, n_points=self.n_points)
        self.background_potential_phase2 = torch.linspace(PHASE2_POTENTIAL, 0., self.n_points, dtype=torch.float64
                                                          ).to(torch.float32).to(self.device)         # :673
        self.hist = {}

    def _deposit_rows(self, k: int):
        """Where the pushes/injections that produce the densities entering step k deposit: the
        exchange slot of that step's parity, or the local density buffer."""
        if self.xch is not None:
            slot = self.xch.slot(k)
            return slot[0], slot[1]
        return self.ion_density, self.electron_density

    # ---- particle initialisation ----------------------------------------------------------
    def load_particles(self, ions_xv: np.ndarray, electrons_xv: np.ndarray) -> None:
        """Identical initial particle arrays on both sides of a parity test: float32 [2, n]."""
        for store, xv in ((self.ions, ions_xv), (self.electrons, electrons_xv)):
            n = xv.shape[1]
            assert n == store.n0
            store.particles[:, :n] = torch.from_numpy(np.ascontiguousarray(xv, dtype=np.float32)).to(self.device)

    def init_synthetic(self, perturb: float = 0.01) -> None:
        """Uniform positions with a small sinusoidal displacement, Maxwellian velocities, and the
        reference's phase-space "dimple" carved out of the electrons by rejection
        (infMagSim_script.py:518-520, 602-613).  Generated on the device (torch RNG)."""
        c = self.cfg
        gen = torch.Generator(device=self.device)
        gen.manual_seed(self.seed)
        L = c.z_max - c.z_min
        for store, dimple in ((self.ions, False), (self.electrons, c.create_electron_dimple and not c.periodic_particles)):
            n = store.n0
            x = torch.rand(n, generator=gen, device=self.device, dtype=torch.float64) * L + c.z_min
            x = x + perturb * torch.sin(2 * math.pi * x / L)
            v = torch.randn(n, generator=gen, device=self.device, dtype=torch.float32) * store.v_th
            if dimple:
                sig = self.v_th_e * c.dimple_velocity_width
                lam = c.dimple_spatial_width * c.debye_length
                todo = torch.arange(n, device=self.device)
                while todo.numel() > 0:   # re-draw the velocity of rejected samples (:536-613)
                    vv, xx = v[todo].double(), x[todo]
                    dim = c.dimple_height * torch.exp(-(vv - c.dimple_velocity) ** 2 / (2 * sig ** 2)) \
                        / (1. + 0.0002 * torch.exp(torch.abs(xx) / lam))
                    rej = torch.rand(todo.numel(), generator=gen, device=self.device, dtype=torch.float64) < dim
                    todo = todo[rej]
                    if todo.numel():
                        v[todo] = torch.randn(todo.numel(), generator=gen, device=self.device,
                                              dtype=torch.float32) * store.v_th
            store.particles[0, :n] = x.float()
            store.particles[1, :n] = v

    # ---- start-up (infMagSim_script.py:696-716) --------------------------------------------
    def initialize(self) -> None:
        c = self.cfg
        den_i, den_e = self._deposit_rows(0)
        for store, den in ((self.ions, den_i), (self.electrons, den_e)):
            ops.accumulate_density(self.grid, self.object_mask, self.background_density, store.largest_index,
                                   store.particles, den, store.empty_slots, store.current_empty_slot,
                                   periodic_particles=c.periodic_particles, use_pure_c_version=True, ctx=self.ctx)
        for store in (self.ions, self.electrons):
            ops.initialize_mover(self.grid, self.object_mask, self.potential, self.dt, store.charge_to_mass,
                                 store.largest_index, store.particles, store.empty_slots,
                                 store.current_empty_slot, v_b=self.v_b, a_b=0.0,
                                 periodic_particles=c.periodic_particles, use_pure_c_version=True, ctx=self.ctx)
        if self.xch is not None:
            self.xch.publish(0, 0)
        self._plan_sorting()
        if self.sort_intervals["ions"] or self.sort_intervals["electrons"]:
            for n, every in self.sort_intervals.items():
                if every:
                    self.sort_particles((n,), key_shift=self.sort_key_shift, lookahead_steps=0.5 * every)

    def _plan_sorting(self) -> None:
        """Sort intervals for the large-grid mover (csrc/mover_win.cu): a segment of a sorted store
        spreads by v*dt per step.  Stores are sorted by where the particles will be half an interval
        later, so the spread first shrinks, then grows; re-sort before ~2 thermal speeds have used
        up the window's margin on either side."""
        c = self.cfg
        window = self.ctx._lib.espic_mover_window_cells(self.ctx.handle, self.n_points)
        self.sort_key_shift = 0
        while (((self.n_points - 2) >> self.sort_key_shift) + 1) >= 512:   # one radix pass
            self.sort_key_shift += 1
        auto = {}
        for name, v_th in (("ions", self.v_th_i / math.sqrt(c.sigma)), ("electrons", self.v_th_e)):
            if window <= 0 or window >= c.n_cells:   # whole grid in shared memory, or a window that covers it
                auto[name] = 0
                continue
            margin = 0.5 * window - (1 << self.sort_key_shift)
            drift = max(v_th, abs(self.v_b)) * self.dt / self.dz          # cells per step at one thermal speed
            # measured optimum on the 65536-cell grid: 2.05 thermal speeds of margin with walls (injected particles pile
            # up unsorted in the tail of the store meanwhile; electrons every 32 steps: 26 / 28 / 30 / 32 / 34 / 36 steps
            # give 1.006 / 1.004 / 0.998 / 0.994 / 0.998 / 1.008 ms per step -- it was 2.3, every 28, while the mover
            # was slower), 1.8 in a periodic box
            reach = 1.8 if c.periodic_particles else 2.05
            auto[name] = int(min(1000, max(1, 2 * margin / (reach * drift))))
        self.sort_intervals = {
            "ions": auto["ions"] if c.sort_interval_ions is None else c.sort_interval_ions,
            "electrons": auto["electrons"] if c.sort_interval_electrons is None else c.sort_interval_electrons}

    # ---- one timestep (infMagSim_script.py:756-1023) ---------------------------------------
    def reduce_densities(self) -> None:
        if self.xch is None:
            dd.all_reduce_densities(self.densities, self.world, self.pg)   # :763, :773 in one call

    def injection_counts(self, k: int):
        """Expected wall flux with stochastic rounding (:965-975, 1005-1009)."""
        c = self.cfg
        L = c.z_max - c.z_min
        growth = math.exp(max(0, k - c.density_growth_start_step) * c.density_growth_rate * c.step_size)
        w_ion = c.w_ion_perturbation * self.v_th_e / c.debye_length * math.sqrt(1. / c.mass_ratio)   # :663
        if self.ctrl.is_static:
            v_flux, v_extra = float(self.ctrl.box[0, 0]), 0.0
        else:
            v_flux, v_extra = self.ctrl.injection_flux_velocity(k), float(self.ctrl.ions_extra[0, k])   # box[0,k+1]
        e_i = expected_particle_injection(c.n_particles / L, self.v_th_i / math.sqrt(c.sigma),
                                          v_flux + v_extra - c.v_d_i, self.dt)
        e_i *= (1 - c.amplitude_perturbation * math.sin(max(0, k - c.hole_tracking_end_step) * self.dt * w_ion))
        e_i *= growth
        n_i = int(e_i)
        if (e_i - n_i) > self.host_rng.rand():
            n_i += 1
        e_e = expected_particle_injection(c.n_particles / L, self.v_th_e, v_flux, self.dt) * growth
        n_e = int(e_e)
        if (e_e - n_e) > self.host_rng.rand():
            n_e += 1
        return n_i, n_e

    def _step_args(self) -> StepArgs:
        """The espic_step_args block of the fused host call, built once (pointers never change)."""
        if getattr(self, "_args", None) is None:
            c = self.cfg
            a = StepArgs()
            a.grid, a.object_mask = self.grid.data_ptr(), self.object_mask.data_ptr()
            a.potential, a.charge = self.potential.data_ptr(), self.charge_density.data_ptr()
            for dst, s, den in ((a.ions, self.ions, self.ion_density), (a.electrons, self.electrons,
                                                                       self.electron_density)):
                dst.particles, dst.empty_slots = s.particles.data_ptr(), s.empty_slots.data_ptr()
                dst.current_empty_slot = s.current_empty_slot.tensor.data_ptr()
                dst.largest_index = s.largest_index.tensor.data_ptr()
                dst.density, dst.potential = den.data_ptr(), None
                dst.particle_storage_length = s.storage
                dst.largest_allowed_slot = s.empty_slots.shape[0] - 1
                dst.charge_to_mass = s.charge_to_mass
            a.n_points = self.n_points
            a.background_density = self.background_density
            a.debye_length, a.object_potential, a.object_transparency = c.debye_length, -3., 1.
            a.periodic_particles = int(c.periodic_particles)
            a.periodic_potential = int(self.periodic_potential)
            if self.fused_deposit:
                a.ions.acc, a.electrons.acc = self.acc[0].data_ptr(), self.acc[1].data_ptr()
            a.ion_tags, a.electron_tags = self.ions.tags.data_ptr(), self.electrons.tags.data_ptr()
            a.v_th_ions, a.v_th_electrons = self.v_th_i / math.sqrt(c.sigma), self.v_th_e
            self._filter_scratch = torch.empty_like(self.fine_mesh)
            a.fine_mesh, a.filter_scratch = self.fine_mesh.data_ptr(), self._filter_scratch.data_ptr()
            a.n_fine, a.tracking_dz, a.tracking_width = self.fine_mesh.shape[0], self.dz, c.debye_length * 4.
            self._args = a
        return self._args

    def step(self) -> None:
        c, k = self.cfg, self.k
        periodic = c.periodic_particles
        self.reduce_densities()
        ions_mobile = k <= c.time_steps_immobile_ions                                            # :929-939
        moving = c.moving_box_simulation
        if self.ctrl.is_static:                       # frame at rest, no background ramp (the reference's defaults)
            ramp, a_b, v_inj = 0, 0.0, float(self.ctrl.box[0, 0])
            v_inj_ions = v_inj - c.v_d_i
        else:
            ramp = self.ctrl.background_phase(k)      # 0, 1 (input-file ramp) or 2 (:852-862); advances ions_extra
            if not moving:
                self.ctrl.update_frozen(k)            # box[0,k+1] = box[0,k], box[1,k] = 0 (:926-928)
            a_b = self.ctrl.a_b(k)
            v_inj = self.ctrl.injection_velocity(k)   # (box[0,k] + box[0,k+1]) / 2 (:998, :1017)
            v_inj_ions = v_inj + float(self.ctrl.ions_extra[0, k]) - c.v_d_i
        if not ramp and not moving and not c.phase_space_histograms:
            # field solve + both pushes as one host call (espic_pic_step); same kernels, same order
            a = self._step_args()
            if self.xch is not None:   # densities entering step k: slot k&1 of every rank; pushes fill slot (k+1)&1
                nxt_i, nxt_e = self._deposit_rows(k + 1)
                a.ions.density, a.electrons.density = nxt_i.data_ptr(), nxt_e.data_ptr()
                a.exchange_parity, a.exchange_step = k & 1, k
                a.reduced_ion_density = self.ion_density.data_ptr()
                a.reduced_electron_density = self.electron_density.data_ptr()
            else:
                a.exchange_parity = -1
            a.ion_acc_pending, a.electron_acc_pending = int(self._acc_pending[0]), int(self._acc_pending[1])
            if self.keep_solved_densities and not self.fused_deposit:
                if self._solved is None:
                    self._solved = torch.zeros((2, self.n_points), dtype=torch.float32, device=self.device)
                a.density_snapshot = self._solved.data_ptr()
            a.a_b = a_b
            a.fix_ions = int(periodic or k <= c.time_steps_immobile_ions)
            a.fix_electrons = 1
            a.ions.dt = self.dt if ions_mobile else 0.
            a.electrons.dt = 0. if k < c.time_steps_immobile_electrons else self.dt
            track = c.track_hole and k > c.initial_transient_steps and k < self.hole_relative_positions.shape[0]
            a.hole_position_out = self.hole_relative_positions[k:k + 1].data_ptr() if track else None   # :896-897
            device_inject = (not periodic) and isinstance(self.sampler, DeviceUniformSampler)
            if device_inject:   # counts on the host (:965-975, 1005-1009), everything else in the same host call
                n_i, n_e = self.injection_counts(k)
                a.n_inject_ions, a.n_inject_electrons = (n_i if ions_mobile else 0), n_e   # frozen ions: :986
                a.v_d_ions, a.v_d_electrons = v_inj_ions, v_inj
                a.step_index, a.inject_dt = k, self.dt
                # sorted stores (large grids): new particles take the freed slots of their own wall
                a.inject_by_side = int(bool(self.sort_intervals["electrons"] or self.sort_intervals["ions"]))
            else:
                a.n_inject_ions = a.n_inject_electrons = 0
            self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
            self.ctx.check(self.ctx._lib.espic_pic_step(self.ctx.handle, a), "espic_pic_step")
            if self.fused_deposit:   # the solve consumed what was pending; the pushes left the next deposit in acc
                self._acc_pending = [True, True]
            done_tracking, done_inject = True, device_inject
        else:
            done_tracking = self._step_operator_by_operator(k, ions_mobile, ramp)
            done_inject = False
            if moving:   # the control law ran inside: these are the values of step k now
                a_b = self.ctrl.a_b(k)
                v_inj = self.ctrl.injection_velocity(k)
                v_inj_ions = v_inj + float(self.ctrl.ions_extra[0, k]) - c.v_d_i
        den_i, den_e = self._deposit_rows(k + 1)
        if not ions_mobile:
            den_i.fill_(1.0)   # :934 (every rank sets ones; the next all-reduce sums them, as in the reference)
            if self._acc_pending[0]:   # the frozen ions' own deposit is discarded, as the reference discards it
                self.acc[0, :self.n_points].zero_()
                self._acc_pending[0] = False
        if not periodic and not done_inject and any(self._acc_pending):
            self.flush_pending(k + 1)  # the operator-level injection below adds to the float rows
        if (not done_tracking and c.track_hole and k > c.initial_transient_steps
                and k < self.hole_relative_positions.shape[0]):                                   # :896-897
            ops.hole_position_tracking(self.potential, self.dz, c.debye_length * 4., self.grid, self.fine_mesh,
                                       out=self.hole_relative_positions[k:k + 1], ctx=self.ctx)
        if not periodic and not done_inject:
            n_i, n_e = self.injection_counts(k)
            if n_i > 0 and ions_mobile:                                                          # :986-1000
                s = self.ions
                ops.inject_particles(n_i, self.grid, self.dt, self.v_th_i / math.sqrt(c.sigma),
                                     self.background_density, self.sampler, v_inj_ions,
                                     a_b + (0.0 if self.ctrl.is_static else float(self.ctrl.ions_extra[1, k])), k,
                                     s.tags, s.particles, s.empty_slots, s.current_empty_slot, s.largest_index,
                                     den_i, ctx=self.ctx)
            if n_e > 0:                                                                          # :1015-1019
                s = self.electrons
                ops.inject_particles(n_e, self.grid, self.dt, self.v_th_e, self.background_density, self.sampler,
                                     v_inj, a_b, k, s.tags, s.particles, s.empty_slots,
                                     s.current_empty_slot, s.largest_index, den_e, ctx=self.ctx)
        if self.xch is not None:
            self._publish(k + 1)
        self.t += self.dt
        self.k += 1
        for n, every in self.sort_intervals.items():
            if every and self.k % every == 0:
                self.sort_particles((n,), key_shift=self.sort_key_shift, lookahead_steps=0.5 * every)

    def solved_densities(self):
        """(ion, electron) densities the last field solve used: reduced over the ranks, end nodes fixed."""
        if self._solved is not None:
            return self._solved[0], self._solved[1]
        return self.ion_density, self.electron_density

    def flush_pending(self, k: int | None = None) -> None:
        """Converts a pending fused deposit into the float density rows step k (default: the next step to
        run) reads (espic_convert_density: the conversion the field solve would have made, bit for bit)."""
        rows = self._deposit_rows(self.k if k is None else k)
        for sp in (0, 1):
            if self._acc_pending[sp]:
                self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
                self.ctx.check(self.ctx._lib.espic_convert_density(self.ctx.handle, self.acc[sp].data_ptr(),
                                                                   rows[sp].data_ptr(), self.n_points,
                                                                   self.background_density), "espic_convert_density")
                self._acc_pending[sp] = False

    def _publish(self, k: int) -> None:
        """Multi-GPU: makes this rank's densities for step k visible to its peers (slot k&1); a pending
        fused deposit is converted into the slot by the same launch."""
        if any(self._acc_pending):
            P = lambda sp: self.acc[sp].data_ptr() if self._acc_pending[sp] else None
            self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
            self.ctx.check(self.ctx._lib.espic_exchange_publish_fused(self.ctx.handle, k & 1, int(k), P(0), P(1),
                                                                      self.background_density),
                           "espic_exchange_publish_fused")
            self._acc_pending = [False, False]
        else:
            self.xch.publish(k, k)

    def _step_operator_by_operator(self, k: int, ions_mobile: bool, ramp: int) -> bool:
        """The same step through the individual operators, in the reference's order (used when the
        ion potential carries a background ramp, the moving-box controller is on or the per-step
        histograms are wanted).  Returns whether the hole position of this step has been tracked."""
        c = self.cfg
        self.flush_pending(k)   # the operators read and write the float density rows
        self._solved = None     # this mode's Recorder reads the arrays themselves: nothing has been pushed yet when it does
        periodic = c.periodic_particles
        fix_ions = periodic or k <= c.time_steps_immobile_ions
        if self.xch is not None:
            self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
            P = lambda t: t.data_ptr()
            self.ctx.check(self.ctx._lib.espic_field_solve_exchange(
                self.ctx.handle, k & 1, k, P(self.grid), P(self.object_mask), P(self.ion_density),
                P(self.electron_density), P(self.charge_density), int(periodic), int(fix_ions), 1, c.debye_length,
                P(self.potential), -3., 1., 0, int(self.periodic_potential), self.n_points), "field_solve_exchange")
        else:
            ops.field_solve(self.grid, self.object_mask, self.ion_density, self.electron_density,
                            self.charge_density, c.debye_length, self.potential, object_potential=-3.,
                            object_transparency=1., boltzmann_electrons=False, periodic_particles=periodic,
                            periodic_potential=self.periodic_potential, fix_ions=fix_ions, fix_electrons=True,
                            ctx=self.ctx)                                                      # :763-779, 807, 831-836
        tracked = False
        if c.moving_box_simulation:
            # the control law needs this step's hole position before the pushes (:896-897, 917-928): track
            # now and read the two floats back -- the one host synchronisation per step of this mode
            n_hist = self.hole_relative_positions.shape[0]
            if c.track_hole and c.initial_transient_steps < k < n_hist:
                ops.hole_position_tracking(self.potential, self.dz, c.debye_length * 4., self.grid, self.fine_mesh,
                                           out=self.hole_relative_positions[k:k + 1], ctx=self.ctx)
                tracked = True
            pos = float(self.hole_relative_positions[k]) if k < n_hist else 0.0
            prev = float(self.hole_relative_positions[k - 1]) if (k - 1) < n_hist else 0.0   # k = 0 reads the last (zero) entry
            self.ctrl.update(k, pos, prev)
        a_b = self.ctrl.a_b(k)
        den_i, den_e = self._deposit_rows(k + 1)
        potential_ions = self.potential                                                          # :852-864
        if ramp == 1:
            potential_ions = self.potential + self.background_potential
        elif ramp == 2:
            potential_ions = self.potential + self.background_potential_phase2
        if c.phase_space_histograms:
            self.phase_space()
        s = self.ions
        ops.move_particles(self.grid, self.object_mask, potential_ions, self.dt if ions_mobile else 0.,
                           s.charge_to_mass, self.background_density, s.largest_index, s.particles,
                           den_i, s.empty_slots, s.current_empty_slot, a_b,
                           periodic_particles=periodic, use_pure_c_version=True, ctx=self.ctx)
        s = self.electrons
        ops.move_particles(self.grid, self.object_mask, self.potential,
                           0. if k < c.time_steps_immobile_electrons else self.dt, s.charge_to_mass,
                           self.background_density, s.largest_index, s.particles, den_e,
                           s.empty_slots, s.current_empty_slot, a_b, periodic_particles=periodic,
                           use_pure_c_version=True, ctx=self.ctx)                                # :940-950
        return tracked

    def phase_space(self) -> None:
        """The three per-step 100x100 histograms (:865-895); results stay on the device."""
        c = self.cfg
        i, e = self.ions, self.electrons
        vi, ve = 20. * self.v_th_i, 4. * self.v_th_e
        self.hist["ions"] = ops.histogram2d_uniform_grid(i.particles[0], i.particles[1], c.z_min, c.z_max, -vi, vi,
                                                         100, 100, ctx=self.ctx)[0]
        self.hist["electrons"] = ops.histogram2d_uniform_grid(e.particles[0], e.particles[1], c.z_min, c.z_max,
                                                              -ve, ve, 100, 100, ctx=self.ctx)[0]
        self.hist["electrons_inject0"] = ops.histogram2d_uniform_grid(
            e.particles[0], e.particles[1], c.z_min, c.z_max, -ve, ve, 100, 100, tags=e.tags, tag_value=0,
            ctx=self.ctx)[0]
        dd.all_reduce_histograms(self.hist.values(), self.world, self.pg)                        # :876,894,895

    def sort_particles(self, species=("ions", "electrons"), key_shift: int = 0, lookahead_steps: float = 0.0) -> None:
        """Stable cell sort + compaction of the named stores (K6).  Physically a no-op: slot numbers
        change, particle state does not.  key_shift > 0 groups by blocks of 2**key_shift cells only;
        lookahead_steps sorts by the cell each particle reaches that many steps from now."""
        for name in species:
            s = getattr(self, name)
            ops.sort_particles_by_cell(self.grid, s.particles, s.tags, s.empty_slots, s.current_empty_slot,
                                       s.largest_index, periodic_particles=self.cfg.periodic_particles, ctx=self.ctx,
                                       key_shift=key_shift, lookahead=lookahead_steps * self.dt)

    def run(self, n_steps: int, check_every: int = 256) -> None:
        """n_steps timesteps.  The device status word is read every ``check_every`` steps and at the end
        (one host synchronisation each): a fatal bit -- an exchange timeout, i.e. a rank that solved
        on stale peer densities and withdrew its potential -- raises instead of producing fields."""
        i = 0
        while i < n_steps:
            room = n_steps - i
            if check_every:
                room = min(room, check_every - i % check_every)
            m = self._native_steps(room)
            if m > 0:
                self._run_native(m)
            else:
                self.step()
                m = 1
            i += m
            if check_every and i % check_every == 0:
                self.check_status()
        self.check_status()

    # Small runs (BASELINE configs[0]/[4]: 1e5..1e6 particles) take ~30 us of GPU work per step: an interpreter
    # between the launches would be most of the wall clock.  run() therefore hands whole stretches of steps
    # to espic_pic_run (one host call; steps without injection replayed from a CUDA graph) whenever the step
    # needs nothing from the host in between.
    native_run = True
    use_step_graph = True

    def _native_steps(self, limit: int) -> int:
        """How many of the next ``limit`` steps espic_pic_run can take in one call (0: step() must)."""
        c, k = self.cfg, self.k
        if (not self.native_run or not self.fused_deposit or self.xch is not None or self.world > 1
                or not self.ctrl.is_static or c.moving_box_simulation or c.phase_space_histograms):
            return 0
        if not c.periodic_particles and not isinstance(self.sampler, DeviceUniformSampler):
            return 0
        if k > c.time_steps_immobile_ions:     # frozen ions: the host resets their density row every step (:934)
            return 0
        n = min(limit, c.time_steps_immobile_ions - k + 1)
        if k < c.time_steps_immobile_electrons:
            n = min(n, c.time_steps_immobile_electrons - k)
        for every in self.sort_intervals.values():
            if every:
                n = min(n, every - k % every)
        if not c.periodic_particles:
            n = min(n, 4096)   # the device control block of an injecting stretch holds 4096 steps' counts
        return n

    def _run_native(self, n: int) -> None:
        """Steps k .. k+n-1 exactly as step() takes them, from one host call."""
        from ._lib import RunArgs
        c, k = self.cfg, self.k
        a = self._step_args()
        a.exchange_parity = -1
        a.ion_acc_pending, a.electron_acc_pending = int(self._acc_pending[0]), int(self._acc_pending[1])
        a.a_b, a.fix_ions, a.fix_electrons = 0.0, 1, 1
        a.ions.dt = self.dt
        a.electrons.dt = 0. if k < c.time_steps_immobile_electrons else self.dt
        r = RunArgs()
        r.n_steps, r.first_step, r.use_graph = n, k, int(self.use_step_graph)
        keep = None
        if not c.periodic_particles:
            counts = np.array([self.injection_counts(k + i) for i in range(n)], dtype=np.int32).reshape(n, 2)
            keep = (np.ascontiguousarray(counts[:, 0]), np.ascontiguousarray(counts[:, 1]))
            r.n_inject_ions, r.n_inject_electrons = keep[0].ctypes.data, keep[1].ctypes.data
            v_inj = float(self.ctrl.box[0, 0])
            a.v_d_ions, a.v_d_electrons = v_inj - c.v_d_i, v_inj
            a.inject_dt = self.dt
            a.inject_by_side = int(bool(self.sort_intervals["electrons"] or self.sort_intervals["ions"]))
        if c.track_hole:
            r.hole_positions = self.hole_relative_positions.data_ptr()
            r.track_from, r.track_until = c.initial_transient_steps + 1, self.hole_relative_positions.shape[0]
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.ctx.check(self.ctx._lib.espic_pic_run(self.ctx.handle, a, r), "espic_pic_run")
        del keep
        self._acc_pending = [True, True]
        for _ in range(n):
            self.t += self.dt
        self.k += n
        for name, every in self.sort_intervals.items():
            if every and self.k % every == 0:
                self.sort_particles((name,), key_shift=self.sort_key_shift, lookahead_steps=0.5 * every)

    FATAL_STATUS = 64   # ESPIC_ST_EXCHANGE_TIMEOUT (include/espic_b200.h)

    def check_status(self) -> int:
        """Reads and clears the status word; raises on a fatal bit, returns the word otherwise (the other
        bits mirror situations in which the reference only prints: bad index, stack over/underflow...)."""
        st = self.ctx.status()
        self.status_seen = getattr(self, "status_seen", 0) | st
        if st & self.FATAL_STATUS:
            from ._lib import EspicError
            raise EspicError(f"rank {self.rank}: the fused density exchange timed out at or before step {self.k} "
                             "(a peer never published its densities); the potential of that step is NaN")
        return st

    def status_word(self) -> int:
        """Every status bit raised since the run began (check_status clears the device word as it reads)."""
        self.check_status()
        return self.status_seen

    def validate(self) -> dict:
        """Self-checks of a run, to be called outside any timed region (synchronises): status word,
        finite potential, charge conservation of the pending deposit (sum(density)*background_density
        against the live particle count of this rank; the deposit accumulates integers, the only
        rounding is the one float conversion per node), and -- multi-rank -- a hash of the potential
        all-gathered over the ranks (the replicated field solve must give identical bits)."""
        import hashlib
        out = {"status": int(self.status_word()), "step": int(self.k)}
        out["potential_finite"] = bool(torch.isfinite(self.potential).all().item())
        live = self.n_active()
        den_i, den_e = self._deposit_rows(self.k)
        c = self.cfg
        ions_counted = self.k == 0 or (self.k - 1) <= c.time_steps_immobile_ions   # else the row is all ones (:934)
        rel = {}
        for sp, (name, den) in enumerate((("ions", den_i), ("electrons", den_e))):
            if name == "ions" and not ions_counted:
                continue
            if self._acc_pending[sp]:   # fused deposit still in its integer accumulators: the check is EXACT
                fixed = int(self.acc[sp, :self.n_points].sum().item())
                total = fixed / float(1 << 24)
                out.setdefault("charge_conservation_exact", {})[name] = (fixed == live[name] << 24)
            else:
                total = float(den.double().sum().item()) * self.background_density
            rel[name] = abs(total - live[name]) / max(live[name], 1)
        out["live"] = live
        out["charge_conservation_rel_err"] = rel
        out["charge_conserved"] = all(v < 2e-6 for v in rel.values())
        digest = hashlib.sha256(self.potential.cpu().numpy().tobytes()).hexdigest()[:16]
        out["potential_sha16"] = digest
        if self.world > 1:
            import torch.distributed as dist
            digests = [None] * self.world
            dist.all_gather_object(digests, digest, group=self.pg)
            out["ranks_identical"] = len(set(digests)) == 1
        out["ok"] = (out["status"] == 0 and out["potential_finite"] and out["charge_conserved"]
                     and out.get("ranks_identical", True))
        return out

    # ---- diagnostics ------------------------------------------------------------------------
    def energies(self):
        """Kinetic and field energy (the reference computes none; SURVEY.md s.8d defines them):
        KE = 1/2 sum m v^2 / background_density per species (m_i = 1, m_e = 1/mass_ratio),
        FE = 1/2 lambda_D^2 sum E_cell^2 dz."""
        c = self.cfg
        out = {}
        for name, s, mass in (("ions", self.ions, 1.0), ("electrons", self.electrons, 1.0 / c.mass_ratio)):
            alive = s.particles[0] < 0.99 * self.parked
            v = s.particles[1].double()
            out["ke_" + name] = float(0.5 * mass * (v * v * alive).sum().item() / self.background_density)
        e = -(self.potential[1:] - self.potential[:-1]).double() / self.dz
        out["fe"] = float(0.5 * c.debye_length ** 2 * (e * e).sum().item() * self.dz)
        return out

    def n_active(self):
        return {name: int((s.particles[0] < 0.99 * self.parked).sum().item())
                for name, s in (("ions", self.ions), ("electrons", self.electrons))}
