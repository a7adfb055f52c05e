"""Run output in the reference's .npz layout and checkpoint / resume.

The reference writes one .npz after the last step (infMagSim_script.py:1032-1083); its analysis
notebook consumes the keys listed in NPZ_KEYS (infMagSim_analysis.ipynb:81-113).  ``Recorder``
collects the same per-step histories from a device-resident ``Simulation`` (one small D2H copy
per stored step) and writes them under the reference's file-name scheme (:1050-1060).

The reference has no checkpointing: a crash loses the run.  ``save_checkpoint`` /
``load_checkpoint`` dump and restore the complete device state of a ``Simulation`` so that the
1e5-step configuration can be resumed; a resumed periodic run continues bit-identically.
"""
from __future__ import annotations

import os

import numpy as np
import torch

NPZ_KEYS = ("grid", "times", "object_masks", "potentials", "potentials_electrons", "electric_fields",
            "ion_densities", "electron_densities", "charge_derivatives", "ion_hist_n_centers",
            "ion_hist_v_centers", "ion_distribution_functions", "electron_hist_n_centers",
            "electron_hist_v_centers", "electron_distribution_functions",
            "trapped_electron_distribution_functions", "background_ion_density", "background_electron_density",
            "box_velocities", "hole_relative_positions", "hole_velocities", "ions_vel")


class Recorder:
    """Per-step histories of the driver (infMagSim_script.py:899-914), storage_step as in :721."""

    def __init__(self, sim, storage_step: int = 1, store_all_until_step: int = 100):
        self.sim = sim
        self.storage_step, self.store_all_until_step = storage_step, store_all_until_step
        self.times, self.potentials, self.ion_densities, self.electron_densities = [], [], [], []
        self.charge_derivatives, self.histograms = [], {"ions": [], "electrons": [], "electrons_inject0": []}
        self._prev_charge = np.zeros(sim.n_points, np.float32)
        sim.keep_solved_densities = True   # see Simulation.solved_densities

    def record(self) -> None:
        """Call after ``sim.step()``: stores the fields that step solved for (time t - dt)."""
        sim = self.sim
        k = sim.k - 1
        if not (k % self.storage_step == 0 or k < self.store_all_until_step):
            return
        self.times.append(sim.t - sim.dt)
        self.potentials.append(sim.potential.cpu().numpy())
        den_i, den_e = sim.solved_densities()   # reduced, end-fixed: what the reference appends (:906-910)
        self.ion_densities.append(den_i.cpu().numpy())
        self.electron_densities.append(den_e.cpu().numpy())
        charge = sim.charge_density.cpu().numpy()
        self.charge_derivatives.append((charge - self._prev_charge) / np.float32(sim.dt))   # :808-809
        self._prev_charge = charge
        for name, lst in self.histograms.items():
            if name in sim.hist:
                lst.append(sim.hist[name].cpu().numpy())

    def filename_base(self) -> str:
        """infMagSim_script.py:1046-1060 (plain variant)."""
        sim, c = self.sim, self.sim.cfg
        pots = np.array(self.potentials)
        depth = float(np.average(np.amax(pots, axis=1)[100:])) if len(pots) > 100 else float(np.amax(pots)) if len(pots) else 0.
        return ('l' + ('%.4f' % c.debye_length) + '_Vd' + ('%.3f' % c.dimple_velocity) + '_hole_depth' + ('%.3f' % depth)
                + '_np' + ('%.1e' % sim.n_points) + '_ni' + ('%.1e' % (c.n_particles * sim.world)) + '_dt' + ('%.1e' % sim.dt)
                + '_sigma' + ('%.1e' % c.sigma) + '_mratio' + ('%.0f' % c.mass_ratio) + '_nsteps' + ('%.1e' % c.n_steps))

    def save(self, directory: str = "./") -> str:
        sim, c = self.sim, self.sim.cfg
        f32 = lambda a: np.array(a, dtype=np.float32)
        n_st = len(self.times)
        zeros_grid = np.zeros((n_st, sim.n_points), np.float32)
        centers = lambda lo, hi: np.float32(np.linspace(lo + (hi - lo) / 200., hi - (hi - lo) / 200., 100))
        vi, ve = 20. * sim.v_th_i, 4. * sim.v_th_e
        hole = sim.hole_relative_positions.cpu().numpy()
        ctrl = sim.ctrl
        pad = lambda a, n: np.concatenate([a, np.zeros(a.shape[:-1] + (max(0, n - a.shape[-1]),), a.dtype)], axis=-1)[..., :n]
        box = pad(ctrl.box[0], c.n_steps + 1)
        if c.moving_box_simulation:
            hole_vel = pad(ctrl.hole_velocities, len(hole))                                     # :917, kept by the controller
        else:   # the same formula from the stored positions (no per-step host read-back in this mode)
            hole_vel = ((hole - np.roll(hole, 1)).astype(np.float64) / sim.dt + box[:len(hole)]).astype(np.float32)
        path = os.path.join(directory, self.filename_base())
        np.savez(path, grid=sim.grid.cpu().numpy(), times=f32(self.times), object_masks=zeros_grid,
                 potentials=f32(self.potentials), potentials_electrons=f32(self.potentials), electric_fields=zeros_grid,
                 ion_densities=f32(self.ion_densities), electron_densities=f32(self.electron_densities),
                 charge_derivatives=f32(self.charge_derivatives),
                 ion_hist_n_centers=centers(c.z_min, c.z_max), ion_hist_v_centers=centers(-vi, vi),
                 ion_distribution_functions=f32(self.histograms["ions"]),
                 electron_hist_n_centers=centers(c.z_min, c.z_max), electron_hist_v_centers=centers(-ve, ve),
                 electron_distribution_functions=f32(self.histograms["electrons"]),
                 trapped_electron_distribution_functions=f32(self.histograms["electrons_inject0"]),
                 background_ion_density=sim.background_density, background_electron_density=sim.background_density,
                 box_velocities=box, hole_relative_positions=hole, hole_velocities=hole_vel,
                 ions_vel=pad(ctrl.ions_extra, c.n_steps + 1))
        return path + ".npz"


_STORE_FIELDS = ("particles", "tags", "empty_slots")


def _barrier(sim) -> None:
    """Multi-rank runs with the fused exchange: moving gigabytes of checkpoint skews the ranks by far
    more than a step; meet again before anyone publishes or waits (the exchange wait is bounded)."""
    if sim.world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize(sim.device)
        dist.barrier(group=sim.pg)


def save_checkpoint(sim, path: str) -> None:
    """Complete state of one rank: stores, stacks, tags, fields, pending deposits, the host generator,
    the device sampler's stream position (Philox key + calls consumed) and the controller."""
    sim.check_status()
    sim.flush_pending()   # a fused deposit still in its accumulators becomes the float rows stored below
    state = {"k": sim.k, "t": sim.t, "cfg": dict(sim.cfg.__dict__), "rank": sim.rank, "world": sim.world,
             "potential": sim.potential.cpu(), "densities": sim.densities.cpu(), "charge": sim.charge_density.cpu(),
             "hole": sim.hole_relative_positions.cpu(), "host_rng": sim.host_rng.get_state(),
             "rng": sim.ctx.rng_state(),
             "ctrl": {"box": sim.ctrl.box.copy(), "ions_extra": sim.ctrl.ions_extra.copy(),
                      "hole_velocities": sim.ctrl.hole_velocities.copy(),
                      "filter_z": None if sim.ctrl.filter.z is None else sim.ctrl.filter.z.copy()}}
    den_i, den_e = sim._deposit_rows(sim.k)
    state["pending_densities"] = torch.stack([den_i, den_e]).cpu()
    for name, s in (("ions", sim.ions), ("electrons", sim.electrons)):
        state[name] = {f: getattr(s, f).cpu() for f in _STORE_FIELDS}
        state[name]["top"] = s.current_empty_slot[0]
        state[name]["last"] = s.largest_index[0]
    torch.save(state, path)
    _barrier(sim)


def load_checkpoint(sim, path: str) -> None:
    """Restores a checkpoint into a Simulation built with the same configuration."""
    state = torch.load(path, weights_only=False)
    # everything that shapes the state or the trajectory must match; n_steps (history length) and the
    # sorting knobs (physically invisible) may differ
    free = {"n_steps", "sort_interval_ions", "sort_interval_electrons", "track_hole", "phase_space_histograms"}
    mine = dict(sim.cfg.__dict__)
    diff = sorted(k for k in set(mine) | set(state["cfg"]) if k not in free and mine.get(k) != state["cfg"].get(k))
    if diff:
        raise ValueError(f"checkpoint was written for a different configuration: {diff}")
    if state.get("rank", sim.rank) != sim.rank or state.get("world", sim.world) != sim.world:
        raise ValueError(f"checkpoint of rank {state.get('rank')}/{state.get('world')} loaded on rank "
                         f"{sim.rank}/{sim.world}")
    sim.k, sim.t = state["k"], state["t"]
    sim.acc.zero_()   # (flush_pending ran before the checkpoint was written: nothing was pending)
    sim._acc_pending = [False, False]
    sim.potential.copy_(state["potential"])
    sim.densities.copy_(state["densities"])
    sim.charge_density.copy_(state["charge"])
    n_hist = min(sim.hole_relative_positions.shape[0], state["hole"].shape[0])
    sim.hole_relative_positions[:n_hist].copy_(state["hole"][:n_hist])
    sim.host_rng.set_state(state["host_rng"])
    if state.get("rng") is not None:
        sim.ctx.set_rng_state(state["rng"])   # the resumed run draws what the uninterrupted run would have drawn
    den_i, den_e = sim._deposit_rows(sim.k)
    den_i.copy_(state["pending_densities"][0])
    den_e.copy_(state["pending_densities"][1])
    for name, s in (("ions", sim.ions), ("electrons", sim.electrons)):
        for f in _STORE_FIELDS:
            getattr(s, f).copy_(state[name][f])
        s.current_empty_slot[0] = state[name]["top"]
        s.largest_index[0] = state[name]["last"]
    if "ctrl" in state:
        sim.ctrl.box, sim.ctrl.ions_extra = state["ctrl"]["box"].copy(), state["ctrl"]["ions_extra"].copy()
        sim.ctrl.hole_velocities = state["ctrl"]["hole_velocities"].copy()
        sim.ctrl.filter.z = None if state["ctrl"]["filter_z"] is None else state["ctrl"]["filter_z"].copy()
    sim._plan_sorting()   # sort intervals are derived from the configuration, not stored
    _barrier(sim)
    if sim.xch is not None:
        sim.xch.publish(sim.k, sim.k)
