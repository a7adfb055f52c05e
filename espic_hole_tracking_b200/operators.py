"""The reference's Python operator API (infMagSim_cython.py) on torch CUDA tensors.

Same names, argument order, defaults and in-place/None-returning conventions as the reference:

    move_particles      infMagSim_cython.py:32-55
    accumulate_density  infMagSim_cython.py:159-165
    initialize_mover    infMagSim_cython.py:167-176
    seedgenrand_c       infMagSim_cython.py:227-228
    inject_particles    infMagSim_cython.py:230-285
    poisson_solve       infMagSim_cython.py:308-320
    histogram2d_uniform_grid  infMagSim_cython.py:506-528
    electric_field_filter     infMagSim_cython.py:567-574

Arrays are C-contiguous torch tensors on one CUDA device (float32 grids, float32 [2,S]
particle stores, int32 [S] free-slot stack, int32 [1,S] injection tags); every output is an
in-place mutation.  torch is only the tensor container: all arithmetic happens in
libespic_b200.so, reached through ctypes (_lib.py).  Python-float scalars are narrowed to C
float at this boundary exactly like the reference's typed Cython parameters.

The reference passes its two in/out scalars (stack top, largest used index) as one-element
Python lists read at entry and written at exit.  Lists are accepted and honoured (this costs a
device synchronisation per call); pass a ``DeviceInt`` instead to keep the scalar resident on
the GPU and the call asynchronous.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import Context, EspicError, c_void_p

_contexts: dict[int, Context] = {}


def default_context(device=None) -> Context:
    if device is None:
        idx = torch.cuda.current_device()
    else:
        idx = torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
    if idx not in _contexts:
        _contexts[idx] = Context(idx)
    return _contexts[idx]


def set_default_context(ctx: Context) -> None:
    _contexts[ctx.device] = ctx


class DeviceInt:
    """A device-resident int32 that quacks like the reference's one-element list."""

    def __init__(self, value: int = 0, device=None):
        self.tensor = torch.tensor([int(value)], dtype=torch.int32, device=device or "cuda")

    def __getitem__(self, i):
        if i != 0:
            raise IndexError(i)
        return int(self.tensor.item())  # synchronises

    def __setitem__(self, i, value):
        if i != 0:
            raise IndexError(i)
        self.tensor.fill_(int(value))

    def __len__(self):
        return 1

    @property
    def ptr(self):
        return c_void_p(self.tensor.data_ptr())


def _ptr(t: torch.Tensor):
    return c_void_p(t.data_ptr())


def _check(t, name, dtype, ndim, device=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch CUDA tensor, got {type(t).__name__} "
                        "(this package has no CPU path)")
    if not t.is_cuda:
        raise TypeError(f"{name}: tensor is on {t.device}; the kernels need CUDA memory")
    if t.dtype != dtype:
        raise ValueError(f"{name}: Buffer dtype mismatch, expected {dtype} but got {t.dtype}")
    if t.dim() != ndim:
        raise ValueError(f"{name}: Buffer has wrong number of dimensions (expected {ndim}, got {t.dim()})")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be C-contiguous")
    if device is not None and t.device != device:
        raise ValueError(f"{name}: on {t.device}, expected {device}")


def _enter(ctx: Context, device) -> None:
    ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)


class _Scalar:
    """Entry/exit handling of a list-or-DeviceInt scalar."""

    def __init__(self, obj, device):
        self.obj = obj
        if isinstance(obj, DeviceInt):
            self.dev = obj.tensor
            self.host = False
        else:
            self.dev = torch.tensor([int(obj[0])], dtype=torch.int32, device=device)
            self.host = True

    @property
    def ptr(self):
        return c_void_p(self.dev.data_ptr())

    def exit(self):
        if self.host:
            self.obj[0] = int(self.dev.item())


def move_particles(grid, object_mask, potential, dt, charge_to_mass, background_density,
                   largest_index_list, particles, density, empty_slots, current_empty_slot_list, a_b,
                   update_position=True, periodic_particles=False, use_pure_c_version=False, ctx=None):
    """Fused gather/push/boundary/deposit (infMagSim_cython.py:32-55 -> infMagSim_c.c:121-229).
    ``use_pure_c_version`` is accepted and ignored: there is one implementation."""
    dev = particles.device if isinstance(particles, torch.Tensor) else None
    _check(particles, "particles", torch.float32, 2, None)
    _check(grid, "grid", torch.float32, 1, dev)
    _check(object_mask, "object_mask", torch.float32, 1, dev)
    _check(potential, "potential", torch.float32, 1, dev)
    _check(density, "density", torch.float32, 1, dev)
    _check(empty_slots, "empty_slots", torch.int32, 1, dev)
    if particles.shape[0] != 2:
        raise ValueError("particles must have shape [2, S]")
    n_points = grid.shape[0]
    if not (object_mask.shape[0] >= n_points and potential.shape[0] >= n_points
            and density.shape[0] >= n_points):
        raise ValueError("object_mask, potential and density must have at least len(grid) elements")
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    storage = particles.shape[1]
    largest_allowed_slot = empty_slots.shape[0] - 1
    top = _Scalar(current_empty_slot_list, dev)
    if top.host and update_position and (empty_slots.shape[0] < 1 or int(current_empty_slot_list[0]) < 0):
        print('move_particles needs list of empty particle slots')  # infMagSim_cython.py:42-43
    lib = ctx._lib
    if isinstance(largest_index_list, DeviceInt):
        rc = lib.espic_move_particles_dev(
            ctx.handle, _ptr(grid), _ptr(object_mask), _ptr(potential), dt, charge_to_mass,
            background_density, largest_index_list.ptr, _ptr(particles), _ptr(density), _ptr(empty_slots),
            top.ptr, a_b, int(bool(update_position)), int(bool(periodic_particles)), largest_allowed_slot,
            n_points, storage)
    else:
        rc = lib.espic_move_particles(
            ctx.handle, _ptr(grid), _ptr(object_mask), _ptr(potential), dt, charge_to_mass,
            background_density, int(largest_index_list[0]), _ptr(particles), _ptr(density),
            _ptr(empty_slots), top.ptr, a_b, int(bool(update_position)), int(bool(periodic_particles)),
            largest_allowed_slot, n_points, storage)
    ctx.check(rc, "move_particles")
    top.exit()


def accumulate_density(grid, object_mask, background_density, largest_index, particles, density,
                       empty_slots, current_empty_slot_list, periodic_particles=False,
                       use_pure_c_version=False, ctx=None):
    """infMagSim_cython.py:159-165: the mover with dt=0, q/m=1, zero potential, no position update."""
    potential = torch.zeros_like(grid)
    move_particles(grid, object_mask, potential, 0, 1, background_density, largest_index, particles,
                   density, empty_slots, current_empty_slot_list, a_b=0., update_position=False,
                   periodic_particles=periodic_particles, use_pure_c_version=use_pure_c_version, ctx=ctx)


def initialize_mover(grid, object_mask, potential, dt, chage_to_mass, largest_index, particles,
                     empty_slots, current_empty_slot_list, v_b, a_b, periodic_particles=False,
                     use_pure_c_version=False, ctx=None):
    """infMagSim_cython.py:167-176: half-step-back of the velocities, then v -= v_b."""
    density = torch.zeros_like(grid)
    move_particles(grid, object_mask, potential, -dt / 2, chage_to_mass, 1, largest_index, particles,
                   density, empty_slots, current_empty_slot_list, a_b=a_b, update_position=False,
                   periodic_particles=periodic_particles, use_pure_c_version=use_pure_c_version, ctx=ctx)
    ctx = ctx or default_context(particles.device)
    if isinstance(largest_index, DeviceInt):
        n_last = largest_index[0]
    else:
        n_last = int(largest_index[0])
    ctx.check(ctx._lib.espic_shift_velocities(ctx.handle, _ptr(particles), particles.shape[1], n_last,
                                              float(v_b)), "initialize_mover")


def sort_particles_by_cell(grid, particles, particles_hist, empty_slots, current_empty_slot_list,
                           largest_index_list, periodic_particles=False, ctx=None, key_shift=0, lookahead=0.0):
    """Stable cell sort + compaction of a particle store (K6, espic_sort_particles); in place.
    No reference counterpart -- the reference never reorders its store; physically a no-op.
    key_shift > 0 sorts by (cell >> key_shift) only (espic_sort_particles_coarse); particles_hist
    may then be None.  lookahead != 0 sorts by the cell of x + v*lookahead."""
    dev = particles.device
    _check(particles, "particles", torch.float32, 2, None)
    _check(grid, "grid", torch.float32, 1, dev)
    if particles_hist is not None:
        _check(particles_hist, "particles_hist", torch.int32, 2, dev)
    _check(empty_slots, "empty_slots", torch.int32, 1, dev)
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    top = _Scalar(current_empty_slot_list, dev)
    last = _Scalar(largest_index_list, dev)
    rc = ctx._lib.espic_sort_particles_coarse(
        ctx.handle, _ptr(grid), grid.shape[0], _ptr(particles), particles.shape[1],
        _ptr(particles_hist) if particles_hist is not None else None, _ptr(empty_slots), empty_slots.shape[0],
        top.ptr, last.ptr, int(largest_index_list[0]), int(bool(periodic_particles)), int(key_shift),
        float(lookahead))
    ctx.check(rc, "sort_particles_by_cell")
    top.exit()
    last.exit()


def poisson_solve(grid, object_mask, charge, debye_length, potential, object_potential=-4.,
                  object_transparency=1., boltzmann_electrons=False, periodic_potential=False,
                  use_pure_c_version=False, ctx=None):
    """infMagSim_cython.py:308-320 -> infMagSim_c.c:356-541."""
    dev = grid.device if isinstance(grid, torch.Tensor) else None
    _check(grid, "grid", torch.float32, 1, None)
    _check(object_mask, "object_mask", torch.float32, 1, dev)
    _check(charge, "charge", torch.float32, 1, dev)
    _check(potential, "potential", torch.float32, 1, dev)
    n_points = grid.shape[0]
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    rc = ctx._lib.espic_poisson_solve(ctx.handle, _ptr(grid), _ptr(object_mask), _ptr(charge),
                                      debye_length, _ptr(potential), object_potential,
                                      object_transparency, int(bool(boltzmann_electrons)),
                                      int(bool(periodic_potential)), n_points)
    ctx.check(rc, "poisson_solve")


def prepare_charge(ion_density, electron_density, charge_density, periodic_particles=False,
                   fix_ions=True, fix_electrons=True, ctx=None):
    """The driver's end-node fix and charge density (infMagSim_script.py:763-779, 807), in place."""
    dev = ion_density.device
    _check(ion_density, "ion_density", torch.float32, 1, None)
    _check(electron_density, "electron_density", torch.float32, 1, dev)
    _check(charge_density, "charge_density", torch.float32, 1, dev)
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    rc = ctx._lib.espic_prepare_charge(ctx.handle, _ptr(ion_density), _ptr(electron_density),
                                       _ptr(charge_density), ion_density.shape[0],
                                       int(bool(periodic_particles)), int(bool(fix_ions)),
                                       int(bool(fix_electrons)))
    ctx.check(rc, "prepare_charge")


def field_solve(grid, object_mask, ion_density, electron_density, charge_density, debye_length, potential,
                object_potential=-4., object_transparency=1., boltzmann_electrons=False,
                periodic_particles=False, periodic_potential=None, fix_ions=True, fix_electrons=True, ctx=None):
    """prepare_charge + poisson_solve in one kernel launch (infMagSim_script.py:763-779, 807, 831-836)."""
    dev = grid.device if isinstance(grid, torch.Tensor) else None
    _check(grid, "grid", torch.float32, 1, None)
    for name, t in (("object_mask", object_mask), ("ion_density", ion_density),
                    ("electron_density", electron_density), ("charge_density", charge_density),
                    ("potential", potential)):
        _check(t, name, torch.float32, 1, dev)
    if periodic_potential is None:
        periodic_potential = periodic_particles
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    rc = ctx._lib.espic_field_solve(ctx.handle, _ptr(grid), _ptr(object_mask), _ptr(ion_density),
                                    _ptr(electron_density), _ptr(charge_density), int(bool(periodic_particles)),
                                    int(bool(fix_ions)), int(bool(fix_electrons)), debye_length, _ptr(potential),
                                    object_potential, object_transparency, int(bool(boltzmann_electrons)),
                                    int(bool(periodic_potential)), grid.shape[0])
    ctx.check(rc, "field_solve")


def seedgenrand_c(seed, ctx=None):
    """infMagSim_cython.py:227-228.  Seeds the device sampler (Philox key), see inject.cu."""
    ctx = ctx or default_context()
    ctx.check(ctx._lib.espic_seed(ctx.handle, int(seed)), "seedgenrand_c")


def draw_velocities(n_inject, v_th, v_d, device=None, ctx=None) -> torch.Tensor:
    """Device counterpart of draw_velocities_c (infMagSim_c.c:560-666); returns a float32 tensor."""
    ctx = ctx or default_context(device)
    dev = torch.device("cuda", ctx.device)
    out = torch.empty(int(n_inject), dtype=torch.float32, device=dev)
    if n_inject > 0:
        _enter(ctx, dev)
        ctx.check(ctx._lib.espic_draw_velocities(ctx.handle, int(n_inject), v_th, v_d, _ptr(out)),
                  "draw_velocities")
    return out


def device_uniforms(n, device=None, ctx=None) -> torch.Tensor:
    ctx = ctx or default_context(device)
    dev = torch.device("cuda", ctx.device)
    out = torch.empty(int(n), dtype=torch.float32, device=dev)
    if n > 0:
        _enter(ctx, dev)
        ctx.check(ctx._lib.espic_genrand(ctx.handle, int(n), _ptr(out)), "genrand")
    return out


def inject_particles_given(n_inject, grid, dt, background_density, velocities, uniforms, a_b, k,
                           particles_hist, particles, empty_slots, current_empty_slot_list,
                           largest_index_list, density, ctx=None):
    """The body of inject_particles (infMagSim_cython.py:250-285) with both random inputs given
    as device tensors: ``velocities`` (the draw_velocities output) and ``uniforms`` (column 0 of
    the host sampler's batch)."""
    n_inject = int(n_inject)
    if n_inject <= 0:
        return
    dev = particles.device
    _check(particles, "particles", torch.float32, 2, None)
    _check(grid, "grid", torch.float32, 1, dev)
    _check(density, "density", torch.float32, 1, dev)
    _check(empty_slots, "empty_slots", torch.int32, 1, dev)
    _check(particles_hist, "particles_hist", torch.int32, 2, dev)
    _check(velocities, "velocities", torch.float32, 1, dev)
    _check(uniforms, "uniforms", torch.float32, 1, dev)
    if velocities.shape[0] < n_inject or uniforms.shape[0] < n_inject:
        raise ValueError("velocities/uniforms shorter than n_inject")
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    top = _Scalar(current_empty_slot_list, dev)
    last = _Scalar(largest_index_list, dev)
    rc = ctx._lib.espic_inject_particles(
        ctx.handle, n_inject, _ptr(grid), grid.shape[0], dt, background_density, _ptr(velocities),
        _ptr(uniforms), a_b, int(k), _ptr(particles_hist), _ptr(particles), particles.shape[1],
        _ptr(empty_slots), top.ptr, last.ptr, _ptr(density))
    ctx.check(rc, "inject_particles")
    top.exit()
    last.exit()


def inject_particles(n_inject, grid, dt, v_th, background_density, uniform_2d_sampler, v_d, a_b, k,
                     particles_hist, particles, empty_slots, current_empty_slot_list, largest_index_list,
                     density, ctx=None):
    """infMagSim_cython.py:230-285.  ``uniform_2d_sampler.get(n)`` is called first (one [n,2]
    batch, column 0 used -- :245-246) exactly as the reference does; a sampler exposing
    ``get_device(n)`` supplies the fractions without leaving the GPU.  Velocities come from the
    device sampler (see draw_velocities)."""
    n_inject = int(n_inject)
    if n_inject <= 0:
        return
    dev = particles.device
    if hasattr(uniform_2d_sampler, "get_device"):
        uniforms = uniform_2d_sampler.get_device(n_inject)
    else:
        host = np.ascontiguousarray(uniform_2d_sampler.get(n_inject).astype(np.float32).T[0])
        uniforms = torch.from_numpy(host).to(dev, non_blocking=False)
    velocities = draw_velocities(n_inject, v_th, v_d, device=dev, ctx=ctx)
    inject_particles_given(n_inject, grid, dt, background_density, velocities, uniforms, a_b, k,
                           particles_hist, particles, empty_slots, current_empty_slot_list,
                           largest_index_list, density, ctx=ctx)


def electric_field_filter(grid, data_array, L, fine_mesh, ctx=None) -> torch.Tensor:
    """infMagSim_cython.py:567-574 -> infMagSim_c.c:668-683."""
    dev = grid.device
    _check(grid, "grid", torch.float32, 1, None)
    _check(data_array, "data_array", torch.float32, 1, dev)
    _check(fine_mesh, "fine_mesh", torch.float32, 1, dev)
    res = torch.zeros_like(fine_mesh)
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    ctx.check(ctx._lib.espic_field_filter(ctx.handle, grid.shape[0], _ptr(grid), _ptr(data_array), L,
                                          fine_mesh.shape[0], _ptr(fine_mesh), _ptr(res)),
              "electric_field_filter")
    return res


def hole_position_tracking(potential, dz, L, grid, fine_mesh, fine_dz=None, out=None, ctx=None):
    """infMagSim_script.py:746-753: fine_mesh[argmax(filter(gradient(potential, dz)))].
    Returns a one-element device tensor (no synchronisation)."""
    dev = grid.device
    _check(grid, "grid", torch.float32, 1, None)
    _check(potential, "potential", torch.float32, 1, dev)
    _check(fine_mesh, "fine_mesh", torch.float32, 1, dev)
    res = torch.empty_like(fine_mesh)
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=dev)
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    ctx.check(ctx._lib.espic_hole_position(ctx.handle, grid.shape[0], _ptr(grid), _ptr(potential),
                                           float(dz), L, fine_mesh.shape[0], _ptr(fine_mesh), _ptr(res),
                                           _ptr(out)), "hole_position_tracking")
    return out


def histogram2d_uniform_grid(X, Y, x_min, x_max, y_min, y_max, n_bins_x, n_bins_y, tags=None,
                             tag_value=0, ctx=None):
    """infMagSim_cython.py:506-528 -> infMagSim_c.c:543-558.  Returns (hist int32 [nx,ny] device
    tensor, x centres, y centres) like the reference.  ``tags``: optional int32 tensor; only
    entries equal to ``tag_value`` are counted (the driver's injection-history mask)."""
    dev = X.device
    _check(X, "X", torch.float32, 1, None)
    _check(Y, "Y", torch.float32, 1, dev)
    f32 = np.float32
    x_min, x_max, y_min, y_max = f32(x_min), f32(x_max), f32(y_min), f32(y_max)
    step_x = f32((x_max - x_min) / f32(n_bins_x))
    step_y = f32((y_max - y_min) / f32(n_bins_y))
    x_centers = np.float32(np.linspace(float(x_min) + float(step_x) / 2., float(x_max) - float(step_x) / 2.,
                                       n_bins_x))
    y_centers = np.float32(np.linspace(float(y_min) + float(step_y) / 2., float(y_max) - float(step_y) / 2.,
                                       n_bins_y))
    hist = torch.zeros((n_bins_x, n_bins_y), dtype=torch.int32, device=dev)
    n = min(X.shape[0], Y.shape[0])
    tag_ptr = c_void_p(0)
    if tags is not None:
        _check(tags, "tags", torch.int32, tags.dim(), dev)
        tag_ptr = _ptr(tags)
    ctx = ctx or default_context(dev)
    _enter(ctx, dev)
    ctx.check(ctx._lib.espic_phase_space_histogram(ctx.handle, _ptr(X), _ptr(Y), n, tag_ptr, int(tag_value),
                                                   float(x_min), float(step_x), float(y_min), float(step_y),
                                                   n_bins_x, n_bins_y, _ptr(hist)),
              "histogram2d_uniform_grid")
    return hist, x_centers, y_centers


__all__ = [n for n in dir() if not n.startswith("_")]
