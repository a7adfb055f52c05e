"""CPU oracle bindings -- TEST INFRASTRUCTURE ONLY.

Two interchangeable CPU arms with the reference's Python operator signatures
(reference: infMagSim_cython.py:32-55, 159-176, 227-285, 308-320):

* ``load("reference")``  -> oracle/_ref/libref.so, the UNMODIFIED infMagSim_c.c compiled by
  oracle/Makefile with the reference's flags.  Wall injection has no C symbol in the
  reference (its body is Cython, infMagSim_cython.py:230-285): that operator runs the
  reference's own function, cut unchanged out of the reference file and cythonized by
  oracle/build_inject_ref.py into oracle/_ref/inject_ref*.so (``reference_inject_particles``;
  ``inject_given`` uses the twin built against oracle/inject_shim.c so that the velocities can
  be supplied).  Only where those extensions are missing does it fall back to the scalar
  Python restatement below.
* ``load("port")``       -> oracle/libespic_oracle.so, our C restatement (espic_oracle.c).

Nothing under espic_hole_tracking_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
from __future__ import annotations

import ctypes as C
import glob
import importlib.machinery
import importlib.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref.so")
PORT_SO = os.path.join(HERE, "libespic_oracle.so")

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)
_f64p = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, oracle/_ref (libref.so and
    the cythonized injection routine)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


_ext_cache: dict = {}


def _load_ext(name: str):
    """An extension module built by oracle/build_inject_ref.py, or None when it is not there."""
    if name not in _ext_cache:
        hits = sorted(glob.glob(os.path.join(HERE, "_ref", name + ".*.so")))
        mod = None
        if hits:
            loader = importlib.machinery.ExtensionFileLoader(name, hits[0])
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
        _ext_cache[name] = mod
    return _ext_cache[name]


def have_inject_reference() -> bool:
    return _load_ext("inject_ref") is not None and _load_ext("inject_ref_given") is not None


class _GivenUniforms:
    """uniform_2d_sampler stand-in: get(n) returns [n, 2] whose column 0 is the supplied fractions."""

    def __init__(self, uniforms):
        self.u = np.asarray(uniforms, dtype=np.float32)

    def get(self, n):
        out = np.zeros((int(n), 2), dtype=np.float32)
        out[:, 0] = self.u[:int(n)]
        return out


def reference_inject_particles(n_inject, grid, dt, v_th, background_density, uniform_2d_sampler, v_d, a_b, k,
                               particles_hist, particles, empty_slots, current_empty_slot_list, largest_index_list,
                               density, seed=None):
    """THE reference's inject_particles (infMagSim_cython.py:230-285, compiled unchanged) with the
    reference's own draw_velocities_c; ``seed`` re-seeds its MT19937 first (seedgenrand_c)."""
    ext = _load_ext("inject_ref")
    if ext is None:
        raise FileNotFoundError("oracle/_ref/inject_ref*.so missing: run `python oracle/build_inject_ref.py`")
    if seed is not None:
        ext.seedgenrand_c(int(seed))
    ext.inject_particles(int(n_inject), grid, dt, v_th, background_density, uniform_2d_sampler, v_d, a_b, int(k),
                         particles_hist, particles, empty_slots, current_empty_slot_list, largest_index_list, density)


def reference_inject_given(n_inject, grid, dt, background_density, velocities, uniforms, a_b, k, particles_hist,
                           particles, empty_slots, current_empty_slot_list, largest_index_list, density):
    """The same compiled reference body with caller-supplied velocities (oracle/inject_shim.c)."""
    ext = _load_ext("inject_ref_given")
    if ext is None:
        raise FileNotFoundError("oracle/_ref/inject_ref_given*.so missing: run `python oracle/build_inject_ref.py`")
    if int(n_inject) <= 0:
        return
    vel = np.ascontiguousarray(velocities[:int(n_inject)], dtype=np.float32)
    shim = C.CDLL(ext.__file__)
    shim.inject_shim_set_velocities.argtypes = [_f32p, C.c_int]
    shim.inject_shim_set_velocities(vel.ctypes.data_as(_f32p), len(vel))
    ext.inject_particles(int(n_inject), grid, dt, 1.0, background_density, _GivenUniforms(uniforms), 0.0, a_b, int(k),
                         particles_hist, particles, empty_slots, current_empty_slot_list, largest_index_list, density)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_f32p)


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_i32p)


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_f64p)


_MOVE_ARGS = [_f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_float, C.c_int, _f32p, _f32p, _i32p,
              _i32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
_POISSON_ARGS = [_f32p, _f32p, _f32p, C.c_float, _f32p, C.c_float, C.c_float, C.c_int, C.c_int,
                 C.c_int]


class CpuKernels:
    """The reference's operator API on numpy arrays, backed by one of the two CPU libraries."""

    def __init__(self, kind: str):
        if kind not in ("reference", "port"):
            raise ValueError(kind)
        self.kind = kind
        path = REF_SO if kind == "reference" else PORT_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference "
                                    "for the reference arm)")
        lib = C.CDLL(path)
        self.lib = lib
        names = {
            "reference": dict(move="move_particles_c", poisson="poisson_solve_c",
                              tridiag="tridiagonal_solve_c", draw="draw_velocities_c",
                              seed="sgenrand", uniform="genrand",
                              filt="electric_field_filter_c", hist="histogram2d_uniform_grid_c"),
            "port": dict(move="oracle_move_particles", poisson="oracle_poisson_solve",
                         tridiag="oracle_tridiagonal_solve", draw="oracle_draw_velocities",
                         seed="oracle_mt_seed", uniform="oracle_mt_uniform",
                         filt="oracle_field_filter", hist="oracle_histogram2d"),
        }[kind]
        self._move = getattr(lib, names["move"])
        self._move.argtypes, self._move.restype = _MOVE_ARGS, None
        self._poisson = getattr(lib, names["poisson"])
        self._poisson.argtypes, self._poisson.restype = _POISSON_ARGS, None
        self._tridiag = getattr(lib, names["tridiag"])
        self._tridiag.argtypes, self._tridiag.restype = [_f64p] * 5 + [C.c_int], None
        self._draw = getattr(lib, names["draw"])
        self._draw.argtypes, self._draw.restype = [C.c_int, C.c_float, C.c_float, _f32p], None
        self._seed = getattr(lib, names["seed"])
        self._seed.argtypes, self._seed.restype = [C.c_ulong], None
        self._uniform = getattr(lib, names["uniform"])
        self._uniform.argtypes, self._uniform.restype = [], C.c_float
        self._filt = getattr(lib, names["filt"])
        self._filt.argtypes = [C.c_int, _f32p, _f32p, C.c_float, C.c_int, _f32p, _f32p]
        self._filt.restype = None
        self._hist = getattr(lib, names["hist"])
        self._hist.argtypes = [_f32p, _f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                               C.c_int, C.c_int, _i32p]
        self._hist.restype = None
        if kind == "port":
            self._inject = lib.oracle_inject_particles
            self._inject.argtypes = [C.c_int, _f32p, C.c_int, C.c_float, C.c_float, _f32p, _f32p,
                                     C.c_float, C.c_int, _i32p, _f32p, C.c_int, _i32p, _i32p,
                                     _i32p, _f32p]
            self._inject.restype = None

    # ---- infMagSim_cython.py:32-55 -------------------------------------------------------
    def move_particles(self, grid, object_mask, potential, dt, charge_to_mass,
                       background_density, largest_index_list, particles, density, empty_slots,
                       current_empty_slot_list, a_b, update_position=True,
                       periodic_particles=False, use_pure_c_version=True):
        top = C.c_int(int(current_empty_slot_list[0]))
        self._move(_fp(grid), _fp(object_mask), _fp(potential), dt, charge_to_mass,
                   background_density, int(largest_index_list[0]), _fp(particles), _fp(density),
                   _ip(empty_slots), C.byref(top), a_b, int(bool(update_position)),
                   int(bool(periodic_particles)), len(empty_slots) - 1, len(grid),
                   particles.shape[1])
        current_empty_slot_list[0] = top.value

    # ---- infMagSim_cython.py:159-165 -----------------------------------------------------
    def accumulate_density(self, grid, object_mask, background_density, largest_index,
                           particles, density, empty_slots, current_empty_slot_list,
                           periodic_particles=False, use_pure_c_version=True):
        potential = np.zeros_like(grid)
        self.move_particles(grid, object_mask, potential, 0, 1, background_density,
                            largest_index, particles, density, empty_slots,
                            current_empty_slot_list, a_b=0., update_position=False,
                            periodic_particles=periodic_particles)

    # ---- infMagSim_cython.py:167-176 -----------------------------------------------------
    def initialize_mover(self, grid, object_mask, potential, dt, chage_to_mass, largest_index,
                         particles, empty_slots, current_empty_slot_list, v_b, a_b,
                         periodic_particles=False, use_pure_c_version=True):
        density = np.zeros_like(grid)
        self.move_particles(grid, object_mask, potential, -dt / 2, chage_to_mass, 1,
                            largest_index, particles, density, empty_slots,
                            current_empty_slot_list, a_b=a_b, update_position=False,
                            periodic_particles=periodic_particles)
        n = largest_index[0] + 1
        # the reference's Python loop mixes a float32 element with a Python float, i.e. the
        # subtraction runs in double and narrows on store (:174-176)
        particles[1, :n] = (particles[1, :n].astype(np.float64) - float(v_b)).astype(np.float32)

    # ---- infMagSim_cython.py:308-320 -----------------------------------------------------
    def poisson_solve(self, grid, object_mask, charge, debye_length, potential,
                      object_potential=-4., object_transparency=1., boltzmann_electrons=False,
                      periodic_potential=False, use_pure_c_version=True):
        self._poisson(_fp(grid), _fp(object_mask), _fp(charge), debye_length, _fp(potential),
                      object_potential, object_transparency, int(bool(boltzmann_electrons)),
                      int(bool(periodic_potential)), len(grid))

    def tridiagonal_solve(self, a, b, c, d, x):
        self._tridiag(_dp(a), _dp(b), _dp(c), _dp(d), _dp(x), len(d))

    # ---- infMagSim_cython.py:224-228 and infMagSim_c.c:560 --------------------------------
    def seedgenrand_c(self, seed):
        self._seed(int(seed))

    def genrand(self) -> float:
        return float(self._uniform())

    def draw_velocities(self, n, v_th, v_d) -> np.ndarray:
        out = np.zeros(int(n), dtype=np.float32)
        if n > 0:
            self._draw(int(n), v_th, v_d, _fp(out))
        return out

    # ---- infMagSim_cython.py:230-285 -----------------------------------------------------
    def inject_particles(self, n_inject, grid, dt, v_th, background_density, uniform_2d_sampler,
                         v_d, a_b, k, particles_hist, particles, empty_slots,
                         current_empty_slot_list, largest_index_list, density):
        """Draw order as in the reference: host uniforms first (:245), then velocities (:248)."""
        n_inject = int(n_inject)
        if n_inject <= 0:   # the reference's &velocities[0] needs at least one element
            return
        uniforms = np.ascontiguousarray(uniform_2d_sampler.get(n_inject).astype(np.float32).T[0])
        velocities = self.draw_velocities(n_inject, v_th, v_d)
        self.inject_given(n_inject, grid, dt, background_density, velocities, uniforms, a_b, k,
                          particles_hist, particles, empty_slots, current_empty_slot_list,
                          largest_index_list, density)

    def inject_given(self, n_inject, grid, dt, background_density, velocities, uniforms, a_b, k,
                     particles_hist, particles, empty_slots, current_empty_slot_list,
                     largest_index_list, density):
        """Injection with both random inputs supplied (the form parity tests use)."""
        if self.kind == "port":
            top = C.c_int(int(current_empty_slot_list[0]))
            last = C.c_int(int(largest_index_list[0]))
            dummy = np.zeros(1, np.float32)
            self._inject(int(n_inject), _fp(grid), len(grid), dt, background_density,
                         _fp(velocities if n_inject else dummy),
                         _fp(uniforms if n_inject else dummy), a_b, int(k),
                         _ip(particles_hist.reshape(-1)), _fp(particles), particles.shape[1],
                         _ip(empty_slots), C.byref(top), C.byref(last), _fp(density))
            current_empty_slot_list[0] = top.value
            largest_index_list[0] = last.value
        elif _load_ext("inject_ref_given") is not None:
            reference_inject_given(n_inject, grid, dt, background_density, velocities, uniforms, a_b, k,
                                   particles_hist, particles, empty_slots, current_empty_slot_list,
                                   largest_index_list, density)
        else:
            inject_python(n_inject, grid, dt, background_density, velocities, uniforms, a_b, k,
                          particles_hist, particles, empty_slots, current_empty_slot_list,
                          largest_index_list, density)

    # ---- diagnostics ("next" rows) -------------------------------------------------------
    def electric_field_filter(self, grid, data_array, L, fine_mesh) -> np.ndarray:
        res = np.zeros(len(fine_mesh), dtype=np.float32)
        self._filt(len(grid), _fp(grid), _fp(data_array), L, len(fine_mesh), _fp(fine_mesh),
                   _fp(res))
        return res

    def histogram2d_uniform_grid(self, X, Y, x_min, x_max, y_min, y_max, n_bins_x, n_bins_y):
        """infMagSim_cython.py:506-528 (wrapper) over infMagSim_c.c:543-558."""
        f32 = np.float32  # the wrapper's arguments and steps are C floats (:507-511)
        x_min, x_max, y_min, y_max = f32(x_min), f32(x_max), f32(y_min), f32(y_max)
        step_x = f32((x_max - x_min) / f32(n_bins_x))
        step_y = f32((y_max - y_min) / f32(n_bins_y))
        x_centers = np.float32(np.linspace(float(x_min) + float(step_x) / 2.,
                                           float(x_max) - float(step_x) / 2., n_bins_x))
        y_centers = np.float32(np.linspace(float(y_min) + float(step_y) / 2.,
                                           float(y_max) - float(step_y) / 2., n_bins_y))
        hist = np.zeros((n_bins_x, n_bins_y), dtype=np.int32)
        X = np.ascontiguousarray(X, dtype=np.float32)
        Y = np.ascontiguousarray(Y, dtype=np.float32)
        if len(X):
            self._hist(_fp(X), _fp(Y), x_min, step_x, y_min, step_y, n_bins_x, n_bins_y, len(X),
                       _ip(hist.reshape(-1)))
        return hist, x_centers, y_centers


def inject_python(n_inject, grid, dt, background_density, velocities, uniforms, a_b, k,
                  particles_hist, particles, empty_slots, current_empty_slot_list,
                  largest_index_list, density):
    """Scalar float32 restatement of infMagSim_cython.py:230-285 (small cases only).

    Cython's typed locals are C floats, so every step below is a float32 operation except
    the math.fmod quotient, which Cython evaluates through Python floats (double).
    """
    import math
    f32 = np.float32
    largest_index = int(largest_index_list[0])
    top = int(current_empty_slot_list[0])
    n_points = len(grid)
    eps = f32(2.e-6)
    z_min, z_max = f32(grid[0]), f32(grid[n_points - 1])
    dz = f32((z_max - z_min) / f32(n_points - 1))
    dt, a_b, bg = f32(dt), f32(a_b), f32(background_density)
    hist = particles_hist.reshape(-1)
    for l in range(int(n_inject)):
        if top < 0:
            print('no empty slots')
        i = int(empty_slots[top])
        pdt = f32(dt * f32(uniforms[l]))
        hist[i] = k
        particles[1, i] = f32(f32(velocities[l]) - f32(a_b * pdt))
        wall = f32(z_max - eps) if velocities[l] < 0. else f32(z_min + eps)
        particles[0, i] = f32(wall + f32(pdt * particles[1, i]))
        largest_index = max(largest_index, i)
        x = particles[0, i]
        left = int(f32(f32(x - z_min) / dz))
        if left < 0 or left > n_points - 2:
            print('bad left_index:', left, z_min, x, z_max, particles[1, i], pdt)
        else:
            frac = math.fmod(float(x), float(dz)) / float(dz)
            share = f32(frac) if x > 0 else f32(1. + frac)
            density[left] = f32(density[left] + f32(share / bg))
            density[left + 1] = f32(density[left + 1] + f32(f32(f32(1) - share) / bg))
            empty_slots[top] = -1
            top -= 1
    largest_index_list[0] = largest_index
    current_empty_slot_list[0] = top


_cache: dict[str, CpuKernels] = {}


def load(kind: str = "reference") -> CpuKernels:
    if kind not in _cache:
        _cache[kind] = CpuKernels(kind)
    return _cache[kind]


def have(kind: str) -> bool:
    return os.path.exists(REF_SO if kind == "reference" else PORT_SO)
