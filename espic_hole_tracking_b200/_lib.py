"""ctypes binding of libespic_b200.so (C ABI in include/espic_b200.h).

There is exactly one compute path: the sm_100a kernels in this shared library.  If the
library is missing or no Blackwell GPU is present every operator raises; nothing falls back
to CPU arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# ESPIC_LIB_PATH: development knob, an A/B build of the same library (scripts/build_variants.sh)
LIB_PATH = os.environ.get("ESPIC_LIB_PATH") or os.path.join(HERE, "libespic_b200.so")
CSRC = os.path.join(HERE, "csrc")

c_void_p, c_int, c_float, c_double = C.c_void_p, C.c_int, C.c_float, C.c_double
P = c_void_p  # every array argument is passed as a raw device (or host) address

class Species(C.Structure):
    """espic_species of include/espic_b200.h."""
    _fields_ = [("particles", c_void_p), ("empty_slots", c_void_p), ("current_empty_slot", c_void_p),
                ("largest_index", c_void_p), ("density", c_void_p), ("potential", c_void_p),
                ("particle_storage_length", c_int), ("largest_allowed_slot", c_int),
                ("charge_to_mass", c_float), ("dt", c_float), ("acc", c_void_p)]


class StepArgs(C.Structure):
    """espic_step_args of include/espic_b200.h."""
    _fields_ = [("grid", c_void_p), ("object_mask", c_void_p), ("potential", c_void_p), ("charge", c_void_p),
                ("ions", Species), ("electrons", Species), ("n_points", c_int),
                ("background_density", c_float), ("a_b", c_float), ("debye_length", c_float),
                ("object_potential", c_float), ("object_transparency", c_float), ("periodic_particles", c_int),
                ("periodic_potential", c_int),
                ("fix_ions", c_int), ("fix_electrons", c_int), ("exchange_parity", c_int),
                ("exchange_step", c_int), ("reduced_ion_density", c_void_p),
                ("reduced_electron_density", c_void_p), ("ion_acc_pending", c_int), ("electron_acc_pending", c_int),
                ("n_inject_ions", c_int), ("n_inject_electrons", c_int),
                ("v_th_ions", c_float), ("v_d_ions", c_float), ("v_th_electrons", c_float),
                ("v_d_electrons", c_float), ("inject_dt", c_float), ("step_index", c_int), ("inject_by_side", c_int), ("ion_tags", c_void_p),
                ("electron_tags", c_void_p), ("density_snapshot", c_void_p), ("fine_mesh", c_void_p), ("filter_scratch", c_void_p),
                ("hole_position_out", c_void_p), ("n_fine", c_int), ("tracking_dz", c_double),
                ("tracking_width", c_float)]


class RunArgs(C.Structure):
    """espic_run_args of include/espic_b200.h."""
    _fields_ = [("n_steps", c_int), ("first_step", c_int), ("n_inject_ions", c_void_p),
                ("n_inject_electrons", c_void_p), ("hole_positions", c_void_p), ("track_from", c_int),
                ("track_until", c_int), ("use_graph", c_int)]


# name -> (restype, argtypes); mirrors include/espic_b200.h one to one
SIGNATURES = {
    "espic_abi_version": (c_int, []),
    "espic_create": (c_int, [C.POINTER(c_void_p), c_int]),
    "espic_destroy": (None, [c_void_p]),
    "espic_set_stream": (c_int, [c_void_p, c_void_p]),
    "espic_last_error": (C.c_char_p, [c_void_p]),
    "espic_status": (c_int, [c_void_p, C.POINTER(c_int)]),
    "espic_device_info": (c_int, [c_void_p, C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int)]),
    "espic_launch_count": (C.c_longlong, [c_void_p]),
    "espic_debug_cta_cycles": (c_int, [c_void_p, C.POINTER(C.c_longlong), c_int]),
    "espic_timing_enable": (c_int, [c_void_p, c_int]),
    "espic_timing_read": (c_int, [c_void_p, C.POINTER(c_double), C.POINTER(c_int), C.POINTER(c_double)]),
    "espic_move_particles": (c_int, [c_void_p, P, P, P, c_float, c_float, c_float, c_int, P, P, P, P,
                                     c_float, c_int, c_int, c_int, c_int, c_int]),
    "espic_move_particles_dev": (c_int, [c_void_p, P, P, P, c_float, c_float, c_float, P, P, P, P, P,
                                         c_float, c_int, c_int, c_int, c_int, c_int]),
    "espic_mover_window_cells": (c_int, [c_void_p, c_int]),
    "espic_sort_particles": (c_int, [c_void_p, P, c_int, P, c_int, P, P, c_int, P, P, c_int, c_int]),
    "espic_sort_particles_coarse": (c_int, [c_void_p, P, c_int, P, c_int, P, P, c_int, P, P, c_int, c_int, c_int,
                                            c_float]),
    "espic_shift_velocities": (c_int, [c_void_p, P, c_int, c_int, c_double]),
    "espic_poisson_solve": (c_int, [c_void_p, P, P, P, c_float, P, c_float, c_float, c_int, c_int, c_int]),
    "espic_prepare_charge": (c_int, [c_void_p, P, P, P, c_int, c_int, c_int, c_int]),
    "espic_field_solve": (c_int, [c_void_p, P, P, P, P, P, c_int, c_int, c_int, c_float, P, c_float, c_float,
                                  c_int, c_int, c_int]),
    "espic_exchange_create": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "espic_exchange_open": (c_int, [c_void_p, c_void_p]),
    "espic_exchange_slot": (c_int, [c_void_p, c_int, C.POINTER(c_void_p)]),
    "espic_exchange_publish": (c_int, [c_void_p, c_int, c_int]),
    "espic_exchange_publish_fused": (c_int, [c_void_p, c_int, c_int, P, P, c_float]),
    "espic_convert_density": (c_int, [c_void_p, P, P, c_int, c_float]),
    "espic_field_solve_exchange": (c_int, [c_void_p, c_int, c_int, P, P, P, P, P, c_int, c_int, c_int, c_float, P,
                                           c_float, c_float, c_int, c_int, c_int]),
    "espic_pic_step": (c_int, [c_void_p, C.POINTER(StepArgs)]),
    "espic_pic_run": (c_int, [c_void_p, C.POINTER(StepArgs), C.POINTER(RunArgs)]),
    "espic_inject_particles": (c_int, [c_void_p, c_int, P, c_int, c_float, c_float, P, P, c_float, c_int,
                                       P, P, c_int, P, P, P, P]),
    "espic_seed": (c_int, [c_void_p, C.c_ulong]),
    "espic_draw_velocities": (c_int, [c_void_p, c_int, c_float, c_float, P]),
    "espic_genrand": (c_int, [c_void_p, c_int, P]),
    "espic_rng_get_state": (c_int, [c_void_p, C.POINTER(C.c_ulonglong)]),
    "espic_rng_set_state": (c_int, [c_void_p, C.POINTER(C.c_ulonglong)]),
    "espic_exchange_set_timeout": (c_int, [c_void_p, c_double]),
    "espic_selftest_quotient": (c_int, [c_void_p, P, c_int, C.POINTER(C.c_ulonglong)]),
    "espic_field_filter": (c_int, [c_void_p, c_int, P, P, c_float, c_int, P, P]),
    "espic_hole_position": (c_int, [c_void_p, c_int, P, P, c_double, c_float, c_int, P, P, P]),
    "espic_phase_space_histogram": (c_int, [c_void_p, P, P, c_int, P, c_int, c_float, c_float, c_float,
                                            c_float, c_int, c_int, P]),
    # reference-named host-pointer drop-ins
    "move_particles_c": (None, [P, P, P, c_float, c_float, c_float, c_int, P, P, P, P, c_float, c_int,
                                c_int, c_int, c_int, c_int]),
    "move_particles_c_minimal": (None, [P, P, P, c_float, c_float, c_float, c_int, P, P, P, P, c_int,
                                        c_int, c_int, c_int, c_int]),
    "tridiagonal_solve_c": (None, [P, P, P, P, P, c_int]),
    "espic_host_release": (None, []),
    "poisson_solve_c": (None, [P, P, P, c_float, P, c_float, c_float, c_int, c_int, c_int]),
    "draw_velocities_c": (None, [c_int, c_float, c_float, P]),
    "sgenrand": (None, [C.c_ulong]),
    "genrand": (c_float, []),
    "electric_field_filter_c": (None, [c_int, P, P, c_float, c_int, P, P]),
    "histogram2d_uniform_grid_c": (None, [P, P, c_float, c_float, c_float, c_float, c_int, c_int, c_int, P]),
    "inject_particles_c": (None, [c_int, P, c_int, c_float, c_float, P, P, c_float, c_int, P, P, c_int,
                                  P, P, P, P]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libespic_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def _stale() -> bool:
    """True when a CUDA source or the public header is newer than the built library."""
    if not os.path.exists(LIB_PATH):
        return False
    built = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "espic_b200.h"))
    return any(os.path.exists(f) and os.path.getmtime(f) > built + 1.0 for f in srcs)


def lib() -> C.CDLL:
    """The loaded shared library with argument types declared.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if _stale() and shutil.which("nvcc") or (_stale() and os.path.exists("/usr/local/cuda/bin/nvcc")):
            build()  # never run kernels older than their sources
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C espic_hole_tracking_b200/csrc`). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library out of sync
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class EspicError(RuntimeError):
    pass


class Context:
    """Owns one espic_ctx (device scratch + the stream kernels are enqueued on)."""

    def __init__(self, device: int = 0):
        self._lib = lib()
        handle = c_void_p()
        rc = self._lib.espic_create(C.byref(handle), int(device))
        if rc != 0:
            raise EspicError(
                f"espic_create(device={device}) failed with code {rc}: a Blackwell (sm_100) GPU is "
                "required; this package has no CPU path")
        self.handle = handle
        self.device = int(device)

    def check(self, rc: int, what: str = "") -> None:
        if rc != 0:
            msg = self._lib.espic_last_error(self.handle).decode(errors="replace")
            raise EspicError(f"{what} failed (code {rc}): {msg}")

    def set_stream(self, cuda_stream: int) -> None:
        self.check(self._lib.espic_set_stream(self.handle, c_void_p(cuda_stream)), "espic_set_stream")

    def status(self) -> int:
        st = c_int(0)
        self.check(self._lib.espic_status(self.handle, C.byref(st)), "espic_status")
        return st.value

    def rng_state(self):
        """(key, calls consumed, seeded) of the device sampler (checkpoints)."""
        st = (C.c_ulonglong * 3)()
        self.check(self._lib.espic_rng_get_state(self.handle, st), "espic_rng_get_state")
        return [int(v) for v in st]

    def set_rng_state(self, state) -> None:
        st = (C.c_ulonglong * 3)(*[int(v) for v in state])
        self.check(self._lib.espic_rng_set_state(self.handle, st), "espic_rng_set_state")

    def set_exchange_timeout(self, seconds: float) -> None:
        self.check(self._lib.espic_exchange_set_timeout(self.handle, float(seconds)), "espic_exchange_set_timeout")

    def launch_count(self) -> int:
        return int(self._lib.espic_launch_count(self.handle))

    def device_info(self):
        a, b, c = c_int(), c_int(), c_int()
        self.check(self._lib.espic_device_info(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"sm_count": a.value, "mover_grid": b.value, "mover_block": c.value}

    def timing_enable(self, on: bool = True) -> None:
        self.check(self._lib.espic_timing_enable(self.handle, int(on)), "espic_timing_enable")

    def timing_read(self):
        mean, n, tot = c_double(), c_int(), c_double()
        self.check(self._lib.espic_timing_read(self.handle, C.byref(mean), C.byref(n), C.byref(tot)))
        return {"mean_ms": mean.value, "launches": n.value, "total_ms": tot.value}

    def close(self) -> None:
        if getattr(self, "handle", None):
            self._lib.espic_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
