"""CPU checks of the drop-in boundary: the C-ABI library is built, loads, and exports every
function include/espic_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "espic_b200.h")


@pytest.fixture(scope="module")
def built_lib():
    from espic_hole_tracking_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"enum\s*\{.*?\};", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(names))


def test_header_declares_the_reference_symbols():
    names = declared_functions()
    for ref_symbol in ("move_particles_c", "poisson_solve_c", "draw_velocities_c", "sgenrand", "genrand",
                       "electric_field_filter_c", "histogram2d_uniform_grid_c"):
        assert ref_symbol in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(handle, n)]
    assert not missing, missing


def test_python_signature_table_matches_header(built_lib):
    assert sorted(built_lib.SIGNATURES) == declared_functions()
    assert built_lib.lib().espic_abi_version() == 1


def test_no_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built_lib.EspicError):
        built_lib.Context(0)


def test_product_never_imports_the_oracle():
    """The oracle is the checker, never part of the product path."""
    pkg = os.path.join(ROOT, "espic_hole_tracking_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libespic_oracle|oracle[./]cpu|oracle/_ref|libref\.so", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{f} reaches into oracle/"


def test_operator_signatures_match_reference():
    """Argument names/order of the operator layer equal the reference's."""
    import inspect
    from espic_hole_tracking_b200 import operators as ops
    expect = {
        "move_particles": ["grid", "object_mask", "potential", "dt", "charge_to_mass", "background_density",
                           "largest_index_list", "particles", "density", "empty_slots",
                           "current_empty_slot_list", "a_b", "update_position", "periodic_particles",
                           "use_pure_c_version"],
        "accumulate_density": ["grid", "object_mask", "background_density", "largest_index", "particles",
                               "density", "empty_slots", "current_empty_slot_list", "periodic_particles",
                               "use_pure_c_version"],
        "initialize_mover": ["grid", "object_mask", "potential", "dt", "chage_to_mass", "largest_index",
                             "particles", "empty_slots", "current_empty_slot_list", "v_b", "a_b",
                             "periodic_particles", "use_pure_c_version"],
        "inject_particles": ["n_inject", "grid", "dt", "v_th", "background_density", "uniform_2d_sampler",
                             "v_d", "a_b", "k", "particles_hist", "particles", "empty_slots",
                             "current_empty_slot_list", "largest_index_list", "density"],
        "poisson_solve": ["grid", "object_mask", "charge", "debye_length", "potential", "object_potential",
                          "object_transparency", "boltzmann_electrons", "periodic_potential",
                          "use_pure_c_version"],
    }
    for name, args in expect.items():
        got = [p for p in inspect.signature(getattr(ops, name)).parameters if p != "ctx"]
        assert got == args, (name, got)
    sig = inspect.signature(ops.poisson_solve).parameters
    assert sig["object_potential"].default == -4. and sig["object_transparency"].default == 1.
    assert inspect.signature(ops.move_particles).parameters["a_b"].default is inspect.Parameter.empty


def test_ctypes_struct_mirrors_match_the_header_layout(tmp_path):
    """espic_species / espic_step_args / espic_run_args as ctypes sees them against what a C compiler makes of
    include/espic_b200.h: same size, every field at the same offset (a silent mismatch would hand the library
    garbage pointers)."""
    import shutil
    import subprocess
    from espic_hole_tracking_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    mirrors = {"espic_species": _lib.Species, "espic_step_args": _lib.StepArgs, "espic_run_args": _lib.RunArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
