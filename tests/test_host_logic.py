"""Host-side logic of the driver mirror (no GPU): the reference's input.txt format, the wall-flux
injection count law, the work split of the mover and the ctypes mirror of the step-argument block."""
import ctypes
import math
import os
import re

import numpy as np

from espic_hole_tracking_b200 import _lib
from espic_hole_tracking_b200.simulation import SimConfig, expected_particle_injection

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_input_file_format(tmp_path):
    """Whitespace-separated `key value` lines, keys as in infMagSim_script.py:33-79."""
    text = """n_steps 2000
n_particles 125000
n_cells 512
mass_ratio 1836
Te/Ti 1
step_size 0.3
dimple_velocity 0
dimple_height 0.95
dimple_velocity_width 0.3
dimple_spatial_width 4
hole_tracking_end_step 0
w_ion_perturbation 0
hole_pushing_start_step 0
hole_pushing_end_step 0
background_density_growth_start_step 0
unknown_key 7
"""
    path = tmp_path / "input.txt"
    path.write_text(text)
    cfg = SimConfig.from_input_file(str(path), periodic_particles=True)
    assert cfg.n_steps == 2000 and cfg.n_particles == 125000 and cfg.n_cells == 512
    assert cfg.mass_ratio == 1836.0 and cfg.sigma == 1.0 and cfg.step_size == 0.3
    assert cfg.dimple_height == 0.95 and cfg.density_growth_start_step == 0
    assert cfg.periodic_particles is True
    assert cfg.v_d_i == 0.0 and cfg.amplitude_perturbation == 0.0        # the reference's defaults (:24-31)


def test_injection_count_law():
    """2 n v_th/sqrt(2 pi) exp(-beta^2) dt + n v_d erf(beta) dt (infMagSim_script.py:414-416): for zero
    drift it is twice the one-sided Maxwellian wall flux; a drift adds its net flux."""
    n, v_th, dt = 1e6 / 6.0, 42.85, 8.75e-4
    zero = expected_particle_injection(n, v_th, 0.0, dt)
    assert zero == 2 * n * v_th / math.sqrt(2 * math.pi) * dt
    drift = expected_particle_injection(n, v_th, 10.0, dt)
    beta = 10.0 / (math.sqrt(2) * v_th)
    assert abs(drift - (zero * math.exp(-beta ** 2) + n * 10.0 * math.erf(beta) * dt)) < 1e-9 * drift
    # Monte-Carlo: particles of a drifting Maxwellian crossing either wall of a unit-density slab in dt
    rng = np.random.default_rng(0)
    v = rng.standard_normal(4_000_000) * v_th + 10.0
    flux = np.abs(v).mean() * n * dt
    assert abs(flux / drift - 1) < 2e-3


def test_step_args_struct_matches_header():
    """The ctypes mirror of espic_step_args / espic_species has the header's fields in order."""
    src = open(os.path.join(ROOT, "include", "espic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def fields(name):
        body = re.search(r"typedef struct \{((?:(?!typedef).)*?)\} " + name + ";", src, flags=re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:  # "float v_th_ions, v_d_ions" -> the last identifier of every comma-separated piece
                out += [re.findall(r"[A-Za-z_][A-Za-z0-9_]*", piece)[-1] for piece in decl.split(",")]
        return out

    assert fields("espic_species") == [f[0] for f in _lib.Species._fields_]
    want = fields("espic_step_args")
    got = [f[0] for f in _lib.StepArgs._fields_]
    assert got == want, (got, want)
    # pointer / scalar kinds agree where it matters for the layout
    kinds = dict(_lib.StepArgs._fields_)
    assert kinds["tracking_dz"] is ctypes.c_double and kinds["hole_position_out"] is ctypes.c_void_p
    assert kinds["ions"] is _lib.Species


def test_mover_work_split_is_balanced_and_ordered():
    """Python restatement of csrc/mover.cu:Split -- contiguous ascending ranges, lengths differ by <= 1."""
    def split(n_slots, n_warps):
        n_units = (n_slots >> 2) // 64
        base, extra = divmod(n_units, n_warps)
        first = [w * base + min(w, extra) for w in range(n_warps)]
        count = [base + (1 if w < extra else 0) for w in range(n_warps)]
        return n_units, first, count
    for n_slots, n_warps in ((100_000_000, 4736), (1000, 4736), (0, 32), (256 * 4736 + 255, 4736)):
        n_units, first, count = split(n_slots, n_warps)
        assert sum(count) == n_units and max(count) - min(count) <= 1
        assert all(first[w] + count[w] == first[w + 1] for w in range(n_warps - 1))
        assert n_slots - n_units * 256 < 256 + 4          # what the last warp's generic tail handles
