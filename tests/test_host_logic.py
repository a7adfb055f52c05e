"""Host-side logic of the driver mirror (no GPU): the reference's input.txt format, the wall-flux
injection count law, the work split of the mover and the ctypes mirror of the step-argument block."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

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


# ---- moving-box controller and background-acceleration bookkeeping (controller.py) ---------------

def _reference_box_loop(pos, dt, v_b_0, step_size, v_th_e, debye, end_step, start_step=100):
    """infMagSim_script.py:730-745, 917-928 transcribed (Python 3, numpy scalars as the reference has them)."""
    from scipy import signal
    n = len(pos)
    tracked, pos = pos, np.zeros(n, dtype=np.float32)      # hole_relative_positions, filled as the loop goes (:896-897)
    dt, v_th_e = np.float64(dt), np.float64(v_th_e)
    box = np.zeros([2, n + 1], dtype=np.float32)
    box[0] = v_b_0 * np.ones(n + 1).astype(np.float32)
    hole_velocities = np.zeros(n, dtype=np.float32)
    alpha = 20. * 1e-3 * v_th_e ** 2 / debye ** 2
    beta = 6. * v_th_e / debye
    B, A = signal.butter(2, .005 * step_size, output='ba')
    for k in range(n):
        pos[k] = tracked[k]
        hole_velocities[k] = (pos[k] - pos[k - 1]) / dt + box[0, k]
        f = signal.filtfilt(B, A, hole_velocities[0:k + 1], padlen=0)
        if k < end_step and k > start_step:
            box[0, k + 1] = np.average(box[0, max(0, k - 0):k + 1]) + alpha * dt * (pos[k] - 0.) \
                + beta * dt * (f[k] - box[0, k])
            box[1, k] = (box[0, k + 1] - box[0, k]) / dt
        else:
            box[0, k + 1] = box[0, k]
            box[1, k] = 0.
    return box, hole_velocities


def test_zero_phase_tail_equals_filtfilt_last_sample():
    signal = pytest.importorskip("scipy.signal")
    from espic_hole_tracking_b200.controller import ZeroPhaseTail, butterworth2, lfilter_zi2
    for wn in (.005 * 0.3, .005, .05):
        B, A = signal.butter(2, wn, output='ba')
        b, a = butterworth2(wn)
        np.testing.assert_allclose(b, B, rtol=1e-12)
        np.testing.assert_allclose(a, A, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(lfilter_zi2(B, A), signal.lfilter_zi(B, A), rtol=1e-9, atol=1e-10)
        x = (np.cumsum(np.random.default_rng(1).standard_normal(300)) * 0.1 + 3).astype(np.float32)
        tail = ZeroPhaseTail(B, A)
        for k in range(len(x)):
            want = signal.filtfilt(B, A, x[:k + 1], padlen=0)[k]
            assert abs(tail.push(x[k]) - want) <= 1e-12 * max(1.0, abs(want)), k


def test_box_controller_follows_the_reference_loop():
    pytest.importorskip("scipy.signal")
    from espic_hole_tracking_b200.controller import BoxController
    n, step_size, debye, mass_ratio = 400, 0.3, 0.125, 1836.0
    v_th_e = np.sqrt(mass_ratio)
    dt = step_size * debye / v_th_e
    rng = np.random.default_rng(7)
    # a hole drifting away from the box centre, jittered by the 1001-point tracking mesh
    pos = np.zeros(n, np.float32)
    pos[81:] = (np.round((0.002 * np.arange(n - 81) + 0.01 * rng.standard_normal(n - 81)) / 0.006) * 0.006).astype(np.float32)
    for moving, end in ((True, 350), (False, 350), (True, 0)):
        want_box, want_hv = _reference_box_loop(pos, dt, 0.5, step_size, v_th_e, debye, end if moving else -1)
        ctrl = BoxController(n, dt, 6. / 512, step_size, v_th_e, debye, v_b_0=0.5, moving_box=moving,
                             hole_tracking_end_step=end)
        for k in range(n):
            assert ctrl.background_phase(k) == 0
            ctrl.update(k, pos[k], pos[k - 1] if k else 0.0)   # at k = 0 the reference reads the still-zero last entry
        np.testing.assert_allclose(ctrl.hole_velocities, want_hv, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(ctrl.box, want_box, rtol=2e-6, atol=1e-5 * np.abs(want_box).max())
        if moving and end:
            assert np.abs(ctrl.box[1]).max() > 0 and not np.allclose(ctrl.box[0], 0.5)   # the law did act
        else:
            assert np.all(ctrl.box[0] == np.float32(0.5)) and np.all(ctrl.box[1] == 0)


def test_background_acceleration_velocity_offset():
    """ions_extra (infMagSim_script.py:852-862): constant acceleration while a ramp phase is on, the
    accumulated velocity kept afterwards; phase 2 is hard-wired to steps 22000-25000."""
    from espic_hole_tracking_b200.controller import BoxController
    n_points, dz, dt = 513, 6. / 512, 8.752e-4
    ctrl = BoxController(30000, dt, dz, 0.3, 42.85, 0.125, background_acceleration=5.0, hole_pushing_start_step=10,
                         hole_pushing_end_step=20, n_points=n_points)
    bp = np.linspace(5.0, 0., num=n_points).astype(np.float32)
    slope = (bp[1] - bp[0]) / dz
    phases = [ctrl.background_phase(k) for k in range(40)]
    assert phases[:10] == [0] * 10 and phases[10:20] == [1] * 10 and phases[20:] == [0] * 20
    assert np.all(ctrl.ions_extra[1, 10:20] == np.float32(slope)) and np.all(ctrl.ions_extra[1, 20:40] == 0)
    np.testing.assert_allclose(ctrl.ions_extra[0, 19], 10 * slope * dt, rtol=1e-5)
    assert np.all(ctrl.ions_extra[0, 20:40] == ctrl.ions_extra[0, 19])
    for k in range(40, 22001):
        phase = ctrl.background_phase(k)
    assert phase == 2
    off = BoxController(100, dt, dz, 0.3, 42.85, 0.125, background_acceleration=5.0, hole_pushing_start_step=0,
                        hole_pushing_end_step=50, n_points=n_points)
    assert all(off.background_phase(k) == 0 for k in range(100)) and not off.ions_extra.any()
