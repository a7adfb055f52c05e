"""Host-side scalar bookkeeping of the driver loop: the moving-box controller and the velocity
offset of the background ion acceleration (infMagSim_script.py:730-745, 852-862, 917-928).

The reference keeps three per-step histories -- ``box`` (box velocity and acceleration),
``ions_extra`` (velocity and acceleration the artificial background potential gives the ions) and
``hole_velocities`` -- as float32 arrays indexed by step, feeds ``box[1,k]`` to the movers as the
frame acceleration a_b and the box velocity to the wall injection, and saves all three in its
.npz.  None of this touches particle data; it stays on the host here too.

The control law needs the zero-phase filtered hole velocity at the CURRENT step,
``signal.filtfilt(B, A, hole_velocities[0:k+1], padlen=0)[k]`` (:918), which the reference
recomputes over the whole history at every step (O(n_steps^2) in total).  The last sample of a
forward-backward pass is the FIRST output of the backward pass, and that depends on the forward
pass only through its last output: with scipy's initial conditions (zi * first input sample on
either pass) it is ``(B[0] + zi[0]) * y_forward[k]``.  ``ZeroPhaseTail`` carries the forward
filter state from step to step and returns exactly that, O(1) per step
(tests/test_host_logic.py checks it against scipy.signal.filtfilt).

Arithmetic follows the reference's numpy-1 promotion: float32 history arrays, float64 scalars.
"""
from __future__ import annotations

import math

import numpy as np

HOLE_TRACKING_START_STEP = 100          # infMagSim_script.py:735
PHASE2_START, PHASE2_END = 22000, 25000  # :158-159
PHASE2_POTENTIAL = 5.0                   # :673


def butterworth2(wn: float):
    """(B, A) of the 2nd-order low-pass Butterworth filter ``signal.butter(2, wn)`` (:743):
    bilinear transform of the analogue prototype, cut-off ``wn`` as a fraction of Nyquist."""
    k = math.tan(0.5 * math.pi * wn)
    norm = 1.0 / (1.0 + math.sqrt(2.0) * k + k * k)
    b0 = k * k * norm
    return (np.array([b0, 2.0 * b0, b0]),
            np.array([1.0, 2.0 * (k * k - 1.0) * norm, (1.0 - math.sqrt(2.0) * k + k * k) * norm]))


def lfilter_zi2(b, a):
    """scipy.signal.lfilter_zi for a second-order section in transposed direct form II: the state
    that makes the response to a unit step start in its steady state."""
    y_ss = (b[0] + b[1] + b[2]) / (a[0] + a[1] + a[2])
    z1 = b[2] - a[2] * y_ss
    z0 = b[1] + z1 - a[1] * y_ss
    return np.array([z0, z1])


class ZeroPhaseTail:
    """filtfilt(B, A, x[0:k+1], padlen=0)[k] for growing k, one sample at a time."""

    def __init__(self, b, a):
        self.b, self.a = np.asarray(b, np.float64), np.asarray(a, np.float64)
        self.zi = lfilter_zi2(self.b, self.a)
        self.z = None

    def push(self, x: float) -> float:
        b, a = self.b, self.a
        x = float(x)
        if self.z is None:                      # forward pass starts from zi * x[0]
            self.z = self.zi * x
        y = b[0] * x + self.z[0]                # transposed direct form II, as scipy's lfilter
        z0 = b[1] * x + self.z[1] - a[1] * y
        z1 = b[2] * x - a[2] * y
        self.z = np.array([z0, z1])
        return float(b[0] * y + self.zi[0] * y)  # first output of the backward pass, started from zi * y


class BoxController:
    def __init__(self, n_steps: int, dt: float, dz: float, step_size: float, v_th_e: float, debye_length: float,
                 v_b_0: float = 0.0, moving_box: bool = False, hole_tracking_end_step: int = 0,
                 background_acceleration: float = 0.0, hole_pushing_start_step: int = 0,
                 hole_pushing_end_step: int = 0, n_points: int = 2):
        self.n_steps, self.dt, self.dz = int(n_steps), float(dt), float(dz)
        self.moving_box = bool(moving_box)
        self.hole_tracking_end_step = int(hole_tracking_end_step)
        self.box = np.zeros((2, self.n_steps + 1), np.float32)                       # :730
        self.box[0] = np.float32(v_b_0)                                              # :736
        self.ions_extra = np.zeros((2, self.n_steps + 1), np.float32)                # :731
        self.hole_velocities = np.zeros(self.n_steps, np.float32)                    # :733
        self.alpha = 20. * 1e-3 * v_th_e ** 2 / debye_length ** 2                    # :739
        self.beta = 6. * v_th_e / debye_length                                       # :740
        self.filter = ZeroPhaseTail(*butterworth2(.005 * step_size))                 # :737-738, 743
        # artificial background acceleration of the ions (:151-159, 672-673)
        self.set_background_acceleration = 0 < hole_pushing_start_step < self.n_steps
        self.phase1 = (int(hole_pushing_start_step), int(hole_pushing_end_step))
        bp = np.linspace(background_acceleration, 0., num=n_points, endpoint=True).astype(np.float32)
        bp2 = np.linspace(PHASE2_POTENTIAL, 0., num=n_points, endpoint=True).astype(np.float32)
        self._slopes = (float(bp[1] - bp[0]) / self.dz, float(bp2[1] - bp2[0]) / self.dz)
        # nothing ever changes: box[0] stays v_b_0, box[1] and ions_extra stay zero; the step skips the bookkeeping
        self.is_static = not self.moving_box and not self.set_background_acceleration

    def _ensure(self, k: int) -> None:
        """Runs longer than cfg.n_steps keep going: the histories grow."""
        if k + 1 >= self.box.shape[1]:
            grow = max(k + 2 - self.box.shape[1], self.box.shape[1])
            last_v = self.box[0, -1]
            self.box = np.concatenate([self.box, np.zeros((2, grow), np.float32)], axis=1)
            self.box[0, -grow:] = last_v
            self.ions_extra = np.concatenate([self.ions_extra, np.zeros((2, grow), np.float32)], axis=1)
        if k >= self.hole_velocities.shape[0]:
            self.hole_velocities = np.concatenate(
                [self.hole_velocities, np.zeros(max(k + 1 - self.hole_velocities.shape[0], self.hole_velocities.shape[0]),
                                                np.float32)])

    def background_phase(self, k: int) -> int:
        """0: no background potential on the ions at step k; 1 / 2: the input file's ramp / the
        hard-wired second phase (:852-862).  Also advances ions_extra to step k."""
        self._ensure(k)
        phase = 0
        if self.set_background_acceleration and self.phase1[0] <= k < self.phase1[1]:
            phase = 1
        elif self.set_background_acceleration and PHASE2_START <= k < PHASE2_END:
            phase = 2
        prev = float(self.ions_extra[0, k - 1])          # k = 0 reads the (zero) last element, as the reference does
        if phase:
            self.ions_extra[1, k] = self._slopes[phase - 1]
            self.ions_extra[0, k] = prev + float(self.ions_extra[1, k]) * self.dt
        else:
            self.ions_extra[0, k] = prev
        return phase

    def update(self, k: int, position: float, previous_position: float) -> None:
        """After the hole position of step k is known (:917-928): hole velocity, its zero-phase
        filtered value, and the control law for the box velocity of step k+1 / acceleration of step k."""
        self._ensure(k)
        box, dt = self.box, self.dt
        pos, prev = np.float32(position), np.float32(previous_position)
        self.hole_velocities[k] = float(pos - prev) / dt + float(box[0, k])
        filtered = self.filter.push(self.hole_velocities[k])
        if self.moving_box and HOLE_TRACKING_START_STEP < k < self.hole_tracking_end_step:
            box[0, k + 1] = (float(box[0, k]) + self.alpha * dt * float(pos)               # W_smoothing = 0 (:738)
                             + self.beta * dt * (filtered - float(box[0, k])))
            box[1, k] = (float(box[0, k + 1]) - float(box[0, k])) / dt
        else:
            box[0, k + 1] = box[0, k]
            box[1, k] = 0.

    def update_frozen(self, k: int) -> None:
        """The same with the controller off (:926-928) and without the velocity diagnostics, which
        need the hole position on the host: Recorder fills them in from the stored positions."""
        self._ensure(k)
        self.box[0, k + 1] = self.box[0, k]
        self.box[1, k] = 0.

    # the quantities the step needs
    def a_b(self, k: int) -> float:
        return float(self.box[1, k])

    def injection_flux_velocity(self, k: int) -> float:           # :966, :974
        return float(self.box[0, k + 1])

    def injection_velocity(self, k: int) -> float:                # :998, :1017
        return (float(self.box[0, k]) + float(self.box[0, k + 1])) / 2.
