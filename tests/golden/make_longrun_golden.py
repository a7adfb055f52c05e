"""Regenerates tests/golden/longrun_c1_*.npz: the ORACLE side of the long hole-tracking run (BASELINE
configs[4]) -- oracle/loop.py (the Python-3 restatement of infMagSim_script.py:756-1023) over the compiled,
unmodified reference kernels, script-default parameters (512 cells, walls + injection, electron dimple,
mass ratio 1836, step_size 0.3), from the seeded initial arrays of oracle.workloads.make_hole_plasma.
Every RECORD_EVERY steps it records what SURVEY.md s.8d lists for this configuration: max potential, the
hole position found by the reference's own matched filter (electric_field_filter_c on np.gradient(phi),
infMagSim_script.py:746-753), kinetic energies KE = 1/2 sum m v^2 / bg per species, field energy
FE = 1/2 lambda^2 sum E^2 dz, and the live populations.

    python tests/golden/make_longrun_golden.py [n_per_species] [steps]      # defaults 100000 10000; ~2 min

The GPU run starts from the same arrays (tests/test_gpu_longrun.py, scripts/longrun_configs4.py); the two
sides use different injection generators (MT19937 vs Philox), so trajectories decorrelate and the
comparison is statistical.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cpu as ocpu, loop as oloop, workloads as wl  # noqa: E402

RECORD_EVERY = 100


def diagnostics(cpu, k_, fine):
    out = {}
    out["max_phi"] = float(cpu.potential.max())
    sig = np.gradient(cpu.potential, cpu.dz).astype(np.float32)
    out["hole"] = float(fine[np.argmax(k_.electric_field_filter(cpu.grid, sig, np.float32(4 * wl.DEBYE_LENGTH), fine))])
    for name, s, mass in (("i", cpu.ions, 1.0), ("e", cpu.electrons, 1.0 / cpu.mass_ratio)):
        alive = s.particles[0] < 0.99 * 2 * wl.Z_MAX
        v = s.particles[1][alive].astype(np.float64)
        out["ke_" + name] = 0.5 * mass * float((v * v).sum()) / cpu.bg
        out["n_" + name] = int(alive.sum())
    e = -(cpu.potential[1:].astype(np.float64) - cpu.potential[:-1]) / cpu.dz
    out["fe"] = 0.5 * wl.DEBYE_LENGTH ** 2 * float((e * e).sum()) * cpu.dz
    return out


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
    steps = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000
    n_cells = 512
    kind = "reference"
    k_ = ocpu.load(kind)
    ions, electrons = wl.make_hole_plasma(n, seed=384)
    cpu = oloop.CpuSimulation(n, n_cells, False, kind=kind)
    cpu.load_particles(ions, electrons)
    cpu.initialize()
    fine = np.linspace(-3, 3, 1001).astype(np.float32)
    rec = {}
    t0 = time.perf_counter()
    for k in range(steps):
        cpu.step()
        if k % RECORD_EVERY == RECORD_EVERY - 1:
            d = diagnostics(cpu, k_, fine)
            d["step"] = k
            for key, val in d.items():
                rec.setdefault(key, []).append(val)
    out = os.path.join(HERE, f"longrun_c1_n{n:.0e}.npz".replace("+0", ""))
    np.savez_compressed(out, n=n, n_cells=n_cells, steps=steps, seed=384, record_every=RECORD_EVERY,
                        **{k: np.array(v) for k, v in rec.items()})
    print(f"{out}: {steps} oracle steps in {time.perf_counter() - t0:.0f} s")


if __name__ == "__main__":
    main()
