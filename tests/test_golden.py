"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
compiled, unmodified reference): the C restatement on CPU, and -- marked gpu -- the CUDA path
through the operator layer.  These run without /root/reference and without oracle/_ref."""
import glob
import os

import numpy as np
import pytest

from tests.helpers import assert_bits_equal

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MOVER_FILES = sorted(glob.glob(os.path.join(GOLDEN, "mover_*.npz")))
POISSON_FILES = sorted(glob.glob(os.path.join(GOLDEN, "poisson_*.npz")))


def test_fixtures_are_present():
    assert len(MOVER_FILES) == 4 and len(POISSON_FILES) == 2
    assert os.path.exists(os.path.join(GOLDEN, "sampler_and_injection.npz"))


# ---------------------------------------------------------------- CPU: the oracle port
@pytest.mark.parametrize("path", MOVER_FILES, ids=os.path.basename)
def test_port_mover_against_golden(port, path):
    f = np.load(path)
    periodic = "periodic" in os.path.basename(path)
    p, stack, top = f["particles_in"].copy(), f["stack_in"].copy(), [int(f["top_in"])]
    den = np.zeros_like(f["grid"])
    for step in range(3):
        port.move_particles(f["grid"], np.zeros_like(f["grid"]), f["potential"], f["dt"], f["qm"], f["bg"],
                            [int(f["last"])], p, den, stack, top, a_b=np.float32(0.25 * step),
                            periodic_particles=periodic)
        assert_bits_equal(den, f["densities"][step], f"density step {step}")
    assert top == [int(f["top_out"])]
    assert_bits_equal(p, f["particles_out"], "particles")
    assert_bits_equal(stack, f["stack_out"], "stack")


@pytest.mark.parametrize("path", POISSON_FILES, ids=os.path.basename)
def test_port_poisson_against_golden(port, path):
    f = np.load(path)
    zero = np.zeros_like(f["grid"])
    cases = (("robin", zero, {}), ("periodic", zero, dict(periodic_potential=True)),
             ("object", f["mask"], dict(object_potential=-3.0, object_transparency=0.35)),
             ("boltzmann", zero, dict(boltzmann_electrons=True)))
    for name, mask, kw in cases:
        pot = f["potential_in"].copy()
        port.poisson_solve(f["grid"], mask, f["charge"], np.float32(0.125), pot, **kw)
        assert_bits_equal(pot, f[name], name)


def test_port_sampler_and_injection_against_golden(port):
    f = np.load(os.path.join(GOLDEN, "sampler_and_injection.npz"))
    port.seedgenrand_c(384)
    assert_bits_equal(port.draw_velocities(400, np.float32(42.85), np.float32(0.0)), f["vel_384_4285_0"])
    port.seedgenrand_c(385)
    vel = port.draw_velocities(400, np.float32(1.0), np.float32(0.8))
    assert_bits_equal(vel, f["vel_385_1_08"])
    p, stack, tags = f["particles_in"].copy(), f["stack_in"].copy(), np.zeros_like(f["tags_out"])
    top, last, den = [int(f["top_in"])], [int(f["last_in"])], f["density_in"].copy()
    port.inject_given(400, f["grid"], np.float32(0.01), np.float32(4.0), vel, f["uniforms"], np.float32(0.2), 17, tags, p,
                      stack, top, last, den)
    assert top == [int(f["top_out"])] and last == [int(f["last_out"])]
    assert_bits_equal(p, f["particles_out"])
    assert_bits_equal(stack, f["stack_out"])
    assert_bits_equal(tags, f["tags_out"])
    assert_bits_equal(den, f["density_out"])


# ---------------------------------------------------------------- GPU: the CUDA path
@pytest.fixture(scope="module")
def ops():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import espic_hole_tracking_b200 as pkg
    return pkg


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("path", MOVER_FILES, ids=os.path.basename)
def test_cuda_mover_against_golden(ops, path):
    import torch
    f = np.load(path)
    periodic = "periodic" in os.path.basename(path)
    p, stack, top = _dev(f["particles_in"]), _dev(f["stack_in"]), [int(f["top_in"])]
    grid, phi = _dev(f["grid"]), _dev(f["potential"])
    den = torch.zeros(len(f["grid"]), device="cuda")
    for step in range(3):
        ops.move_particles(grid, torch.zeros_like(grid), phi, float(f["dt"]), float(f["qm"]), float(f["bg"]),
                           [int(f["last"])], p, den, stack, top, 0.25 * step, periodic_particles=periodic,
                           use_pure_c_version=True)
        want = f["densities"][step]
        # ~1 particle per cell: the float32 reference sum is exact to a few ulp, so 1e-6 is a real bar
        assert np.abs(den.cpu().numpy() - want).max() <= 1e-6 * want.max(), step
    assert top == [int(f["top_out"])]
    assert_bits_equal(p.cpu().numpy(), f["particles_out"], "particles")
    t = top[0]
    assert_bits_equal(stack.cpu().numpy()[:t + 1], f["stack_out"][:t + 1], "stack")


@pytest.mark.gpu
@pytest.mark.parametrize("path", POISSON_FILES, ids=os.path.basename)
def test_cuda_poisson_against_golden(ops, path):
    f = np.load(path)
    zero = np.zeros_like(f["grid"])
    cases = (("robin", zero, {}), ("periodic", zero, dict(periodic_potential=True)),
             ("object", f["mask"], dict(object_potential=-3.0, object_transparency=0.35)),
             ("boltzmann", zero, dict(boltzmann_electrons=True)))
    for name, mask, kw in cases:
        pot = _dev(f["potential_in"])
        ops.poisson_solve(_dev(f["grid"]), _dev(mask), _dev(f["charge"]), 0.125, pot, **kw)
        err = np.abs(pot.cpu().numpy().astype(np.float64) - f[name]).max() / np.abs(f[name]).max()
        assert err <= 1e-6, (name, err)


@pytest.mark.gpu
def test_cuda_injection_against_golden(ops):
    import torch
    f = np.load(os.path.join(GOLDEN, "sampler_and_injection.npz"))
    p, stack = _dev(f["particles_in"]), _dev(f["stack_in"])
    tags = torch.zeros(f["tags_out"].shape, dtype=torch.int32, device="cuda")
    top, last, den = [int(f["top_in"])], [int(f["last_in"])], _dev(f["density_in"])
    ops.inject_particles_given(400, _dev(f["grid"]), 0.01, 4.0, _dev(f["vel_385_1_08"]), _dev(f["uniforms"]), 0.2, 17, tags,
                               p, stack, top, last, den)
    assert top == [int(f["top_out"])] and last == [int(f["last_out"])]
    assert_bits_equal(p.cpu().numpy(), f["particles_out"])
    assert_bits_equal(stack.cpu().numpy(), f["stack_out"])
    assert_bits_equal(tags.cpu().numpy(), f["tags_out"])
    assert np.abs(den.cpu().numpy() - f["density_out"]).max() <= 2e-6 * f["density_out"].max()
