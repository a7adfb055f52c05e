"""espic_hole_tracking_b200 -- the per-timestep hot path of ctzhou/ESPIC_hole_tracking's 1-D
infinite-B electrostatic PIC as hand-written sm_100a CUDA kernels behind the reference's own
operator API.  See DESIGN.md / INTEGRATION.md at the repository root."""
from ._lib import Context, EspicError, build, lib  # noqa: F401
from .operators import (  # noqa: F401
    DeviceInt, accumulate_density, draw_velocities, electric_field_filter, histogram2d_uniform_grid,
    hole_position_tracking, initialize_mover, inject_particles, inject_particles_given, move_particles,
    poisson_solve, prepare_charge, field_solve, sort_particles_by_cell, seedgenrand_c, default_context, set_default_context)

__version__ = "0.1.0"
