"""Host-side logic of the particle-sharded (multi-GPU) run.

The reference parallelises as one MPI rank per core with private particle shards and
replicated grids: rank r seeds its generators with 384 + r^3
(infMagSim_script.py:277), deposits with background_density = n*P*dz/(z_max-z_min) (:402, :510)
so that the P per-rank grids sum to a density normalised to 1, and exchanges exactly one thing
per step: the sum all-reduce of the ion and electron density grids (:763, :773).  Here ranks are
GPUs of one box, the two grids are the rows of ONE [2, n_points] buffer, and the exchange is a
single torch.distributed all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).  The
field solve stays replicated: every rank solves on identical reduced data.
"""
from __future__ import annotations

import torch


def rank_seed(base_seed: int, rank: int) -> int:
    """infMagSim_script.py:277."""
    return base_seed + rank * rank * rank


def background_density(n_particles_per_rank: int, world_size: int, dz: float, z_min: float, z_max: float) -> float:
    """infMagSim_script.py:402, :510 (n_engines = world size)."""
    return n_particles_per_rank * world_size * dz / (z_max - z_min)


def all_reduce_densities(densities: torch.Tensor, world_size: int, group=None) -> None:
    """Sum the [2, n_points] ion+electron buffer over ranks, in place; a no-op for one rank."""
    if world_size <= 1:
        return
    import torch.distributed as dist
    if densities.dim() != 2 or densities.shape[0] != 2 or not densities.is_contiguous():
        raise ValueError("densities must be a contiguous [2, n_points] tensor")
    dist.all_reduce(densities, op=dist.ReduceOp.SUM, group=group)


def all_reduce_histograms(hists, world_size: int, group=None) -> None:
    """int32 phase-space histograms (infMagSim_script.py:876, 894, 895)."""
    if world_size <= 1:
        return
    import torch.distributed as dist
    for h in hists:
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
