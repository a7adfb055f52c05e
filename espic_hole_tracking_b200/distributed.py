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


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap memory the library allocated."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class DensityExchange:
    """The fused replacement of the density all-reduce (espic_exchange_* in include/espic_b200.h).

    Every rank owns two parity slots of [2, n_points] floats in its HBM, IPC-mapped by its peers.
    ``slot(parity)`` is a torch view of the local slot: the movers and the injector of a step
    deposit into it, ``publish`` makes it visible, and the next step's field-solve kernel sums all
    ranks' slots itself over NVLink (rank order => bit-identical fields on every rank).  The IPC
    handles travel once, at set-up, through torch.distributed.
    """

    def __init__(self, ctx, n_points: int, world_size: int, rank: int, device, group=None):
        import ctypes as C

        import torch.distributed as dist
        self.ctx, self.world, self.rank = ctx, world_size, rank
        handle = (C.c_ubyte * 64)()
        ctx.check(ctx._lib.espic_exchange_create(ctx.handle, int(n_points), int(world_size), int(rank), handle),
                  "espic_exchange_create")
        handles = [None] * world_size
        if world_size > 1:
            dist.all_gather_object(handles, bytes(handle), group=group)
        else:
            handles[0] = bytes(handle)
        blob = (C.c_ubyte * (64 * world_size)).from_buffer_copy(b"".join(handles))
        ctx.check(ctx._lib.espic_exchange_open(ctx.handle, blob), "espic_exchange_open")
        n_pad = (n_points + 31) & ~31
        self._slots = []
        for parity in (0, 1):
            ptr = C.c_void_p()
            ctx.check(ctx._lib.espic_exchange_slot(ctx.handle, parity, C.byref(ptr)), "espic_exchange_slot")
            t = torch.as_tensor(_CudaArray(ptr.value, (2, n_pad)), device=device)
            if t.data_ptr() != ptr.value:
                raise RuntimeError("torch copied the exchange slot instead of wrapping it")
            self._slots.append(t[:, :n_points])
        if world_size > 1:
            dist.barrier(group=group)  # every peer has mapped every buffer before anyone publishes

    def slot(self, parity: int) -> torch.Tensor:
        """[2, n_points] view (rows are n_pad apart): row 0 ion, row 1 electron density of this rank."""
        return self._slots[parity & 1]

    def publish(self, parity: int, step: int) -> None:
        self.ctx.check(self.ctx._lib.espic_exchange_publish(self.ctx.handle, parity & 1, int(step)),
                       "espic_exchange_publish")
