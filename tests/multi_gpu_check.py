"""Run under torchrun on >= 2 GPUs.

Part 1: the P-way sharded run on P GPUs (fused peer-memory density exchange) against the ORACLE loop
(oracle/loop.py over the compiled reference, infMagSim_script.py:756-1023) on the UNION of the shards on
one CPU -- the reference's own statement of what a sharded run computes ("P-way sharded
result == 1-GPU result up to summation order"), at 4096 and 65536 cells, periodic and wall-bounded.
Part 2: the fused exchange against the NCCL all-reduce path, step by step, from identical particle arrays.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espic_hole_tracking_b200.simulation import SimConfig, Simulation  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ok = True


def shard_particles(r, n, seed=11):
    rng = np.random.default_rng(seed + 1000 * r)
    out = []
    for v_th in (1.0, float(np.sqrt(1836.0))):
        x = rng.uniform(-3.0, 3.0, n)
        x = x + 0.05 * np.sin(2 * np.pi * x / 6.0)
        out.append(np.stack([x.astype(np.float32), (rng.standard_normal(n) * v_th).astype(np.float32)]))
    return out


def check_vs_oracle(n_cells, periodic, n=200_000, steps=10):
    """Every rank pushes its own shard on its GPU; rank 0 also runs the oracle loop on all shards
    concatenated (background density n*P*dz/L on both sides, ref :402, :510).  Wall injection draws
    from generators that differ by construction (MT19937 vs Philox), so the wall-bounded case runs with
    the injection counts forced to zero on both sides: removal at the walls is still exercised."""
    from oracle import cpu as ocpu, loop as oloop
    cfg = SimConfig(n_steps=steps + 2, n_particles=n, n_cells=n_cells, periodic_particles=periodic,
                    periodic_potential=False, track_hole=False, extra_storage_factor=1.25,
                    sort_interval_electrons=4 if n_cells > 20000 else None)
    sim = Simulation(cfg, rank=rank, world_size=world, device=dev, exchange="p2p")
    sim.injection_counts = lambda k: (0, 0)
    sim.load_particles(*shard_particles(rank, n))
    sim.initialize()
    one = None
    if rank == 0:
        kind = "reference" if ocpu.have("reference") else "port"
        one = oloop.CpuSimulation(world * n, n_cells, periodic, kind=kind, n_ranks=1, rank=0, periodic_potential=False)
        shards = [shard_particles(r, n) for r in range(world)]
        one.load_particles(np.concatenate([s[0] for s in shards], axis=1), np.concatenate([s[1] for s in shards], axis=1))
        one.initialize()
    worst_phi = worst_den = den0 = worst_sum = 0.0
    good = True
    identical = True
    for k in range(steps):
        sim.step()   # solves step k's field from the all-rank sum of step k's densities, then pushes
        gathered = [torch.empty_like(sim.potential) for _ in range(world)]
        dist.all_gather(gathered, sim.potential)
        identical &= all(torch.equal(gathered[0], g) for g in gathered)   # replicated solve: identical bits on every rank
        good &= identical
        if rank == 0:
            one.field_step()
            phi = sim.potential.cpu().numpy().astype(np.float64)
            scale = np.abs(one.potential).max()
            worst_phi = max(worst_phi, float(np.abs(phi - one.potential).max() / scale))
            # reduced, end-fixed densities the GPU solve used (written by the fused kernel) against the oracle's.
            # Step 0 (identical particle arrays on both sides): equal up to float32 summation order -- the oracle's own
            # serial accumulation is ~1e-6 off the exact sum at these counts.  From step 1 on the two runs are
            # free-running: the potentials differ by summation order, the electrons (q/m = -1836) turn that into
            # position differences that grow step by step, and the reference's mirrored deposit
            # (infMagSim_c.c:219-225) is discontinuous at the nodes, so node values differ by whole particle weights;
            # what stays comparable is the potential, the totals and the populations.
            one_particle = 1.0 / one.bg
            for g, c in ((sim.ion_density, one.ion_density), (sim.electron_density, one.electron_density)):
                gd = g.cpu().numpy().astype(np.float64)
                diff = np.abs(gd - c)
                if k == 0:
                    den0 = max(den0, float(diff.max() / c.max()))
                worst_den = max(worst_den, float(diff.max() / one_particle))
                worst_sum = max(worst_sum, abs(float(gd[1:-1].sum() - c[1:-1].astype(np.float64).sum())) / float(c.sum()))
            one.push_step()
            one.t += one.dt
            one.k += 1
    live = sim.n_active()
    counts = torch.tensor([live["ions"], live["electrons"]], device=dev)
    dist.all_reduce(counts)
    st = sim.status_word()
    if rank == 0:
        act = one.n_active()
        # the densities differ from the oracle's by float32 summation order only (its serial accumulation is
        # itself ~1e-6 off the exact sum at these counts); the solve amplifies that by ~(L/lambda_D)^2/8 ~ 300
        # in the net-charge mode of the Robin system (same bound as tests/test_distributed_cpu.py)
        good &= den0 <= 5e-6 and worst_den <= 12.0 and worst_sum <= 2e-5 and worst_phi <= 3e-3
        # free-running: a particle within an ulp of a wall may leave a step earlier on one side
        good &= abs(int(counts[0]) - act["ions"]) <= 3 and abs(int(counts[1]) - act["electrons"]) <= 3
        # rank 0's own shard against the oracle's first n slots: same trajectories up to the field difference
        x_gpu = sim.ions.particles[0, :n].cpu().numpy()
        x_cpu = one.ions.particles[0, :n]
        both = (x_gpu < 5.9) & (x_cpu < 5.9)
        sorted_store = n_cells > 20000   # the windowed mover re-sorts: slots no longer correspond
        dx = 0.0 if sorted_store else float(np.abs(x_gpu[both] - x_cpu[both]).max())
        good &= dx <= 2e-5
        print(f"ORACLE cells={n_cells} periodic={periodic} ranks={world}: max|phi-phi_oracle|/max|phi| = {worst_phi:.3e} "
              f"(tol 3e-3), density at step 0: max error / peak = {den0:.2e} (tol 5e-6), later steps (free-running): interior total off "
              f"by {worst_sum:.1e} (tol 2e-5), largest node difference {worst_den:.2f} particle weights; live (all ranks) {counts.tolist()} "
              f"vs oracle {[act['ions'], act['electrons']]}, max ion |dx| = {dx:.2e}, status {st}, "
              f"ranks bit-identical: {bool(identical)}", flush=True)
    good &= st == 0
    return bool(good)


for n_cells in [int(c) for c in os.environ.get("ESPIC_ORACLE_CELLS", "4096,65536").split(",") if c]:
    for periodic in (True, False):
        ok &= check_vs_oracle(n_cells, periodic)

# 1024 cells: field solve as a 2-CTA cluster; 65536 cells: 8 CTAs with 9 rows per thread, the windowed mover and
# its automatic re-sorting (local to each rank) under the exchange
cells_list = [int(c) for c in os.environ.get("ESPIC_CHECK_CELLS", "1024,65536").split(",")]
for n_cells, periodic in [(c, p) for c in cells_list for p in (True, False)]:
    cfg = SimConfig(n_steps=40, n_particles=400_000, n_cells=n_cells, periodic_particles=periodic,
                    periodic_potential=False, track_hole=False, sort_interval_electrons=5 if n_cells > 20000 else None)
    sims = {}
    for mode in ("nccl", "p2p"):
        sim = Simulation(cfg, rank=rank, world_size=world, device=dev, exchange=mode)
        sim.init_synthetic()   # same seed per rank in both modes -> identical shards
        sim.initialize()
        sims[mode] = sim
    worst = 0.0
    for k in range(12):
        for sim in sims.values():
            sim.step()
        a, b = sims["nccl"].potential, sims["p2p"].potential
        scale = float(a.abs().max())
        worst = max(worst, float((a - b).abs().max()) / scale)
        # every rank must hold bit-identical fields in the fused path
        gathered = [torch.empty_like(b) for _ in range(world)]
        dist.all_gather(gathered, b)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        ok &= same
    act = [sims[m].n_active() for m in sims]
    st = [sims[m].status_word() for m in sims]
    # with 2 ranks a+b is the same sum in any order: bitwise; more ranks: summation order only
    tol = 0.0 if world == 2 and periodic else (5e-2 if not periodic else 3e-3)
    good = worst <= tol and st == [0, 0] and (periodic is False or act[0] == act[1])
    ok &= good
    if rank == 0:
        print(f"cells={n_cells} periodic={periodic}: max |phi_p2p - phi_nccl| / max|phi| over 12 steps = {worst:.3e} (tol {tol}); "
              f"ranks bit-identical: {same}; status {st}; active {act}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
