#!/usr/bin/env python
"""bench.py -- particle-steps/s of the PIC hot path on N B200s (one rank per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 1e8 ions + 1e8 electrons per GPU, 4096-cell periodic grid,
synthetic seeded particles; N>1 keeps the per-GPU work fixed (weak scaling) and adds the per-step
density all-reduce.  A "step" is one full timestep over both species: end-node fix + charge
density, field solve, ion push, electron push (each push = fused gather/push/boundary/deposit).

Keys of the JSON line beyond the base contract:
  roofline      the mover kernel: 16 algorithmic bytes per particle-step over its
                mean launch duration (CUDA events recorded on the launching stream inside the
                timed region) against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own C path (oracle/_ref, the compiled infMagSim_c.c) driven by the
                Python-3 restatement of its step loop on a bounded sample, one process per core
  e2e           the same steps through the public operator API with the per-step host traffic of
                the reference driver inside the timed region (control block H2D from pinned
                memory; potential + both densities + stack tops D2H; host sync every step).
                Particle state stays resident in HBM, as it stays resident in host RAM in the
                reference.
  e2e_host_arrays  the literal drop-in: move_particles_c/poisson_solve_c called with HOST arrays
                (particles cross PCIe in both directions every call).
  checks        self-validation after every timed loop, outside it (Simulation.validate): device status
                word, finite potential, charge conservation of the deposit against the live particle
                count, and -- N>1 -- an all-gathered hash of the potential equal on all ranks.  A failed
                check makes the process exit non-zero: no number without a valid run.
  configs3      (N=1) BASELINE configs[3] in front of the driver: 65536 cells, absorbing walls + injection,
                electron dimple, hole tracking every step, 1e8+1e8 particles, four electron sort
                intervals inside the timed region, through Simulation.run() (steps replayed from CUDA graphs);
                its roofline for k_move_win<walls> comes from a second pass of as many steps with an event
                around every mover launch (ms_per_step_instrumented).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec"
N_PER_SPECIES = 100_000_000
N_CELLS = 4096
ALGO_BYTES_PER_PARTICLE_STEP = 16


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            in_region = (t0 - 0.05) <= ts <= (t1 + 0.15)
            try:
                if in_region:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
            except ValueError:
                continue
            if in_region:
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args):
    """--impl reference: the reference's own C implementation of the path (oracle/_ref when the
    reference compiled in the build container, else our C port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu, loop
    kind = "reference" if cpu.have("reference") else "port"
    if kind == "port" and not cpu.have("port"):
        cpu.build()
    cores = os.cpu_count() or 1
    n_sample = 2_000_000  # per species per process: each step is a bounded sample of the workload
    res = loop.timed_cpu_sample(kind, n_sample, N_CELLS, True, steps=args.steps, warmup=args.warmup,
                                periodic_potential=False)
    value = res["value"]
    sample = (f"{cores} processes x ({n_sample:.0e} ions + {n_sample:.0e} electrons), {N_CELLS}-cell periodic grid, "
              f"{args.steps} timed steps after {args.warmup} warm-up; same loop as the GPU arm")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"1e8 ions + 1e8 electrons, {N_CELLS}-cell periodic grid (configs[1]); CPU arm timed on a "
                               "bounded sample", "cells": N_CELLS, "periodic": True},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=N_PER_SPECIES, help="per species per GPU")
    ap.add_argument("--cells", type=int, default=N_CELLS)
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1: density exchange fused into the field-solve kernel over NVLink peer memory (p2p, "
                         "default) or one NCCL all-reduce per step")
    ap.add_argument("--walls", action="store_true",
                    help="probe, not the headline: absorbing walls + injection (the reference's default mode) "
                         "instead of periodic particles")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-arrays", action="store_true")
    ap.add_argument("--no-configs3", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from espic_hole_tracking_b200.simulation import SimConfig, Simulation

    n = int(args.particles)
    cfg = SimConfig(n_steps=2 * (args.steps + args.warmup) + 8, n_particles=n, n_cells=args.cells,
                    periodic_particles=not args.walls, periodic_potential=False,
                    extra_storage_factor=1.05 if args.walls else 1.0, track_hole=False)
    sim = Simulation(cfg, rank=rank, world_size=world, device=device, exchange=args.exchange)
    sim.init_synthetic()
    sim.initialize()
    ctx = sim.ctx
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value, roofline) ----------------
    for _ in range(args.warmup):
        sim.step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    launches0 = ctx.launch_count()
    ctx.timing_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        sim.step()
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    timing = ctx.timing_read()
    ctx.timing_enable(False)
    launches = ctx.launch_count() - launches0
    clk = clocks.stop(t0, t1) if rank == 0 else None
    checks = {"after_device_timed_loop": sim.validate()}
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    particle_steps = 2.0 * n * args.steps * world
    value = particle_steps / (ms * 1e-3)

    peak, peak_src = measured_peak()
    mover_ms = timing["mean_ms"]
    achieved = ALGO_BYTES_PER_PARTICLE_STEP * n / (mover_ms * 1e-3) / 1e9 if mover_ms > 0 else 0.0
    window = ctx._lib.espic_mover_window_cells(ctx.handle, sim.n_points)
    flavour = "walls" if args.walls else "periodic"
    kernel = (f"espic::k_move_win<{flavour}> ({window}-cell shared-memory window per store segment)" if window
              else f"espic::k_move<{flavour}, smem tables>")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": kernel,
                "launch_ms_mean": mover_ms, "launches_timed": timing["launches"],
                "mover_share_of_step": timing["total_ms"] / ms if ms > 0 else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PARTICLE_STEP * n}
    traffic_path = os.path.join(ROOT, "profiles", "mover_traffic.json")
    if os.path.exists(traffic_path):
        try:
            with open(traffic_path) as fh:
                tr = json.load(fh)
            if int(tr.get("particles_per_launch", 0)) == n and not window and not args.walls:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---------------- end-to-end through the operator API with host traffic ----------------
    ctrl_host = torch.zeros(4, dtype=torch.float32).pin_memory()
    ctrl_dev = torch.zeros(4, dtype=torch.float32, device=device)
    out_host = [torch.zeros((3, sim.n_points), dtype=torch.float32).pin_memory() for _ in range(2)]
    tops_host = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    h2d = ctrl_host.numel() * 4
    d2h = out_host[0].numel() * 4 + tops_host[0].numel() * 4
    max_phi = []

    out_np = [o.numpy() for o in out_host]
    tops_np = [t.numpy() for t in tops_host]

    def consume(b):  # the per-step host work of the driver: max(potential) (:1024-1027), stack tops
        done[b].synchronize()
        max_phi.append(float(out_np[b][0].max()))
        assert int(tops_np[b][0]) >= -1

    def e2e_iteration(i):
        b = i & 1
        ctrl_host[0] = sim.a_b
        ctrl_host[1] = float(sim.k)
        ctrl_dev.copy_(ctrl_host, non_blocking=True)            # per-step control block (box acceleration, step)
        sim.step()
        out_host[b][0].copy_(sim.potential, non_blocking=True)  # what the reference driver stores each step
        out_host[b][1:3].copy_(sim.densities, non_blocking=True)
        tops_host[b][0:1].copy_(sim.ions.current_empty_slot.tensor, non_blocking=True)
        tops_host[b][1:2].copy_(sim.electrons.current_empty_slot.tensor, non_blocking=True)
        done[b].record()
        if i > 0:
            consume(1 - b)      # read step i-1's results while step i runs (one-step lag, every step is read)

    for i in range(args.warmup):  # warm the host side of the loop too (pinned copies, event waits)
        e2e_iteration(i)
    consume((args.warmup - 1) & 1)
    max_phi.clear()
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        e2e_iteration(i)
    consume((args.steps - 1) & 1)
    barrier()
    e_ms = (time.perf_counter() - e0) * 1e3
    assert len(max_phi) == args.steps
    checks["after_e2e_loop"] = sim.validate()
    checks["ok"] = all(c["ok"] for c in checks.values())
    if world > 1:
        t = torch.tensor([e_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    e2e = {"value": particle_steps / (e_ms * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h,
           "note": "Simulation.step() (public driver API, one espic_pic_step host call per step), particle state "
                   "resident in HBM; per step: control block H2D from pinned memory, potential + both densities + "
                   "stack tops D2H into pinned memory and read on the host (max potential) with a one-step lag; "
                   "wall clock"}

    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n:.3g} ions + {n:.3g} electrons per GPU, {args.cells}-cell "
                               f"{'wall-bounded grid with injection' if args.walls else 'periodic grid'} "
                               "(BASELINE configs[1]; N>1 = configs[2]-style weak scaling with the density all-reduce)",
                   "particles_per_species_per_gpu": n, "cells": args.cells, "periodic": not args.walls,
                   "field_solve": "periodic particles, Robin-end potential (periodic_potential=False): the reference's "
                                  "periodic_potential branch is numerically broken on 4097 points (DESIGN.md s.8)",
                   "l2_policy": "inputs (1.6 GB of particle state per GPU) are larger than the 126 MB L2; no flush needed",
                   "parallelism": (f"particle shards x{world}; density exchange: " +
                                   {"p2p": "fused into the field-solve kernel (every rank sums all ranks' "
                                           f"[2,{args.cells + 1}] slots over NVLink peer memory, no collective launch)",
                                    "nccl": f"one NCCL fp32 all-reduce of [2,{args.cells + 1}] per step",
                                    "none": "none (single rank)"}[sim.exchange_mode])},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "checks": checks,
    }

    # ---------------- the literal host-array drop-in (rank 0, N=1 only) ----------------
    if rank == 0 and world == 1 and not args.no_host_arrays:
        try:
            line["e2e_host_arrays"] = host_array_leg(sim, n, args.cells)
        except Exception as exc:  # reported, never hidden
            line["e2e_host_arrays"] = {"error": repr(exc)}

    # ---------------- BASELINE configs[3] (rank 0, N=1 only) ----------------
    del sim
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_configs3 and not args.walls and args.cells == N_CELLS:
        try:
            line["configs3"] = configs3_leg(n, device, peak, peak_src)
            checks["configs3"] = line["configs3"]["checks"]
            checks["ok"] = checks["ok"] and line["configs3"]["checks"]["ok"]
        except Exception as exc:  # reported, never hidden
            line["configs3"] = {"error": repr(exc)}
            checks["ok"] = False
        torch.cuda.empty_cache()

    # ---------------- CPU baseline (rank 0, N=1 only, bounded sample) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu, loop
        kind = "reference" if cpu.have("reference") else "port"
        if kind == "port" and not cpu.have("port"):
            cpu.build()
        cores = os.cpu_count() or 1
        n_sample = 2_000_000
        res = loop.timed_cpu_sample(kind, n_sample, args.cells, True, steps=8, warmup=1, periodic_potential=False)
        line["cpu_baseline"] = {
            "value": res["value"], "unit": "particle-steps/s", "cores": cores, "kind": kind,
            "sample": f"{cores} processes x ({n_sample:.0e} ions + {n_sample:.0e} electrons), {args.cells}-cell periodic "
                      f"grid, 8 timed steps after 1 warm-up ({res['seconds']:.1f} s)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not checks["ok"]:
        sys.stderr.write("bench.py: self-checks FAILED: " + json.dumps(checks) + "\n")
        sys.exit(1)


def configs3_leg(n, device, peak, peak_src):
    """BASELINE configs[3] as the reference would run it: 65536 cells, absorbing walls with injection (the
    reference's default boundary), the electron phase-space dimple, hole tracking every step; 1e8+1e8
    particles.  Timed over (at least) two electron sort intervals, sorts inside the timed region."""
    import torch

    from espic_hole_tracking_b200.simulation import SimConfig, Simulation
    cells = 65536
    cfg = SimConfig(n_steps=1200, n_particles=n, n_cells=cells, periodic_particles=False, extra_storage_factor=1.1,
                    track_hole=True, initial_transient_steps=0, create_electron_dimple=True)
    sim = Simulation(cfg, device=device)
    sim.init_synthetic()
    sim.initialize()
    k_e = max(int(sim.sort_intervals["electrons"]), 1)
    warm = 100   # past the first, most violent part of the dimple's start-up transient
    steps = max(4 * k_e, 40)   # four whole electron sort intervals
    # start the timed region right after an electron sort so that it holds whole intervals
    while (sim.k + warm) % k_e != 0:
        warm += 1
    # Pass A, throughput: the public driver call Simulation.run() -- stretches of steps from one host call each,
    # replayed from step graphs (espic_pic_run), sorts on schedule -- with nothing recorded between the kernels.
    sim.run(warm, check_every=0)          # warm-up through the same call: the step graphs exist before the clock starts
    torch.cuda.synchronize()
    live0 = sim.n_active()
    ctx = sim.ctx
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    sim.run(steps, check_every=0)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    live1 = sim.n_active()
    # Pass B, roofline: the same number of steps (whole sort intervals again) one espic_pic_step at a time with a
    # CUDA event on either side of every mover launch -- events cannot be captured into the graphs, and they cost
    # ~20 us per step, which is why the throughput above is not taken from this pass.
    ctx.timing_enable(True)
    evb0, evb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evb0.record()
    for _ in range(steps):
        sim.step()
    evb1.record()
    torch.cuda.synchronize()
    ms_b = evb0.elapsed_time(evb1)
    timing = ctx.timing_read()
    ctx.timing_enable(False)
    live_mean = 0.5 * (live0["ions"] + live0["electrons"] + live1["ions"] + live1["electrons"])
    per_launch = 0.5 * live_mean                       # live particles one mover launch advances
    mover_ms = timing["mean_ms"]
    achieved = ALGO_BYTES_PER_PARTICLE_STEP * per_launch / (mover_ms * 1e-3) / 1e9 if mover_ms > 0 else 0.0
    window = ctx._lib.espic_mover_window_cells(ctx.handle, sim.n_points)
    chk = sim.validate()
    pos = sim.hole_relative_positions[warm:warm + 2 * steps].cpu().numpy()
    return {
        "metric": METRIC, "value": live_mean * steps / (ms * 1e-3), "unit": "particle-steps/s", "steps": steps,
        "warmup": warm, "ms_per_step": ms / steps, "ms_per_step_instrumented": ms_b / steps,
        "api": "Simulation.run(steps): espic_pic_run, steps replayed from CUDA graphs; roofline from a second, "
               "instrumented pass of Simulation.step()",
        "config": {"workload": f"{n:.3g} ions + {n:.3g} electrons, {cells}-cell wall-bounded grid with injection, electron "
                               "dimple, hole tracking every step (BASELINE configs[3])",
                   "cells": cells, "periodic": False, "sort_interval_steps": dict(sim.sort_intervals),
                   "electron_sort_intervals_timed": steps / k_e,
                   "live_particles": {"start": live0, "end": live1}},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": f"espic::k_move_win<walls> ({window}-cell shared-memory window per store segment)",
                     "launch_ms_mean": mover_ms, "launches_timed": timing["launches"],
                     "mover_share_of_step": timing["total_ms"] / ms_b if ms_b > 0 else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PARTICLE_STEP * per_launch,
                     "note": "mean over every mover launch of the timed region (both species, all steps of the sort "
                             "intervals); algorithmic bytes count LIVE particles only, the kernel also scans dead slots"},
        "gpu_launches": int(launches), "checks": chk,
        "hole_position_last": float(pos[-1]) if len(pos) else None,
    }


def host_array_leg(sim, n, cells):
    """move_particles_c / poisson_solve_c with HOST numpy arrays: one full timestep, timed by wall clock."""
    import ctypes as C

    import numpy as np
    import torch

    from espic_hole_tracking_b200._lib import lib
    L = lib()
    p = lambda a: C.c_void_p(a.ctypes.data)
    grid = sim.grid.cpu().numpy()
    mask = np.zeros_like(grid)
    phi = sim.potential.cpu().numpy()
    stores = []
    for s in (sim.ions, sim.electrons):
        stores.append(dict(p=s.particles.cpu().numpy(), stack=s.empty_slots.cpu().numpy(),
                           top=C.c_int(s.current_empty_slot[0]), last=s.largest_index[0], qm=s.charge_to_mass))
    den = [np.zeros_like(grid), np.zeros_like(grid)]
    bytes_h2d = bytes_d2h = 0

    def one_step():
        nonlocal bytes_h2d, bytes_d2h
        for st, d in zip(stores, den):
            S = st["p"].shape[1]
            L.move_particles_c(p(grid), p(mask), p(phi), sim.dt, st["qm"], sim.background_density, st["last"],
                               p(st["p"]), p(d), p(st["stack"]), C.cast(C.byref(st["top"]), C.c_void_p), 0.0, 1, 1,
                               len(st["stack"]) - 1, len(grid), S)
            # x and v rows both ways, the three grid arrays in, the density out; the free-slot stack is
            # append-only, so only entries pushed by this call come back (none in a periodic run)
            top_before = st.get("top_before", st["top"].value)
            bytes_h2d += 8 * (st["last"] + 1) + 12 * len(grid) + 4
            bytes_d2h += 8 * (st["last"] + 1) + 4 * len(grid) + 4 + 4 * max(0, st["top"].value - top_before)
            st["top_before"] = st["top"].value
        d0, d1 = den
        d0[0] += d0[-1]; d0[-1] = d0[0]
        d1[0] += d1[-1]; d1[-1] = d1[0]
        rho = d0 - d1
        L.poisson_solve_c(p(grid), p(mask), p(rho), sim.cfg.debye_length, p(phi), -3.0, 1.0, 0, 0, len(grid))

    one_step()  # warm-up (allocates the staging buffers, page-locks the arrays)
    reps = 3
    times = []
    for _ in range(reps):   # wall clock over PCIe is noisy (host NUMA placement, other tenants): the best step counts
        bytes_h2d = bytes_d2h = 0
        t0 = time.perf_counter()
        one_step()
        times.append(time.perf_counter() - t0)
    dt = min(times)
    bytes_h2d *= reps
    bytes_d2h *= reps
    L.espic_host_release()   # the numpy arrays above die with this function: give their pages back first
    return {"value": 2.0 * n / dt, "seconds_per_step": times, "unit": "particle-steps/s", "h2d_bytes_per_step": bytes_h2d // reps,
            "d2h_bytes_per_step": bytes_d2h // reps, "steps": reps,
            "note": "reference-named C symbols with host numpy arrays (page-locked by the library on first use; upload, "
                    "push and download of the store pipelined in 8M-slot chunks on three streams); PCIe-bound by "
                    "construction"}


if __name__ == "__main__":
    main()
