// K1: fused gather / leapfrog push / boundary handling / linear-weight deposit.
// Replaces move_particles_c (ref infMagSim_c.c:121-229).
//
// Data layout in HBM: structure of arrays, float x[S] followed by float v[S] (the reference's
// [2,S] store), read and written as float4 (one quad of 4 consecutive slots per lane, 512 B
// contiguous per warp instruction), streaming cache hints.  Algorithmic traffic: 16 B per
// particle-step (x,v in; x,v out).
//
// Per CTA the whole grid lives in shared memory as three tables:
//   s_dv[c]  the velocity kick of cell c, fl(fl(fl(q/m*E_c) - a_b)*dt) with
//            E_c = fl(-(phi[c+1]-phi[c])/dz): bit-identical to what the reference evaluates
//            per particle (piecewise-constant field, ref :166-168), so the gather is one LDS;
//   s_lo/s_hi  64-bit fixed-point density accumulators split into two u32 words so that the
//            deposit is a native ATOMS.ADD (sm_100a has no native shared-memory fp32 or u64
//            atomic add: both compile to CAS loops).  A carry into s_hi is detected from the
//            value the atomic returns.  Integer accumulation makes the deposited density
//            independent of summation order, hence reproducible run to run and identical
//            for any particle ordering.
// The tile is flushed once per CTA with 64-bit global reductions into ctx->acc; k_finish
// converts to the float density the reference's callers expect.
//
// Work split: the store is cut into units of 256 consecutive slots.  The wall variant gives every
// warp of the persistent grid (148 CTAs x 28 warps, register loads) one contiguous, ascending range of
// units; the periodic variant (148 CTAs x 16 warps) deals the units round-robin (warp w: units w,
// w+2368, ...), which keeps the whole grid on one contiguous stretch of the store at any moment (DRAM
// page locality), and brings each unit into a per-warp shared-memory ring as two 1 KB bulk copies (TMA,
// cp.async.bulk completing on an mbarrier) issued by one lane, one unit ahead of the arithmetic.
//
// Removal (walls, absorbing object; in periodic runs only a particle flung past 0.99*2*z_max)
// must reproduce the reference's sequential stack pushes: slots appear on the stack in
// ascending slot order.  A range -- a warp's range of units, or a single unit in the
// round-robin split -- compacts its removed slot ids in order with ballots into the staging
// region behind its first slot and records the count; k_finish scans the counts (skipped when
// nothing was removed) and k_scatter copies the staged ids to their final stack positions -- on the
// step path (espic_pic_step, deposit left in the caller's accumulators) one launch does both,
// k_removal_epilogue, and the periodic variant pushes its practically-never removals itself.
//
// Grids too large for the shared-memory tables (more than ~19k points) are moved by
// k_move_win (mover_win.cu); start-up calls on such grids use this kernel with the kick table
// in global memory (k_build_dv) and 64-bit global reductions for the deposit.
#include <cstdio>
#include <cstdlib>

#include "mover_common.cuh"

namespace espic {


// Removal bookkeeping of the round-robin (per-unit ranges) periodic variant by ONE CTA, the last one to finish
// (MoveArgs::epilogue): nothing to do unless a slot was removed, which a periodic run practically never does;
// then the per-unit counts are scanned and the staged ids pushed on the stack in ascending slot order, as
// k_finish + k_scatter would.  Out of line: none of this may weigh on the register allocation of the main loop.
__device__ __noinline__ void removal_epilogue_strided(const MoveArgs& a, int n_slots) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int total = __ldcg(a.total_removed);
    if (total != 0) {  // uniform
        const int n_chunks = (n_slots >> 2) / kUnitQuads + 1;  // the units + the tail range
        const int per = (n_chunks + (int)blockDim.x - 1) / (int)blockDim.x;
        const int i0 = min(tid * per, n_chunks), i1 = min(i0 + per, n_chunks);
        int sum = 0;
        for (int i = i0; i < i1; ++i) sum += __ldcg(a.counts + i);
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        int before = 0;
        for (int ww = 0; ww < w; ++ww) before += s_warp[ww];
        int run = before + incl - sum;
        const int old_top = *a.top;
        for (int i = i0; i < i1; ++i) {
            const int cnt = __ldcg(a.counts + i);
            if (cnt) {
                const int* src = a.staging + (size_t)i * kUnitSlots;
                for (int q = 0; q < cnt; ++q) {
                    const int dst = old_top + 1 + run + q;
                    if (dst >= 0 && dst <= a.largest_allowed_slot) a.empty_slots[dst] = __ldcg(src + q) + a.slot_offset;
                }
                a.counts[i] = 0;  // per-unit counts are kept zeroed between launches
                run += cnt;
            }
        }
        __syncthreads();
        if (tid == 0) {
            *a.top = old_top + total;
            if (old_top + total > a.largest_allowed_slot) atomicOr(a.status, ESPIC_ST_STACK_OVERFLOW);
            *a.total_removed = 0;
        }
    }
    if (tid == 0) *a.ticket = 0u;
}

// FAST: complete units run the register fast path on float4 quads; everything else (the slots behind the last
// complete unit, unaligned stores, start-up calls with update_position=0, an absorbing
// object, exotic dz) goes quad by quad through quad_generic.
template <bool PERIODIC, bool SMEM, bool STRIDED>
__global__ void __launch_bounds__((move_threads<PERIODIC, SMEM>()), 1) k_move(const MoveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_points = a.n_points;
    const int n_pad = (n_points + 3) & ~3;
    float* s_dv = reinterpret_cast<float*>(smem_raw);
    unsigned* s_lo = reinterpret_cast<unsigned*>(s_dv + n_pad);
    unsigned* s_hi = s_lo + n_pad;

    const Geom g = make_geom(a.grid, n_points);
    const int tid = threadIdx.x;
    bool has_obj;
    const float* dv;
    if (SMEM) {
        int obj = 0;
        for (int j = tid; j < n_points; j += blockDim.x) {
            if (j < n_points - 1) {
                // ref :166-168, hoisted from the particle loop to once per cell
                const float e_cell = __fdiv_rn(-__fsub_rn(__ldg(a.phi + j + 1), __ldg(a.phi + j)), g.dz);
                const float kick = __fsub_rn(__fmul_rn(a.qm, e_cell), a.a_b);
                s_dv[j] = __fmul_rn(kick, a.dt);
            }
            s_lo[j] = 0u;
            s_hi[j] = 0u;
            obj |= (__ldg(a.mask + j) > 0.f);
        }
        has_obj = __syncthreads_or(obj) != 0;
        dv = s_dv;
    } else {
        has_obj = a.flags_global[0] != 0;
        dv = a.dv_global;
    }

    const int lane = tid & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int gw = (tid >> 5) * gridDim.x + blockIdx.x;  // interleave the warps of all CTAs
    const int total_warps = warps_per_cta * gridDim.x;
    const int n_slots = (a.largest_index_dev ? *a.largest_index_dev : a.largest_index) + 1;
    const int n_quads = (n_slots + 3) >> 2;
    const Split split(n_slots, total_warps);
    const bool fast = a.vec_ok && a.update_position != 0 && !has_obj && g.fast_div;  // uniform

    {
        int removed = 0;
        const int u0 = split.first_unit(gw), u1 = u0 + split.unit_count(gw);
        int* stage_ids = a.staging + (size_t)u0 * kUnitSlots;
        if (fast && a.staged) {
            // Units travel global -> shared with cp.async one unit ahead of the arithmetic (no
            // registers held across the wait), 2 stages of 2 KB per warp behind the grid tables:
            // [stage][x of quad 0 | v of quad 0 | x of quad 1 | v of quad 1][32 lanes x 16 B].
            float4* ring = reinterpret_cast<float4*>(smem_raw + a.tile_bytes) + (size_t)(tid >> 5) * (kRingBytesPerWarp / 16);
            const float4* gx = reinterpret_cast<const float4*>(a.px);
            const float4* gv = reinterpret_cast<const float4*>(a.pv);
#if ESPIC_VAR_TMA
            // A/B: the unit travels as two 1 KB bulk copies (x, v) issued by lane 0 and completed on a per-warp,
            // per-stage mbarrier; ring stage = [x of the unit | v of the unit]
            const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem_raw + a.tile_bytes + (size_t)(blockDim.x >> 5) * kRingBytesPerWarp) +
                                  (unsigned)(tid >> 5) * (8u * kTmaStages);
            const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
            if (lane == 0) {
#pragma unroll
                for (int st = 0; st < kTmaStages; ++st) mbar_init(bar0 + 8u * st, 1u);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
            auto issue = [&](int u, int stage) {
                __syncwarp();   // every lane has finished reading this stage
                if (lane == 0) {
                    const unsigned bar = bar0 + 8u * stage, dst = ring_s + 2048u * stage;
                    mbar_expect_tx(bar, 2048u);
                    bulk_load(dst, a.px + (size_t)u * kUnitSlots, 1024u, bar);
                    bulk_load(dst + 1024u, a.pv + (size_t)u * kUnitSlots, 1024u, bar);
                }
            };
            unsigned parity = 0u;
#else
            auto issue = [&](int u, int stage) {
                float4* dst = ring + stage * 128 + lane;
                const int q = u * kUnitQuads + lane;
                cp_async16(dst, gx + q);
                cp_async16(dst + 32, gv + q);
                cp_async16(dst + 64, gx + q + 32);
                cp_async16(dst + 96, gv + q + 32);
                cp_async_commit();
            };
#endif
            // strided: units dealt round-robin, so that at any moment the grid works on one contiguous
            // ~10 MB stretch of each row (DRAM page locality: 0.305 -> 0.295 ms per 1e8 particles)
            // instead of 4736 streams 340 KB apart
            const int du = STRIDED ? total_warps : 1;
            const int ub = STRIDED ? gw : u0, ue = STRIDED ? split.n_units : u1;
#if ESPIC_VAR_TMA
#pragma unroll
            for (int st = 0; st < kTmaStages - 1; ++st)
                if (ub + st * du < ue) issue(ub + st * du, st);
#else
            if (ub < ue) issue(ub, 0);
#endif
            int stage = 0;
#pragma unroll 1
            for (int u = ub; u < ue; u += du, stage = (ESPIC_VAR_TMA ? (stage + 1 == kTmaStages ? 0 : stage + 1) : stage ^ 1)) {
                const int q0 = u * kUnitQuads + lane, q1 = q0 + 32;
#if ESPIC_VAR_TMA
                mbar_wait(bar0 + 8u * stage, (parity >> stage) & 1u);
                parity ^= 1u << stage;   // the stage's next fill completes the other phase
                const float4* src = ring + stage * 128 + lane;
                Quad p0, p1;
                unpack_quad(src[0], src[64], p0);
                unpack_quad(src[32], src[96], p1);
                {   // refill the stage read in the previous iteration
                    const int ahead = u + (kTmaStages - 1) * du;
                    const int fill = stage == 0 ? kTmaStages - 1 : stage - 1;
                    if (ahead < ue) issue(ahead, fill);
                }
#else
                cp_async_wait_all();
                const float4* src = ring + stage * 128 + lane;
                Quad p0, p1;
                unpack_quad(src[0], src[32], p0);
#if !ESPIC_VAR_UNPACK
                unpack_quad(src[64], src[96], p1);
#endif
#endif
#if !ESPIC_VAR_TMA
                if (u + du < ue) issue(u + du, stage ^ 1);  // uniform; fills the OTHER stage
#endif
                if (lane < 16 && u + a.prefetch_units * du < split.n_units) {
                    const float* row = (lane & 8) ? a.pv : a.px;
                    const float* line = row + (size_t)(u + a.prefetch_units * du) * kUnitSlots + (lane & 7) * 32;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
                }
                const unsigned r0 = quad_fast<PERIODIC, SMEM>(a, g, p0, q0, n_slots, dv, s_lo, s_hi);
                // the second quad is read from the ring only now: its stage is not overwritten before the next
                // iteration, and nothing of it has to survive the first quad in registers (or in local memory:
                // both quads held at once spilled 16 words per iteration at 64 registers per thread)
#if ESPIC_VAR_UNPACK
                unpack_quad(src[64], src[96], p1);
#endif
                const unsigned r1 = quad_fast<PERIODIC, SMEM>(a, g, p1, q1, n_slots, dv, s_lo, s_hi);
                if (!PERIODIC || __any_sync(0xffffffffu, (r0 | r1) != 0u)) {
                    if (STRIDED) {  // every unit is its own range: staged behind the unit, counted per unit
                        int removed_here = 0;
                        int* stage_unit = a.staging + (size_t)u * kUnitSlots;
                        stage_removed(r0, q0, lane, stage_unit, removed_here);
                        stage_removed(r1, q1, lane, stage_unit, removed_here);
                        if (lane == 0 && removed_here) {
                            a.counts[u] = removed_here;
                            atomicAdd(a.total_removed, removed_here);
                        }
                    } else {
                        stage_removed(r0, q0, lane, stage_ids, removed);
                        stage_removed(r1, q1, lane, stage_ids, removed);
                    }
                }
            }
        } else if (fast) {
#pragma unroll 1
            for (int u = u0; u < u1; ++u) {
                const int q0 = u * kUnitQuads + lane, q1 = q0 + 32;
                // pull a later unit of this warp's range into L2 (16 lines: 8 of x, 8 of v), so that
                // the register loads below see L2 latency instead of DRAM latency
                if (lane < 16 && u + a.prefetch_units < split.n_units) {
                    const float* row = (lane & 8) ? a.pv : a.px;
                    const float* line = row + (size_t)(u + a.prefetch_units) * kUnitSlots + (lane & 7) * 32;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
                }
                Quad p0, p1;
                load_quad(a, q0, p0);  // both quads' loads are issued before either is used
                load_quad(a, q1, p1);
                const unsigned r0 = quad_fast<PERIODIC, SMEM>(a, g, p0, q0, n_slots, dv, s_lo, s_hi);
                const unsigned r1 = quad_fast<PERIODIC, SMEM>(a, g, p1, q1, n_slots, dv, s_lo, s_hi);
                if (!PERIODIC || __any_sync(0xffffffffu, (r0 | r1) != 0u)) {
                    stage_removed(r0, q0, lane, stage_ids, removed);
                    stage_removed(r1, q1, lane, stage_ids, removed);
                }
            }
        } else if (STRIDED) {  // the host chose per-unit ranges; the device found no fast path (object, start-up call)
#pragma unroll 1
            for (int u = gw; u < split.n_units; u += total_warps) {
                int removed_here = 0;
                int* stage_unit = a.staging + (size_t)u * kUnitSlots;
#pragma unroll 1
                for (int q = u * kUnitQuads + lane; q < (u + 1) * kUnitQuads; q += 32) {
                    const unsigned r = quad_generic<PERIODIC, SMEM>(a, q, n_slots, dv, has_obj, s_lo, s_hi);
                    stage_removed(r, q, lane, stage_unit, removed_here);
                }
                if (lane == 0 && removed_here) {
                    a.counts[u] = removed_here;
                    atomicAdd(a.total_removed, removed_here);
                }
            }
        } else {
#pragma unroll 1
            for (int q = u0 * kUnitQuads + lane; q < u1 * kUnitQuads; q += 32) {
                const unsigned r = quad_generic<PERIODIC, SMEM>(a, q, n_slots, dv, has_obj, s_lo, s_hi);
                stage_removed(r, q, lane, stage_ids, removed);
            }
        }
        if (lane == 0 && !STRIDED) {
            a.counts[gw] = removed;
            if (removed) atomicAdd(a.total_removed, removed);
        }
    }
    if (gw == total_warps - 1) {
        // the slots behind the last complete unit (fewer than 256 + 3): generic, bounds-checked
        int removed = 0;
        int* stage = a.staging + (size_t)split.n_units * kUnitSlots;
#pragma unroll 1
        for (int qb = split.n_units * kUnitQuads; qb < n_quads; qb += 32) {  // uniform trip count
            const int q = qb + lane;
            unsigned r = 0;
            if (q < n_quads) r = quad_generic<PERIODIC, SMEM>(a, q, n_slots, dv, has_obj, s_lo, s_hi);
            stage_removed(r, q, lane, stage, removed);
        }
        if (lane == 0) {
            a.counts[STRIDED ? split.n_units : total_warps] = removed;
            if (removed) atomicAdd(a.total_removed, removed);
        }
    }

    if (SMEM) {
        __syncthreads();
        // flush the tile; start each CTA at a different node so the L2 reductions of
        // concurrent CTAs do not all hit the same address at once
        const int start = (int)(((long long)blockIdx.x * n_points) / gridDim.x);
        for (int j = tid; j < n_points; j += blockDim.x) {
            int jj = j + start;
            if (jj >= n_points) jj -= n_points;
            const unsigned lo = s_lo[jj], hi = s_hi[jj];
            if (lo | hi) atomicAdd(a.acc + jj, ((unsigned long long)hi << 32) | lo);
        }
    }
    if (ESPIC_VAR_EPILOGUE && STRIDED && a.epilogue) {  // uniform
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            removal_epilogue_strided(a, n_slots);
        }
    }
}

// Kick table + object flag for the large-grid path.
__global__ void k_build_dv(const float* __restrict__ grid, const float* __restrict__ mask,
                           const float* __restrict__ phi, int n_points, float qm, float a_b, float dt,
                           float* __restrict__ dv, int* __restrict__ flags) {
    const Geom g = make_geom(grid, n_points);
    int obj = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_points; j += gridDim.x * blockDim.x) {
        if (j < n_points - 1) {
            float e_cell = __fdiv_rn(-__fsub_rn(phi[j + 1], phi[j]), g.dz);
            float kick = __fsub_rn(__fmul_rn(qm, e_cell), a_b);
            dv[j] = __fmul_rn(kick, dt);
        }
        obj |= (mask[j] > 0.f);
    }
    if (__syncthreads_or(obj) && threadIdx.x == 0) atomicOr(flags, 1);
}

// The same for both species of a step in one launch: table / flag pair A for the species pushed first, B behind it.
__global__ void k_build_dv2(const float* __restrict__ grid, const float* __restrict__ mask, int n_points,
                            const float* __restrict__ phi_a, float qm_a, float a_b_a, float dt_a, float* __restrict__ dv_a,
                            const float* __restrict__ phi_b, float qm_b, float a_b_b, float dt_b, float* __restrict__ dv_b,
                            int* __restrict__ flags) {
    const Geom g = make_geom(grid, n_points);
    int obj = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_points; j += gridDim.x * blockDim.x) {
        if (j < n_points - 1) {
            const float ea = __fdiv_rn(-__fsub_rn(phi_a[j + 1], phi_a[j]), g.dz);
            dv_a[j] = __fmul_rn(__fsub_rn(__fmul_rn(qm_a, ea), a_b_a), dt_a);
            const float eb = __fdiv_rn(-__fsub_rn(phi_b[j + 1], phi_b[j]), g.dz);
            dv_b[j] = __fmul_rn(__fsub_rn(__fmul_rn(qm_b, eb), a_b_b), dt_b);
        }
        obj |= (mask[j] > 0.f);
    }
    if (__syncthreads_or(obj) && threadIdx.x == 0) {
        atomicOr(flags, 1);
        atomicOr(flags + 2, 1);
    }
}

struct FinishArgs {
    unsigned long long* acc;
    float* density;
    int* counts;       // in: removed per range
    int* offsets;      // out: exclusive offsets (scan of counts), only when something was removed
    int strided;       // ranges are the 256-slot units (+ the tail) instead of the warps' ranges (+ tail)
    const int* largest_index_dev;
    int largest_index;
    int* top;          // device stack top
    int* scatter_hdr;  // [0] = old top, [1] = total removed, [2] = n_chunks, [3] = running total (input)
    int* status;
    int n_ranges;       // mover warps + 1 (see Split)
    int largest_allowed_slot;
    int n_points;
    double inv_scale;   // 2^-24 / background_density
    int* flags_global;  // [0] object flag, [1] segment counter: reset for the next large-grid launch
    int convert;        // 0: leave the accumulators as they are (a later launch over another part of the
                        // store adds to them and converts: chunked host-array calls)
};

// Fixed-point accumulators -> float density (and re-zero them), by all CTAs; then CTA 0:
// exclusive scan of the per-range removal counts (skipped when the mover removed nothing) and
// stack-top update.
__global__ void __launch_bounds__(1024) k_finish(const FinishArgs a) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (a.convert) {
        for (int j = blockIdx.x * blockDim.x + tid; j < a.n_points; j += gridDim.x * blockDim.x) {
            const unsigned long long v = a.acc[j];
            a.density[j] = (float)((double)v * a.inv_scale);
            a.acc[j] = 0ull;
        }
    }
    if (blockIdx.x != 0) return;  // large grids: the other CTAs only help with the conversion
    const int total = a.scatter_hdr[3];
    int n_chunks = a.n_ranges;
    if (a.strided) {
        const int n_slots = (a.largest_index_dev ? *a.largest_index_dev : a.largest_index) + 1;
        n_chunks = (n_slots >> 2) / kUnitQuads + 1;
    }
    if (total != 0) {  // uniform
        // each thread owns a contiguous run of counts: sum, block-scan the sums, write offsets
        const int per = (n_chunks + (int)blockDim.x - 1) / (int)blockDim.x;
        const int i0 = min(tid * per, n_chunks), i1 = min(i0 + per, n_chunks);
        int sum = 0;
#pragma unroll 8
        for (int i = i0; i < i1; ++i) sum += a.counts[i];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        if (w == 0) {
            int x = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += t;
            }
            s_warp[lane] = x;
        }
        __syncthreads();
        int run = (w ? s_warp[w - 1] : 0) + incl - sum;
#pragma unroll 8
        for (int i = i0; i < i1; ++i) {
            a.offsets[i] = run;
            run += a.counts[i];
        }
    }
    if (tid == 0) {
        const int old_top = *a.top;
        a.scatter_hdr[0] = old_top;
        a.scatter_hdr[1] = total;
        a.scatter_hdr[2] = n_chunks;
        a.scatter_hdr[3] = 0;
        *a.top = old_top + total;
        if (old_top + total > a.largest_allowed_slot) atomicOr(a.status, ESPIC_ST_STACK_OVERFLOW);
        if (a.flags_global) {
            a.flags_global[0] = 0;
            a.flags_global[1] = 0;
        }
    }
}

// Copies the staged slot ids to their final positions on the stack.  One thread per OUTPUT entry: it
// finds its range by bisection of the scanned offsets.  Removals concentrate in a few ranges (on
// a sorted store everything that leaves sits next to a wall; a range there can hold thousands of
// ids while most hold none), so a warp or CTA per range leaves the work to a handful of warps --
// 40-120 us on the 65536-cell runs; per entry it is balanced and the stack is written contiguously.
__global__ void __launch_bounds__(256) k_scatter(const int* __restrict__ staging,
                                                 const int* __restrict__ offsets,
                                                 const int* __restrict__ hdr, int* __restrict__ empty_slots,
                                                 int largest_allowed_slot, const int* largest_index_dev,
                                                 int largest_index, int mover_warps, int* strided_counts,
                                                 int slot_offset, int interleave) {
    const int total = hdr[1];
    if (total == 0) return;
    const int old_top = hdr[0], n_chunks = hdr[2];
    const int n_slots = (largest_index_dev ? *largest_index_dev : largest_index) + 1;
    const Split split(n_slots, mover_warps);
    const int stride = gridDim.x * blockDim.x;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < total; d += stride) {
        // Which staged id goes to stack position old_top+1+d.  The reference's order is ascending slot order
        // (i = d).  interleave (stores kept sorted by position, together with the by-side injection of
        // inject.cu): the ids are taken alternately from the two ends of the ascending list, outermost on top --
        // top: highest slot, below it: lowest slot, then second highest, second lowest ... -- so that the next
        // injection, which pops from the top, gets the slots next to either wall first and what is left over
        // (more particles left than entered) are the innermost slots of both wall regions, at the bottom.
        const int t = total - 1 - d;  // distance from the new top
        const int i = !interleave ? d : ((t & 1) ? (t >> 1) : total - 1 - (t >> 1));
        // last range c with offsets[c] <= i (ranges without removals share their successor's offset)
        int lo = 0, hi = n_chunks - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(offsets + mid) <= i) lo = mid;
            else hi = mid - 1;
        }
        const int c = lo;
        const int* src = strided_counts
                             ? staging + (size_t)c * kUnitSlots
                             : staging + (size_t)(c < mover_warps ? split.first_unit(c) : split.n_units) * kUnitSlots;
        const int dst = old_top + 1 + d;
        if (dst >= 0 && dst <= largest_allowed_slot) empty_slots[dst] = src[i - __ldg(offsets + c)] + slot_offset;
    }
    // per-unit counts are only ever written when non-zero: leave them zeroed for the next launch
    if (strided_counts)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride)
            if (strided_counts[c]) strided_counts[c] = 0;
}

// k_finish + k_scatter in one launch, for the step path (fused deposit: nothing to convert) with contiguous ranges
// (the wall-bounded whole-grid kernel and the windowed kernel): every CTA scans the per-range removal counts into
// shared memory for itself -- a few thousand ints out of L2 -- and copies its share of the staged ids to the stack,
// bisecting the shared-memory offsets; the last CTA to finish moves the stack top and resets the counters.  One
// launch less per species and step.
constexpr int kEpilogueMaxRanges = 12 * 1024 - 64;  // 48 KB of static shared memory

__global__ void __launch_bounds__(256) k_removal_epilogue(const int* __restrict__ staging, const int* __restrict__ counts,
                                                          int n_ranges, int* __restrict__ hdr, int* __restrict__ top,
                                                          int* __restrict__ empty_slots, int largest_allowed_slot,
                                                          const int* largest_index_dev, int largest_index, int mover_warps,
                                                          int slot_offset, int interleave, int* flags_global, int* status,
                                                          unsigned* ticket) {
    __shared__ int s_off[kEpilogueMaxRanges];
    __shared__ int s_warp[8];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int total = hdr[3];   // the movers' running total; every CTA reads it and the old top before the last one resets them
    const int old_top = *top;
    if (total != 0) {  // uniform
        // counts -> shared memory with every load in flight at once (coalesced), then an in-place exclusive scan,
        // each thread a contiguous run (an odd run length keeps the runs on different banks)
#pragma unroll 8
        for (int i = tid; i < n_ranges; i += 256) s_off[i] = __ldcg(counts + i);
        __syncthreads();
        const int per = ((n_ranges + 255) / 256) | 1;
        const int i0 = min(tid * per, n_ranges), i1 = min(i0 + per, n_ranges);
        int sum = 0;
        for (int i = i0; i < i1; ++i) sum += s_off[i];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        int run = incl - sum;
        for (int ww = 0; ww < w; ++ww) run += s_warp[ww];
        for (int i = i0; i < i1; ++i) {
            const int c = s_off[i];
            s_off[i] = run;
            run += c;
        }
        __syncthreads();
        const int n_slots = (largest_index_dev ? *largest_index_dev : largest_index) + 1;
        const Split split(n_slots, mover_warps);
        const int stride = gridDim.x * blockDim.x;
        for (int d = blockIdx.x * blockDim.x + tid; d < total; d += stride) {
            // which staged id goes to stack position old_top+1+d: see k_scatter
            const int t = total - 1 - d;
            const int i = !interleave ? d : ((t & 1) ? (t >> 1) : total - 1 - (t >> 1));
            int lo = 0, hi = n_ranges - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_off[mid] <= i) lo = mid;
                else hi = mid - 1;
            }
            const int* src = staging + (size_t)(lo < mover_warps ? split.first_unit(lo) : split.n_units) * kUnitSlots;
            const int dst = old_top + 1 + d;
            if (dst >= 0 && dst <= largest_allowed_slot) empty_slots[dst] = __ldcg(src + (i - s_off[lo])) + slot_offset;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && tid == 0) {
        *top = old_top + total;
        if (old_top + total > largest_allowed_slot) atomicOr(status, ESPIC_ST_STACK_OVERFLOW);
        hdr[0] = old_top;
        hdr[1] = total;
        hdr[2] = n_ranges;
        hdr[3] = 0;
        if (flags_global) {
            flags_global[0] = 0;
            flags_global[1] = 0;
        }
        *ticket = 0u;
    }
}

// Fixed-point accumulators -> float density (the conversion k_finish makes), accumulators cleared.
__global__ void k_convert_density(unsigned long long* __restrict__ acc, float* __restrict__ density, int n,
                                  double inv_scale) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        density[j] = (float)((double)acc[j] * inv_scale);
        acc[j] = 0ull;
    }
}

// Both species' pending planes in one launch (either may be null).
__global__ void k_convert_density2(unsigned long long* __restrict__ acc0, float* __restrict__ den0,
                                   unsigned long long* __restrict__ acc1, float* __restrict__ den1, int n, double inv_scale) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        if (acc0) {
            den0[j] = (float)((double)acc0[j] * inv_scale);
            acc0[j] = 0ull;
        }
        if (acc1) {
            den1[j] = (float)((double)acc1[j] * inv_scale);
            acc1[j] = 0ull;
        }
    }
}

int launch_convert_density2(espic_ctx* ctx, unsigned long long* acc0, float* den0, unsigned long long* acc1, float* den1,
                            int n_points, float bg) {
    if (n_points < 1 || (!acc0 && !acc1)) return 0;
    k_convert_density2<<<(n_points + 1023) / 1024, 1024, 0, ctx->stream>>>(acc0, den0, acc1, den1, n_points,
                                                                          1.0 / ((double)kWeightOne * (double)bg));
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_convert_density2 launch");
}

int launch_convert_density(espic_ctx* ctx, unsigned long long* acc, float* density, int n_points, float bg) {
    if (n_points < 1) return 0;
    k_convert_density<<<(n_points + 1023) / 1024, 1024, 0, ctx->stream>>>(acc, density, n_points,
                                                                         1.0 / ((double)kWeightOne * (double)bg));
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_convert_density launch");
}

__global__ void k_shift_velocities(float* __restrict__ v, int n, double v_b) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        v[i] = (float)((double)v[i] - v_b);  // ref infMagSim_cython.py:174-176
}

static size_t move_smem_bytes(int n_points) {
    return (size_t)((n_points + 3) & ~3) * kSmemBytesPerPoint;
}

int move_window_cells(espic_ctx* ctx, int n_points) {
    int max_smem = 0;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device) != cudaSuccess) return 0;
    if (move_smem_bytes(n_points) <= (size_t)max_smem) return 0;
    return move_window_size(n_points);
}

template <bool PERIODIC, bool SMEM, bool STRIDED = false>
static int launch_move_variant(espic_ctx* ctx, const MoveArgs& a, size_t smem) {
    auto kern = k_move<PERIODIC, SMEM, STRIDED>;
    // CTA width and CTAs per SM for this tile size, chosen once per context and instantiation (the
    // shared-memory opt-in belongs to the device's copy of the kernel): as many 256-thread CTAs as
    // the tile allows, else wider CTAs so that an SM still holds 32 warps
    espic_ctx::MoveCfg& cfg = ctx->move_cfg[(PERIODIC ? 4 : 0) + (SMEM ? 2 : 0) + (STRIDED ? 1 : 0)];
    size_t& cached_smem = cfg.smem;
    int& cached_per_sm = cfg.per_sm;
    int& cached_threads = cfg.threads;
    if (cached_smem != smem) {
        if (smem > 48 * 1024)
            ESPIC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        constexpr int max_threads = move_threads<PERIODIC, SMEM>();
        int threads = max_threads, per_sm = 0;
        if (const char* e = getenv("ESPIC_MOVE_THREADS")) {  // tuning knob
            const int t = atoi(e);
            if (t >= 256 && t <= max_threads && t % 128 == 0) threads = t;
        }
        for (;;) {
            ESPIC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
            if (per_sm * threads >= 1024 || threads * 2 > max_threads) break;
            threads *= 2;
        }
        if (per_sm < 1) return fail(ctx, ESPIC_E_ARG, "mover kernel does not fit on an SM (smem %zu)", smem);
        if (per_sm * threads > 1024) per_sm = 1024 / threads;
        cached_smem = smem;
        cached_per_sm = per_sm;
        cached_threads = threads;
    }
    const int grid = ctx->sm_count * cached_per_sm;
    ctx->mover_grid = grid;
    ctx->mover_block = cached_threads;
    ctx->mover_ranges = grid * (cached_threads / 32);
    const bool timed = ctx->timing_on && ctx->ev_used + 2 <= ctx->ev_cap;
    if (timed) cudaEventRecord(ctx->ev[ctx->ev_used], ctx->stream);
    kern<<<grid, cached_threads, smem, ctx->stream>>>(a);
    if (timed) {
        cudaEventRecord(ctx->ev[ctx->ev_used + 1], ctx->stream);
        ctx->ev_used += 2;
    }
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_move launch");
}

int launch_move(espic_ctx* ctx, const float* grid, const float* mask, const float* phi, float dt,
                float qm, float bg, int largest_index, const int* largest_index_dev, float* particles,
                float* density, int* empty_slots, int* top, float a_b, int update_position,
                int periodic, int largest_allowed_slot, int n_points, int storage, int minimal, int slot_offset,
                int finish_density, unsigned long long* acc_fused, int* pushed, int stack_interleave,
                const DvPlan* dv_plan) {
    // acc_fused (the step path, espic_pic_step): the deposit is ADDED to these caller-owned fixed-point
    // accumulators and left there -- the next field solve converts it -- and `density` is not touched.
    // slot_offset / finish_density: the call covers slots [slot_offset, slot_offset + largest_index] of the
    // store only; removed slots are reported with their store-wide numbers, and the density is converted
    // (and the accumulators cleared) only when finish_density is set -- a store processed in several
    // consecutive calls (the pipelined host-array drop-in) deposits exactly what one call would.
    if (n_points < 2) return fail(ctx, ESPIC_E_ARG, "n_points must be >= 2");
    if (slot_offset < 0 || (slot_offset & 3)) return fail(ctx, ESPIC_E_ARG, "slot_offset must be a multiple of 4");
    if (storage < 1) return fail(ctx, ESPIC_E_ARG, "particle_storage_length must be >= 1");
    if (!largest_index_dev && largest_index >= storage)
        return fail(ctx, ESPIC_E_ARG, "largest_index %d >= storage %d", largest_index, storage);
    int rc = ensure_acc(ctx, n_points);
    if (rc) return rc;
    const int max_slots = largest_index_dev ? storage : (largest_index + 1);
    const int max_quads = (max_slots + 3) >> 2;
    rc = ensure(ctx, ctx->staging, ((size_t)max_quads * 4 + 2 * kUnitSlots) * sizeof(int), "staging");
    if (rc) return rc;
    const int max_ranges = 64 * 1024;  // >= ranges of the contiguous split (warps or segments x 32) + 1
    const size_t max_units = (size_t)max_quads / kUnitQuads + 2;   // ranges of the round-robin split
    // [8: header][max_ranges + 8: counts, contiguous split][max_units: counts, round-robin split, kept
    //  zeroed between launches]; the scanned offsets live in their own buffer (its length varies from
    //  call to call and must not run into the zeroed counts)
    rc = ensure(ctx, ctx->chunk_counts, ((size_t)max_ranges + 16 + max_units) * sizeof(int), "range counts");
    if (rc) return rc;
    rc = ensure(ctx, ctx->range_offsets, ((max_units > (size_t)max_ranges ? max_units : (size_t)max_ranges) + 8) * sizeof(int),
                "range offsets");
    if (rc) return rc;

    MoveArgs a{};
    a.grid = grid;
    a.mask = mask;
    a.phi = phi;
    a.px = particles + slot_offset;
    a.pv = particles + storage + slot_offset;
    a.staging = (int*)ctx->staging.ptr;
    int* counts_contiguous = (int*)ctx->chunk_counts.ptr + 8;  // first 8 ints: scatter header
    int* counts_strided = counts_contiguous + max_ranges + 8;
    int* offsets = (int*)ctx->range_offsets.ptr;
    a.counts = counts_contiguous;
    a.total_removed = (int*)ctx->chunk_counts.ptr + 3;
    a.acc = acc_fused ? acc_fused : ctx->acc;
    a.status = ctx->status;
    a.ticket = (unsigned*)(ctx->status + 10);
    a.top = top;
    a.empty_slots = empty_slots;
    a.largest_allowed_slot = largest_allowed_slot;
    a.slot_offset = slot_offset;
    a.largest_index_dev = largest_index_dev;
    a.largest_index = largest_index;
    a.n_points = n_points;
    a.dt = dt;
    a.qm = qm;
    a.a_b = a_b;
    a.update_position = update_position;
    a.minimal = minimal;
    {
        static int debug_cta = -1;
        if (debug_cta < 0) debug_cta = getenv("ESPIC_DEBUG_CTA") ? 1 : 0;
        if (debug_cta) {
            rc = ensure(ctx, ctx->debug_ws, 2048 * sizeof(long long), "debug");
            if (rc) return rc;
            a.cta_cycles = (long long*)ctx->debug_ws.ptr;
        }
    }
    a.vec_ok = ((reinterpret_cast<uintptr_t>(particles) & 15u) == 0 && (storage & 3) == 0) ? 1 : 0;
    if (minimal) a.vec_ok = 0;  // the stripped variant (never on the step path) runs the generic slot update only
    static int prefetch_env = -2;
    if (prefetch_env == -2) {
        const char* e = getenv("ESPIC_PREFETCH_UNITS");  // tuning knob; -1 = measured default, 0 = off
        prefetch_env = e ? atoi(e) : -1;
    }

    const size_t tile = move_smem_bytes(n_points);
    int max_smem = 0;
    ESPIC_CUDA(ctx, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    const bool use_smem = tile <= (size_t)max_smem;
    // cp.async staging ring: 4 KB per warp of a 1024-thread CTA, when the tables leave room for it
    const size_t ring = (size_t)((periodic ? kMoveThreadsStaged : kMoveThreads) / 32) * (kRingBytesPerWarp + 8 * kTmaStages * ESPIC_VAR_TMA);
    static int staging_env = -1;
    if (staging_env < 0) {
        const char* e = getenv("ESPIC_STAGING");  // tuning knob
        staging_env = e ? atoi(e) : 1;
    }
    // measured: staging gains 4 % in the periodic variant and loses 5 % in the wall variant (its
    // removal bookkeeping already strains the 64-register budget), so only the former uses it
    a.staged = (use_smem && (periodic || ESPIC_VAR_STAGE_WALLS) && staging_env && tile + ring <= (size_t)max_smem && a.vec_ok) ? 1 : 0;
    a.tile_bytes = (int)tile;
    static int strided_env = -1;
    if (strided_env < 0) {
        const char* e = getenv("ESPIC_STRIDED");  // tuning knob; 0 = contiguous ranges per warp
        strided_env = e ? atoi(e) : 1;
    }
    a.strided = (a.staged && strided_env && periodic) ? 1 : 0;
    if (a.strided) a.counts = counts_strided;
    // windowed / global-table launches never take the round-robin kernel
    a.epilogue = (ESPIC_VAR_EPILOGUE && acc_fused && a.strided && periodic && use_smem) ? 1 : 0;
    a.direct_push = (!a.epilogue && acc_fused && pushed && a.strided && periodic && use_smem) ? 1 : 0;
    a.pushed = pushed;
    const size_t smem = use_smem ? tile + (a.staged ? ring : 0) : 0;
    // L2 prefetch distance, measured on B200 (1e8 particles, 4096 cells): one unit ahead of the next
    // fetch is best in both modes -- 0.373 ms without it, 0.305 ms with it (staged)
    a.prefetch_units = prefetch_env < 0 ? (a.staged ? 2 : 1) : (prefetch_env == 0 ? (1 << 28) : prefetch_env);
    int* flags = nullptr;
    if (!use_smem) {
        // [16 ints: flag pair A, flag pair B (zeroed at allocation / by the epilogues)][table A][table B]
        const size_t n_tab = ((size_t)n_points + 63) & ~(size_t)63;
        rc = ensure(ctx, ctx->dv_table, (16 + 2 * n_tab) * sizeof(float), "kick tables");
        if (rc) return rc;
        const int role = dv_plan ? dv_plan->role : 0;
        flags = (int*)ctx->dv_table.ptr + (role == 2 ? 2 : 0);
        float* dvt = (float*)ctx->dv_table.ptr + 16 + (role == 2 ? n_tab : 0);
        if (role == 1) {
            k_build_dv2<<<(n_points + 255) / 256, 256, 0, ctx->stream>>>(grid, mask, n_points, phi, qm, a_b, dt, dvt,
                                                                         dv_plan->phi_b, dv_plan->qm_b, dv_plan->a_b_b,
                                                                         dv_plan->dt_b, dvt + n_tab, flags);
            ctx->launches++;
        } else if (role == 0) {
            k_build_dv<<<(n_points + 255) / 256, 256, 0, ctx->stream>>>(grid, mask, phi, n_points, qm, a_b, dt, dvt, flags);
            ctx->launches++;
        }
        a.dv_global = dvt;
        a.flags_global = flags;
    }
    // large grid, production call: a window of the grid in shared memory per segment of the store
    static int window_env = -1;
    if (window_env < 0) {
        const char* e = getenv("ESPIC_WINDOW");  // 0: always the all-global kernel
        window_env = e ? atoi(e) : 1;
    }
    if (!use_smem && window_env && a.vec_ok && update_position && !minimal && move_window_size(n_points) > 0) {
        rc = launch_move_window(ctx, a, periodic);
    } else if (periodic && a.strided) {
        rc = launch_move_variant<true, true, true>(ctx, a, smem);
    } else if (periodic) {
        rc = use_smem ? launch_move_variant<true, true>(ctx, a, smem) : launch_move_variant<true, false>(ctx, a, 0);
    } else {
        rc = use_smem ? launch_move_variant<false, true>(ctx, a, smem) : launch_move_variant<false, false>(ctx, a, 0);
    }
    if (rc) return rc;
    if (a.epilogue || a.direct_push) return 0;  // removal bookkeeping done inside the mover; the deposit stays in acc_fused

    FinishArgs f{};
    f.acc = a.acc;
    f.density = density;
    f.counts = a.counts;
    f.offsets = offsets;
    f.strided = a.strided;
    f.largest_index_dev = largest_index_dev;
    f.largest_index = largest_index;
    f.top = top;
    f.scatter_hdr = (int*)ctx->chunk_counts.ptr;
    f.status = ctx->status;
    const int mover_warps = ctx->mover_ranges;
    if (mover_warps + 1 > max_ranges) return fail(ctx, ESPIC_E_ARG, "mover grid too large for the range table");
    f.n_ranges = mover_warps + 1;
    f.largest_allowed_slot = largest_allowed_slot;
    f.n_points = n_points;
    f.inv_scale = 1.0 / ((double)kWeightOne * (double)bg);
    f.flags_global = flags;
    f.convert = acc_fused ? 0 : finish_density;
    static int fused_epilogue_env = -1;
    if (fused_epilogue_env < 0) {
        const char* e = getenv("ESPIC_FUSED_EPILOGUE");  // tuning knob; 0 = k_finish + k_scatter
        fused_epilogue_env = e ? atoi(e) : 1;
    }
    if (fused_epilogue_env && acc_fused && !a.strided && f.n_ranges <= kEpilogueMaxRanges) {
        int eg = 1 + max_slots / 16384;   // a small store needs a few CTAs, not a scan per SM
        if (eg > 4 * ctx->sm_count) eg = 4 * ctx->sm_count;
        k_removal_epilogue<<<eg, 256, 0, ctx->stream>>>(a.staging, a.counts, f.n_ranges, f.scatter_hdr, top, empty_slots,
                                                        largest_allowed_slot, largest_index_dev, largest_index, mover_warps,
                                                        slot_offset, stack_interleave, flags, ctx->status,
                                                        (unsigned*)(ctx->status + 11));
        ctx->launches++;
        return check_cuda(ctx, cudaGetLastError(), "mover epilogue launch");
    }
    k_finish<<<n_points <= 8192 ? 1 : (n_points + 4095) / 4096, 1024, 0, ctx->stream>>>(f);
    ctx->launches++;
    ESPIC_CUDA(ctx, cudaGetLastError());
    k_scatter<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(a.staging, offsets, f.scatter_hdr, empty_slots,
                                                      largest_allowed_slot, largest_index_dev, largest_index,
                                                      mover_warps, a.strided ? a.counts : nullptr, slot_offset,
                                                      stack_interleave);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "mover epilogue launch");
}

// Every float a in [0, span]: fast quotient vs the division instruction.
__global__ void k_selftest_quotient(const float* __restrict__ grid, int n_points, unsigned long long* counts) {
    const Geom g = make_geom(grid, n_points);
    const unsigned hi = __float_as_uint(g.span);
    unsigned long long bad_cell = 0, bad_bits = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= hi; i += stride) {
        const float a = __uint_as_float((unsigned)i);
        const float qf = quotient_dz<true>(g, a);
        const float qr = __fdiv_rn(a, g.dz);
        if (__float2int_rz(qf) != __float2int_rz(qr)) ++bad_cell;
        if (a >= 1e-30f && __float_as_uint(qf) != __float_as_uint(qr)) ++bad_bits;
    }
    if (!g.fast_div) bad_cell = bad_bits = 0;  // the kernels use the division instruction for such grids
    if (bad_cell) atomicAdd(counts, bad_cell);
    if (bad_bits) atomicAdd(counts + 1, bad_bits);
}

int launch_selftest_quotient(espic_ctx* ctx, const float* grid, int n_points, unsigned long long* counts_out) {
    int rc = ensure(ctx, ctx->inject_ws, 64, "selftest workspace");
    if (rc) return rc;
    unsigned long long* d = (unsigned long long*)((char*)ctx->inject_ws.ptr + 32);
    ESPIC_CUDA(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
    k_selftest_quotient<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(grid, n_points, d);
    ctx->launches++;
    ESPIC_CUDA(ctx, cudaGetLastError());
    ESPIC_CUDA(ctx, cudaMemcpyAsync(counts_out, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    ESPIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int launch_shift_velocities(espic_ctx* ctx, float* particles, int storage, int largest_index, double v_b) {
    const int n = largest_index + 1;
    if (n <= 0) return 0;
    const int grid = (n + 255) / 256 < 8 * ctx->sm_count ? (n + 255) / 256 : 8 * ctx->sm_count;
    k_shift_velocities<<<grid, 256, 0, ctx->stream>>>(particles + storage, n, v_b);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_shift_velocities launch");
}

}  // namespace espic
