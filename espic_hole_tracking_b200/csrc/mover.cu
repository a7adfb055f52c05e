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
// Removal (non-periodic walls, absorbing object) must reproduce the reference's sequential
// stack pushes: slots appear on the stack in ascending slot order.  Each warp owns whole
// chunks of 4096 consecutive slots, compacts the removed slot ids of a chunk in order with
// ballots into that chunk's staging region, and records the count; k_finish scans the counts
// (one CTA) and k_scatter copies the staged ids to their final stack positions.
//
// Grids too large for the shared-memory tables (more than ~19k points) use the same kernel
// with the kick table in global memory (L1/L2-resident) and 64-bit global reductions for
// the deposit.
#include <cstdio>

#include "espic_internal.cuh"

namespace espic {

constexpr int kMoveThreads = 256;
constexpr int kIterPerChunk = 32;                       // warp iterations (of 32 quads) per chunk
constexpr int kQuadsPerChunk = 32 * kIterPerChunk;      // 1024 quads
constexpr int kChunkSlots = 4 * kQuadsPerChunk;         // 4096 slots
constexpr int kSmemBytesPerPoint = 12;

struct MoveArgs {
    const float* grid;
    const float* mask;
    const float* phi;
    float* px;
    float* pv;
    int* staging;
    int* counts;
    unsigned long long* acc;
    int* status;
    const int* largest_index_dev;
    const float* dv_global;   // only for the large-grid path
    const int* flags_global;  // [0] = object present (large-grid path)
    int largest_index;
    int n_points;
    float dt, qm, a_b;
    int update_position;
    int vec_ok;  // 1: float4 path (16-byte aligned rows), 0: scalar path
};

enum { ACT_NONE = 0, ACT_DEPOSIT = 1, ACT_REMOVE = 2 };

// One slot of the reference's loop body (ref infMagSim_c.c:143-217), everything except the
// deposit itself.  Bit-exact in x and v: each rounded operation of the reference is one
// rounded intrinsic here.
template <bool PERIODIC>
__device__ __forceinline__ int push_one(const Geom& g, float& x, float& v, const float* dv,
                                        float dt, bool update_position, bool has_obj,
                                        const float* __restrict__ grid,
                                        const float* __restrict__ mask, int& cell, bool& bad_any) {
    bool alive;
    if (PERIODIC) {
        alive = x < g.live_thr;
        if (alive) x = fold_periodic(g, x);
    } else {
        alive = (x > g.lo_edge) && (x < g.hi_edge);
    }
    const bool alive_at_entry = alive;
    cell = 0;
    if (alive_at_entry) {
        bool bad;
        cell = cell_of<PERIODIC>(g, x, bad);
        bad_any |= bad;
        cell = min(max(cell, 0), g.last_cell);
        v = __fadd_rn(v, dv[cell]);
        if (update_position) {
            x = __fadd_rn(x, __fmul_rn(v, dt));
            if (PERIODIC) {
                alive = x < g.live_thr;
                if (alive) x = fold_periodic(g, x);
            } else {
                alive = (x > g.lo_edge) && (x < g.hi_edge);
            }
            if (alive) {
                cell = cell_of<PERIODIC>(g, x, bad);
                bad_any |= bad;
                cell = min(max(cell, 0), g.last_cell);
            }
        }
    }
    bool swallowed = false;
    if (has_obj && alive) {  // ref :200-210, evaluated in double there
        float m0 = __ldg(mask + cell);
        if (m0 > 0.f) {
            float m1 = __ldg(mask + cell + 1);
            double covered = (m1 > 0.f) ? 1. - (double)m0 : (double)m0;
            swallowed = (double)x > (double)__ldg(grid + cell) + covered * (double)g.dz;
        }
    }
    if (alive_at_entry || x < g.live_thr) {  // ref :211
        if (!alive || swallowed) {
            x = g.parked;
            return ACT_REMOVE;
        }
        return ACT_DEPOSIT;
    }
    return ACT_NONE;
}

template <bool SMEM>
__device__ __forceinline__ void deposit_one(const Geom& g, float x, int cell, unsigned* s_lo,
                                            unsigned* s_hi, unsigned long long* acc) {
    const unsigned w0 = left_share_fixed(g, x);
    const unsigned w1 = kWeightOne - w0;
    if (SMEM) {
        unsigned old0 = atomicAdd(s_lo + cell, w0);
        unsigned old1 = atomicAdd(s_lo + cell + 1, w1);
        if (old0 + w0 < old0) atomicAdd(s_hi + cell, 1u);
        if (old1 + w1 < old1) atomicAdd(s_hi + cell + 1, 1u);
    } else {
        atomicAdd(acc + cell, (unsigned long long)w0);
        atomicAdd(acc + cell + 1, (unsigned long long)w1);
    }
}

__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

template <bool PERIODIC, bool SMEM>
__global__ void __launch_bounds__(kMoveThreads) k_move(const MoveArgs a) {
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
                float e_cell = __fdiv_rn(-__fsub_rn(__ldg(a.phi + j + 1), __ldg(a.phi + j)), g.dz);
                float kick = __fsub_rn(__fmul_rn(a.qm, e_cell), a.a_b);
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
    const int gw = (tid >> 5) * gridDim.x + blockIdx.x;  // interleave warps of all CTAs
    const int total_warps = warps_per_cta * gridDim.x;
    const int n_slots = (a.largest_index_dev ? *a.largest_index_dev : a.largest_index) + 1;
    const int n_quads = (n_slots + 3) >> 2;
    const int n_chunks = (n_quads + kQuadsPerChunk - 1) / kQuadsPerChunk;
    const bool upd = a.update_position != 0;
    const float dt = a.dt;
    bool bad_any = false;

    float4* x4 = reinterpret_cast<float4*>(a.px);
    float4* v4 = reinterpret_cast<float4*>(a.pv);

    for (int chunk = gw; chunk < n_chunks; chunk += total_warps) {
        int removed = 0;
        int* stage = a.staging + (size_t)chunk * kChunkSlots;
        const int qbase = chunk * kQuadsPerChunk;
#pragma unroll 1
        for (int it = 0; it < kIterPerChunk; it += 2) {
            const int q[2] = {qbase + it * 32 + lane, qbase + (it + 1) * 32 + lane};
            if (qbase + it * 32 >= n_quads) break;  // warp-uniform
            float X[2][4], V[2][4];
            // issue all loads of both quads before touching either (memory-level parallelism)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (q[u] < n_quads) {
                    if (a.vec_ok) {
                        float4 xx = ld_stream(x4 + q[u]);
                        float4 vv = ld_stream(v4 + q[u]);
                        X[u][0] = xx.x; X[u][1] = xx.y; X[u][2] = xx.z; X[u][3] = xx.w;
                        V[u][0] = vv.x; V[u][1] = vv.y; V[u][2] = vv.z; V[u][3] = vv.w;
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            int s = 4 * q[u] + c;
                            X[u][c] = (s < n_slots) ? a.px[s] : g.parked;
                            V[u][c] = (s < n_slots) ? a.pv[s] : 0.f;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                unsigned rem_mask = 0;
                if (q[u] < n_quads) {
                    int cell[4];
                    int act[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool in_range = (4 * q[u] + c) < n_slots;
                        act[c] = ACT_NONE;
                        cell[c] = 0;
                        if (in_range)
                            act[c] = push_one<PERIODIC>(g, X[u][c], V[u][c], dv, dt, upd, has_obj,
                                                        a.grid, a.mask, cell[c], bad_any);
                    }
                    if (a.vec_ok) {
                        st_stream(x4 + q[u], make_float4(X[u][0], X[u][1], X[u][2], X[u][3]));
                        st_stream(v4 + q[u], make_float4(V[u][0], V[u][1], V[u][2], V[u][3]));
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            int s = 4 * q[u] + c;
                            if (s < n_slots) {
                                a.px[s] = X[u][c];
                                a.pv[s] = V[u][c];
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (act[c] == ACT_DEPOSIT)
                            deposit_one<SMEM>(g, X[u][c], cell[c], s_lo, s_hi, a.acc);
                        rem_mask |= (act[c] == ACT_REMOVE) ? (1u << c) : 0u;
                    }
                }
                // ordered compaction of removed slot ids (ascending slot order)
                if (__ballot_sync(0xffffffffu, rem_mask != 0u)) {
                    const unsigned b0 = __ballot_sync(0xffffffffu, rem_mask & 1u);
                    const unsigned b1 = __ballot_sync(0xffffffffu, rem_mask & 2u);
                    const unsigned b2 = __ballot_sync(0xffffffffu, rem_mask & 4u);
                    const unsigned b3 = __ballot_sync(0xffffffffu, rem_mask & 8u);
                    const unsigned lt = (1u << lane) - 1u;
                    int pos = removed + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) +
                              __popc(b3 & lt);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (rem_mask & (1u << c)) stage[pos++] = 4 * q[u] + c;
                    removed += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
                }
            }
        }
        if (lane == 0) a.counts[chunk] = removed;
    }
    if (bad_any) atomicOr(a.status, ESPIC_ST_BAD_LEFT_INDEX);

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

struct FinishArgs {
    unsigned long long* acc;
    float* density;
    int* counts;     // in: removed per chunk; out: exclusive offsets
    int* top;        // device stack top
    int* scatter_hdr;  // [0] = old top, [1] = total removed, [2] = n_chunks
    int* status;
    const int* largest_index_dev;
    int largest_index;
    int largest_allowed_slot;
    int n_points;
    double inv_scale;  // 2^-24 / background_density
    int* flags_global;  // reset for the next large-grid launch
};

// One CTA: fixed-point accumulators -> float density (and re-zero them), exclusive scan of
// the per-chunk removal counts, stack-top update.
__global__ void __launch_bounds__(1024) k_finish(const FinishArgs a) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x;
    for (int j = tid; j < a.n_points; j += blockDim.x) {
        const unsigned long long v = a.acc[j];
        a.density[j] = (float)((double)v * a.inv_scale);
        a.acc[j] = 0ull;
    }
    if (tid == 0) {
        s_carry = 0;
        if (a.flags_global) a.flags_global[0] = 0;
    }
    __syncthreads();
    const int n_slots = (a.largest_index_dev ? *a.largest_index_dev : a.largest_index) + 1;
    const int n_quads = (n_slots + 3) >> 2;
    const int n_chunks = (n_quads + kQuadsPerChunk - 1) / kQuadsPerChunk;
    for (int base = 0; base < n_chunks; base += blockDim.x) {
        const int i = base + tid;
        const int c = (i < n_chunks) ? a.counts[i] : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += t;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            int w = s_warp[tid];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, d);
                if (tid >= d) w += t;
            }
            s_warp[tid] = w;  // inclusive over warps
        }
        __syncthreads();
        const int warp_off = (tid >> 5) ? s_warp[(tid >> 5) - 1] : 0;
        const int carry = s_carry;
        if (i < n_chunks) a.counts[i] = carry + warp_off + incl - c;
        __syncthreads();
        if (tid == blockDim.x - 1) s_carry = carry + warp_off + incl;
        __syncthreads();
    }
    if (tid == 0) {
        const int total = s_carry;
        const int old_top = *a.top;
        a.scatter_hdr[0] = old_top;
        a.scatter_hdr[1] = total;
        a.scatter_hdr[2] = n_chunks;
        *a.top = old_top + total;
        if (old_top + total > a.largest_allowed_slot) atomicOr(a.status, ESPIC_ST_STACK_OVERFLOW);
    }
}

// Copies each chunk's staged slot ids to their final position on the stack.
__global__ void __launch_bounds__(256) k_scatter(const int* __restrict__ staging,
                                                 const int* __restrict__ offsets,
                                                 const int* __restrict__ hdr, int* __restrict__ empty_slots,
                                                 int largest_allowed_slot) {
    const int total = hdr[1];
    if (total == 0) return;
    const int old_top = hdr[0], n_chunks = hdr[2];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int c = gw; c < n_chunks; c += nw) {
        const int off = offsets[c];
        const int end = (c + 1 < n_chunks) ? offsets[c + 1] : total;
        const int cnt = end - off;
        const int* src = staging + (size_t)c * kChunkSlots;
        for (int r = lane; r < cnt; r += 32) {
            const int dst = old_top + 1 + off + r;
            if (dst >= 0 && dst <= largest_allowed_slot) empty_slots[dst] = src[r];
        }
    }
}

__global__ void k_shift_velocities(float* __restrict__ v, int n, double v_b) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        v[i] = (float)((double)v[i] - v_b);  // ref infMagSim_cython.py:174-176
}

static size_t move_smem_bytes(int n_points) {
    return (size_t)((n_points + 3) & ~3) * kSmemBytesPerPoint;
}

template <bool PERIODIC, bool SMEM>
static int launch_move_variant(espic_ctx* ctx, const MoveArgs& a, size_t smem) {
    auto kern = k_move<PERIODIC, SMEM>;
    // occupancy of this variant for this tile size, queried once (one device per process)
    static size_t cached_smem = (size_t)-1;
    static int cached_per_sm = 0;
    if (cached_smem != smem) {
        if (smem > 48 * 1024)
            ESPIC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        ESPIC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMoveThreads, smem));
        if (per_sm < 1) return fail(ctx, ESPIC_E_ARG, "mover kernel does not fit on an SM (smem %zu)", smem);
        cached_smem = smem;
        cached_per_sm = per_sm;
    }
    const int grid = ctx->sm_count * cached_per_sm;
    ctx->mover_grid = grid;
    ctx->mover_block = kMoveThreads;
    const bool timed = ctx->timing_on && ctx->ev_used + 2 <= ctx->ev_cap;
    if (timed) cudaEventRecord(ctx->ev[ctx->ev_used], ctx->stream);
    kern<<<grid, kMoveThreads, smem, ctx->stream>>>(a);
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
                int periodic, int largest_allowed_slot, int n_points, int storage) {
    if (n_points < 2) return fail(ctx, ESPIC_E_ARG, "n_points must be >= 2");
    if (storage < 1) return fail(ctx, ESPIC_E_ARG, "particle_storage_length must be >= 1");
    if (!largest_index_dev && largest_index >= storage)
        return fail(ctx, ESPIC_E_ARG, "largest_index %d >= storage %d", largest_index, storage);
    int rc = ensure_acc(ctx, n_points);
    if (rc) return rc;
    const int max_slots = largest_index_dev ? storage : (largest_index + 1);
    const int max_quads = (max_slots + 3) >> 2;
    const int max_chunks = (max_quads + kQuadsPerChunk - 1) / kQuadsPerChunk;
    rc = ensure(ctx, ctx->staging, (size_t)(max_chunks > 0 ? max_chunks : 1) * kChunkSlots * sizeof(int), "staging");
    if (rc) return rc;
    rc = ensure(ctx, ctx->chunk_counts, ((size_t)max_chunks + 8) * sizeof(int), "chunk counts");
    if (rc) return rc;

    MoveArgs a{};
    a.grid = grid;
    a.mask = mask;
    a.phi = phi;
    a.px = particles;
    a.pv = particles + storage;
    a.staging = (int*)ctx->staging.ptr;
    a.counts = (int*)ctx->chunk_counts.ptr + 4;  // first 4 ints: scatter header
    a.acc = ctx->acc;
    a.status = ctx->status;
    a.largest_index_dev = largest_index_dev;
    a.largest_index = largest_index;
    a.n_points = n_points;
    a.dt = dt;
    a.qm = qm;
    a.a_b = a_b;
    a.update_position = update_position;
    a.vec_ok = ((reinterpret_cast<uintptr_t>(particles) & 15u) == 0 && (storage & 3) == 0) ? 1 : 0;

    const size_t smem = move_smem_bytes(n_points);
    int max_smem = 0;
    ESPIC_CUDA(ctx, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    const bool use_smem = smem <= (size_t)max_smem;
    int* flags = nullptr;
    if (!use_smem) {
        rc = ensure(ctx, ctx->dv_table, (size_t)n_points * sizeof(float) + 64, "kick table");
        if (rc) return rc;
        flags = (int*)ctx->dv_table.ptr;  // first 16 ints are flags (zeroed at allocation / by k_finish)
        float* dvt = (float*)ctx->dv_table.ptr + 16;
        k_build_dv<<<(n_points + 255) / 256, 256, 0, ctx->stream>>>(grid, mask, phi, n_points, qm, a_b, dt, dvt, flags);
        ctx->launches++;
        a.dv_global = dvt;
        a.flags_global = flags;
    }
    if (periodic) {
        rc = use_smem ? launch_move_variant<true, true>(ctx, a, smem) : launch_move_variant<true, false>(ctx, a, 0);
    } else {
        rc = use_smem ? launch_move_variant<false, true>(ctx, a, smem) : launch_move_variant<false, false>(ctx, a, 0);
    }
    if (rc) return rc;

    FinishArgs f{};
    f.acc = ctx->acc;
    f.density = density;
    f.counts = a.counts;
    f.top = top;
    f.scatter_hdr = (int*)ctx->chunk_counts.ptr;
    f.status = ctx->status;
    f.largest_index_dev = largest_index_dev;
    f.largest_index = largest_index;
    f.largest_allowed_slot = largest_allowed_slot;
    f.n_points = n_points;
    f.inv_scale = 1.0 / ((double)kWeightOne * (double)bg);
    f.flags_global = flags;
    k_finish<<<1, 1024, 0, ctx->stream>>>(f);
    ctx->launches++;
    ESPIC_CUDA(ctx, cudaGetLastError());
    const int sgrid = max_chunks < 8 * ctx->sm_count ? (max_chunks + 7) / 8 : ctx->sm_count;
    k_scatter<<<sgrid > 0 ? sgrid : 1, 256, 0, ctx->stream>>>(a.staging, a.counts, f.scatter_hdr, empty_slots,
                                                            largest_allowed_slot);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "mover epilogue launch");
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
