// K1w: the mover for grids whose tables do not fit in shared memory (more than ~19k points, e.g.
// the 65536-cell hole-tracking runs).  Same arithmetic as k_move (mover.cu), same results bit
// for bit; what changes is where the grid tables live.
//
// The store is cut into segments of 32 consecutive ranges (one per warp of a CTA).  CTAs take
// segments from a device counter.  For each segment the CTA
//   1. samples 1024 positions of the segment, histograms their cells into 64 buckets and picks
//      the window of kWinCells cells (start on a bucket boundary) that holds the most samples,
//      centre-weighted -- circular (cell index modulo n_cells, the two end NODES kept apart) in
//      the periodic kernel, so that a segment straddling the seam fits one window; clamped to the
//      grid in the wall kernel, where nothing crosses the ends;
//   2. builds the window's tables in shared memory: kick per cell (copied from the global kick
//      table k_build_dv made) and ONE u32 plane of the fixed-point deposit -- a word that wraps
//      (every 256 full-weight deposits on a node) credits 2^32 to the global accumulator directly,
//      which buys a window half as wide again as two planes would allow (8 instead of 12 B/cell);
//   3. runs the slot update on its ranges, one quad (4 slots) per lane and instruction;
//   4. flushes the plane into the global accumulators.
//
// Step 3 is straight-line code for every slot: table index = cell - w_lo, a slot that takes no part in
// the gather or the deposit (dead, leaving, stray) given a dummy word pair of its lane with zero weights
// instead of being branched around.  Stray slots -- an electron that has outrun its segment, a particle
// injected since the last sort -- are patched per lane behind one vote each for the gather (kick from the
// global table) and the deposit (64-bit global reductions); the warp keeps going, nothing is re-done.
// ncu on the first version, which mapped, clamped and tested every slot on its own: 109 instructions per
// particle; an all-or-nothing variant (a warp with one stray takes an out-of-line general path) lost more
// than it gained: one stray in 128 slots is the rule late in a sort interval.
//
// Tried and dropped (A/B builds on the production-like 65536-cell run, 1e8 + 1e8, ms per step against 1.13):
// kick table left in global memory (L1/L2 gathers) with the whole shared memory as a 55296-cell deposit plane
// -- no stray for 58 steps, but 1.80 (28-step sort interval) to 1.89 (58): the gather sits in the dependency
// chain of every slot; a no-wrap instantiation of the mapping next to the general one (1.43: the duplicated
// loop costs more in instruction fetch than it saves); an all-or-nothing slow path per warp (one stray in 128
// slots is the rule late in a sort interval); 1024-thread CTAs at 64 registers (1.15); the bulk-copy (TMA) staging
// ring of k_move<periodic> -- its 64 KB have to come out of the window (20480 cells, sort every 21 steps instead of
// 28) and the CTA has to be 512 threads wide: 1.248 (periodic grid: mover 0.449 ms against 0.363); narrower CTAs on
// register loads (768 -> 1.133, 640 -> 1.172, 512 -> 1.241); L2 prefetch 2 / 3 / 4 units ahead (1.132 / 1.145 / 1.161);
// register loads issued one unit ahead (against 1.030: 896 threads 1.114, 768 -> 1.069, 640 -> 1.079 -- spills or too few warps).
//
// After espic_sort_particles_coarse() a segment's particles sit within a few hundred cells of each
// other and drift apart by |v| dt per step; the window leaves ~13700 cells of margin either side,
// so nearly all slots stay on the shared-memory path for many steps.  Because the deposit accumulates
// integers the density does not depend on which path a slot took, nor on the window, the segment
// schedule or the particle order.
#include <cstdio>
#include <cstdlib>

#include "mover_common.cuh"

namespace espic {

constexpr int kWinCells = 27648;
constexpr int kWinNodes = kWinCells + 2;  // both nodes of every cell, plus one when the window wraps past the seam
constexpr int kWinBuckets = 64;
#ifndef ESPIC_WIN_BOUNDS
#define ESPIC_WIN_BOUNDS 896    // CTA width of k_move_win: 28 warps at 72 registers (production-like 65536-cell step, ms:
                                // 768 -> 1.142, 896 -> 1.138, 1024 -> 1.146)
#endif
constexpr int kWinThreads = ESPIC_WIN_BOUNDS;

// cell -> index into the window tables.  Cells at or above w_lo map to cell - w_lo; cells that
// wrapped (below w_lo) map one further so that node n_cells (right end) and node 0 (left end)
// stay distinct.  `inside` says whether the cell belongs to the window at all.
struct WinMap {
    int w_lo, n_cells;
    __device__ __forceinline__ int operator()(int cell, bool& inside) const {
        const int t = cell - w_lo;
        const int m = t >> 31;
        const int rel = t + (n_cells & m);
        inside = (unsigned)rel < (unsigned)kWinCells;
        return rel - m;
    }
    // the same without the test, for the fast path: an index >= kWinCells means "outside"
    template <bool WRAP>
    __device__ __forceinline__ int index(int cell) const {
        const int t = cell - w_lo;
        if (!WRAP) return t;  // negative (cell below the window) is huge as unsigned
        const int m = t >> 31;
        return t + (n_cells & m) - m;
    }
    // table index -> grid node (inverse of the above for the left node of a cell)
    __device__ __forceinline__ int node_of(int j) const {
        const int seam = n_cells - w_lo;  // index of node n_cells when the window reaches it
        return j <= seam ? w_lo + j : j - seam - 1;
    }
};

struct KickWindow {
    const float* s_dv;
    const float* dv_global;
    WinMap map;
    __device__ __forceinline__ float operator()(int c) const {
        bool inside;
        const int j = map(c, inside);
        if (inside) return s_dv[j];
        return __ldg(dv_global + c);
    }
};

// One quad (see the header).  Every slot runs the straight-line code with a SAFE table index: a slot that takes
// no part (dead on entry, leaving, stray) has a table index outside the window by construction -- its "cell" is
// that of a parked or out-of-domain position -- and is given the lane's own dummy word pair jd instead, with zero
// weights (one shared dummy address would serialise: a third of the slots of a wall segment leave every step).
// Live strays -- table index >= limit -- are patched per lane behind one vote each for the gather and the deposit:
// kick from the global table, deposit by 64-bit global reductions; a live stray whose CELL is not even inside the
// grid makes the quad `rare` (redone by quad_generic, which clamps and reports as the reference's printf does).
// limit: table indices below it are cells of the window.
template <bool PERIODIC>
__device__ __forceinline__ unsigned quad_win(const MoveArgs& a, const Geom& g, Quad& p, int q, int n_slots,
                                             const KickWindow& kick, unsigned* s_lo, int jd, unsigned limit) {
    constexpr bool WRAP = PERIODIC;  // wall windows are clamped to the grid: index = cell - w_lo
    const WinMap& map = kick.map;
    const unsigned last = (unsigned)g.last_cell;
    int j0[4], c0[4];
    unsigned out0 = 0, alive = 0;
    bool rare = false;
    float xe[4];
    // ---- entry: liveness, cell, table index (ref infMagSim_c.c:146-165)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float b0;
        bool ok = true;
        if (PERIODIC) {  // a live slot inside the domain; everything else (dead slots included) is `rare`
            const float a0 = __fsub_rn(p.x[c], g.z_lo);
            xe[c] = __fadd_rn(a0, g.z_lo);
            b0 = __fsub_rn(xe[c], g.z_lo);
            rare |= !(a0 >= 0.f);
            rare |= !(b0 < g.cell_span);
        } else {  // dead slots are common (removed particles wait for re-injection)
            ok = (p.x[c] > g.lo_edge) && (p.x[c] < g.hi_edge);
            alive |= ok ? (1u << c) : 0u;
            xe[c] = p.x[c];
            b0 = __fsub_rn(p.x[c], g.z_lo);
        }
        c0[c] = __float2int_rz(quotient_dz<true>(g, b0));
        const int j = map.index<WRAP>(c0[c]);
        const bool out = (unsigned)j >= limit;
        j0[c] = out ? jd : j;
        out0 |= (out && ok) ? (1u << c) : 0u;
    }
    // ---- gather (ref :166-168)
    float kv[4];
    if (!__any_sync(0xffffffffu, out0 != 0u)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) kv[c] = kick.s_dv[j0[c]];
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool out = (out0 >> c) & 1u;
            const bool bad = out && (unsigned)c0[c] > last;  // a live slot whose cell is not a cell of the grid
            rare |= bad;
            kv[c] = (out && !bad) ? __ldg(kick.dv_global + c0[c]) : kick.s_dv[j0[c]];
        }
    }
    // ---- push, boundaries, exit cell (ref :166-189, :211-217)
    float xd[4];
    int j1[4], c1[4];
    unsigned dep = 0, rem = 0, out1 = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float v1 = __fadd_rn(p.v[c], kv[c]);
        const float x1 = __fadd_rn(xe[c], __fmul_rn(v1, a.dt));
        bool deposits = true;
        if (PERIODIC) {
            const float a1 = __fsub_rn(x1, g.z_lo);
            rare |= !(a1 > g.fold_lo);
            rare |= !(a1 < g.fold_hi);
            const float off = __fsub_rn(a1, (a1 >= g.span) ? g.span : 0.f);
            xd[c] = __fadd_rn(off, (off >= 0.f) ? g.z_lo : g.z_hi);
            const float b2 = __fsub_rn(xd[c], g.z_lo);
            rare |= !(b2 < g.cell_span);
            c1[c] = __float2int_rz(quotient_dz<true>(g, b2));
            p.x[c] = xd[c];
            p.v[c] = v1;
        } else {
            const bool ok0 = (alive >> c) & 1u;
            const bool ok1 = (x1 > g.lo_edge) && (x1 < g.hi_edge);  // false for NaN: parked, as the reference parks it
            deposits = ok0 && ok1;
            const bool removed = ok0 ? !ok1 : (p.x[c] < g.live_thr);
            xd[c] = x1;
            c1[c] = __float2int_rz(quotient_dz<true>(g, __fsub_rn(x1, g.z_lo)));
            p.x[c] = removed ? g.parked : (ok0 ? x1 : p.x[c]);
            p.v[c] = ok0 ? v1 : p.v[c];
            dep |= deposits ? (1u << c) : 0u;
            rem |= removed ? (1u << c) : 0u;
        }
        const int j = map.index<WRAP>(c1[c]);
        const bool out = (unsigned)j >= limit;
        j1[c] = out ? jd : j;
        out1 |= (out && deposits) ? (1u << c) : 0u;
    }
    if (PERIODIC) dep = 0xFu;
    const bool strays = __any_sync(0xffffffffu, out1 != 0u);
    if (strays) {
#pragma unroll
        for (int c = 0; c < 4; ++c) rare |= ((out1 >> c) & 1u) && (unsigned)c1[c] > last;
    }
    if (__any_sync(0xffffffffu, rare))
        return quad_generic<PERIODIC, false>(a, q, n_slots, a.dv_global, false, nullptr, nullptr);
    store_quad(a, q, p);
    // ---- deposit (ref :219-225): eight atomics, one running maximum of the returned words as the carry test;
    //      slots that do not deposit here (dead, removed, stray) add zero to the lane's dummy words
    const unsigned here = dep & ~out1;
    unsigned old[8], top = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned w0, w1;
        node_weights_fixed(g, xd[c], c1[c], w0, w1);
        const bool on = (here >> c) & 1u;
        w0 = on ? w0 : 0u;
        w1 = on ? w1 : 0u;
        old[2 * c] = atomicAdd(s_lo + j1[c], w0);
        old[2 * c + 1] = atomicAdd(s_lo + j1[c] + 1, w1);
        top = max(top, max(old[2 * c], old[2 * c + 1]));
    }
    if (__any_sync(0xffffffffu, top >= 0xFF000000u)) {  // a word may have wrapped (old value within 2^24 of 2^32): look,
#pragma unroll                                         // and credit its 2^32 to the global accumulator of the node
        for (int c = 0; c < 4; ++c) {
            unsigned w0, w1;
            node_weights_fixed(g, xd[c], c1[c], w0, w1);
            const bool on = (here >> c) & 1u;
            if (on && old[2 * c] > ~w0) atomicAdd(a.acc + map.node_of(j1[c]), 1ull << 32);
            if (on && old[2 * c + 1] > ~w1) atomicAdd(a.acc + map.node_of(j1[c] + 1), 1ull << 32);
        }
    }
    if (strays) {  // slots whose deposit cell lies outside the window: straight into the global accumulators
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if ((out1 >> c) & 1u) {
                unsigned w0, w1;
                node_weights_fixed(g, xd[c], c1[c], w0, w1);
                atomicAdd(a.acc + c1[c], (unsigned long long)w0);
                atomicAdd(a.acc + c1[c] + 1, (unsigned long long)w1);
            }
        }
    }
    return rem;
}

template <bool PERIODIC>
__global__ void __launch_bounds__(kWinThreads, 1) k_move_win(const __grid_constant__ MoveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_dv = reinterpret_cast<float*>(smem_raw);
    unsigned* s_lo = reinterpret_cast<unsigned*>(s_dv + kWinNodes);
    __shared__ int s_bucket[kWinBuckets];
    __shared__ int s_score[kWinBuckets];
    __shared__ int s_seg, s_w_lo;

#if ESPIC_DEBUG_CYCLES
    const long long t_entry = a.cta_cycles ? clock64() : 0ll;
#endif
    const int n_points = a.n_points, n_cells = n_points - 1;
    const Geom g = make_geom(a.grid, n_points);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int n_ranges = a.n_segments * warps_per_cta;
    const int n_slots = (a.largest_index_dev ? *a.largest_index_dev : a.largest_index) + 1;
    const int n_quads = (n_slots + 3) >> 2;
    const Split split(n_slots, n_ranges);
    const bool has_obj = a.flags_global[0] != 0;
    const bool fast = !has_obj && g.fast_div;  // uniform; the launcher has checked the rest
    const int bucket_cells = (n_cells + kWinBuckets - 1) / kWinBuckets;
    const int window_buckets = kWinCells / bucket_cells;

    for (;;) {
        if (tid == 0) s_seg = atomicAdd(a.seg_counter, 1);
        if (tid < kWinBuckets) s_bucket[tid] = 0;
        __syncthreads();
        // segments are handed out from both ends of the store inwards: the two wall segments -- where a third of the
        // slots is removed and re-injected every step, several times the work of an interior segment -- start first
        const int ticket = s_seg;
        if (ticket >= a.n_segments) break;
        const int seg = (ticket & 1) ? a.n_segments - 1 - (ticket >> 1) : (ticket >> 1);
#if ESPIC_DEBUG_CYCLES
        const long long t_seg = a.cta_cycles ? clock64() : 0ll;
#endif
        const int r0 = seg * warps_per_cta;
        const int seg_u0 = split.first_unit(r0), seg_u1 = split.first_unit(r0 + warps_per_cta);

        // 1. where are this segment's particles?
        const long long seg_slots = (long long)(seg_u1 - seg_u0) * kUnitSlots;
        if (seg_slots == 0 && seg != a.n_segments - 1) {  // uniform: nothing here (a store much smaller than the grid)
            if (lane == 0) a.counts[r0 + warp] = 0;
            __syncthreads();
            continue;
        }
        if (seg_slots > 0) {
            const float x = __ldg(a.px + (size_t)seg_u0 * kUnitSlots + (size_t)((tid * seg_slots) / blockDim.x));
            const bool live = PERIODIC ? (x < g.live_thr) : (x > g.lo_edge && x < g.hi_edge);
            if (live) {
                bool bad = false;
                const int c = cell_of<PERIODIC, false>(g, PERIODIC ? fold_periodic(g, x) : x, bad);
                atomicAdd(&s_bucket[c / bucket_cells], 1);
            }
        }
        __syncthreads();
        if (tid < kWinBuckets) {
            int score = 0;
            for (int i = 0; i < window_buckets; ++i) {
                const int b = tid + i;  // periodic: circular; walls: buckets past the end hold nothing
                const int cnt = PERIODIC ? s_bucket[b & (kWinBuckets - 1)] : (b < kWinBuckets ? s_bucket[b] : 0);
                score += cnt * min(i + 1, window_buckets - i);
            }
            s_score[tid] = score;
        }
        __syncthreads();
        if (tid == 0) {
            int best = 0;
            for (int b = 1; b < kWinBuckets; ++b)
                if (s_score[b] > s_score[best]) best = b;
            int w_lo = best * bucket_cells;
            if (n_cells <= kWinCells) w_lo = 0;                              // a window as large as the grid starts at cell 0
            else if (!PERIODIC) w_lo = min(w_lo, n_cells - kWinCells);       // wall windows never reach past the last cell
            s_w_lo = w_lo;
        }
        __syncthreads();
        const WinMap map{s_w_lo, n_cells};

        // 2. the window's tables
        for (int j = tid; j < kWinNodes; j += blockDim.x) {
            const int node = map.node_of(j);
            s_dv[j] = (node >= 0 && node < n_cells) ? __ldg(a.dv_global + node) : 0.f;
            s_lo[j] = 0u;
        }
        __syncthreads();

        // 3. this warp's range
        {
            const KickWindow kick{s_dv, a.dv_global, map};
            const unsigned limit = PERIODIC ? (unsigned)kWinCells : (unsigned)min(kWinCells, n_cells - map.w_lo);
            // the lane's own dummy word pair in the middle of the table, for slots that take no part (zero weights)
            const int jd = min(kWinCells / 2, n_cells / 2) + 2 * lane;
            const int range = r0 + warp;
            int removed = 0;
            const int u0 = split.first_unit(range), u1 = u0 + split.unit_count(range);
            int* stage_ids = a.staging + (size_t)u0 * kUnitSlots;
            if (fast) {
#pragma unroll 1
                for (int u = u0; u < u1; ++u) {
                    const int q0 = u * kUnitQuads + lane, q1 = q0 + 32;
                    if (lane < 16 && u + a.prefetch_units < u1) {
                        const float* row = (lane & 8) ? a.pv : a.px;
                        const float* line = row + (size_t)(u + a.prefetch_units) * kUnitSlots + (lane & 7) * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
                    }
                    Quad p0, p1;
                    load_quad(a, q0, p0);
                    load_quad(a, q1, p1);
                    const unsigned rem0 = quad_win<PERIODIC>(a, g, p0, q0, n_slots, kick, s_lo, jd, limit);
                    const unsigned rem1 = quad_win<PERIODIC>(a, g, p1, q1, n_slots, kick, s_lo, jd, limit);
                    if (!PERIODIC || __any_sync(0xffffffffu, (rem0 | rem1) != 0u)) {
                        stage_removed(rem0, q0, lane, stage_ids, removed);
                        stage_removed(rem1, q1, lane, stage_ids, removed);
                    }
                }
            } else {
#pragma unroll 1
                for (int q = u0 * kUnitQuads + lane; q < u1 * kUnitQuads; q += 32) {
                    const unsigned r = quad_generic<PERIODIC, false>(a, q, n_slots, a.dv_global, has_obj, nullptr, nullptr);
                    stage_removed(r, q, lane, stage_ids, removed);
                }
            }
            if (lane == 0) {
                a.counts[range] = removed;
                if (removed) atomicAdd(a.total_removed, removed);
            }
            if (range == n_ranges - 1) {
                // the slots behind the last complete unit: generic, bounds-checked, all global
                int removed_tail = 0;
                int* stage = a.staging + (size_t)split.n_units * kUnitSlots;
#pragma unroll 1
                for (int qb = split.n_units * kUnitQuads; qb < n_quads; qb += 32) {
                    const int q = qb + lane;
                    unsigned r = 0;
                    if (q < n_quads)
                        r = quad_generic<PERIODIC, false>(a, q, n_slots, a.dv_global, has_obj, nullptr, nullptr);
                    stage_removed(r, q, lane, stage, removed_tail);
                }
                if (lane == 0) {
                    a.counts[n_ranges] = removed_tail;
                    if (removed_tail) atomicAdd(a.total_removed, removed_tail);
                }
            }
        }
        __syncthreads();

        // 4. flush
        for (int j = tid; j < kWinNodes; j += blockDim.x) {
            const unsigned lo = s_lo[j];
            if (lo) atomicAdd(a.acc + map.node_of(j), (unsigned long long)lo);
        }
        __syncthreads();
#if ESPIC_DEBUG_CYCLES
        // development aid (build with -DESPIC_DEBUG_CYCLES=1, run with ESPIC_DEBUG_CTA=1): cycles per segment, ions in
        // [512, 1024), electrons in [0, 512); per CTA from 1024 on (scripts/cta_cycles_probe.py)
        if (a.cta_cycles && tid == 0 && seg < 512) a.cta_cycles[(a.qm > 0.f ? 512 : 0) + seg] = clock64() - t_seg;
#endif
    }
#if ESPIC_DEBUG_CYCLES
    if (a.cta_cycles && tid == 0) a.cta_cycles[1024 + blockIdx.x] = clock64() - t_entry;
#endif
}

// Cells of the window for a grid whose whole-grid tables do not fit (the caller has checked that).  Grids of
// up to kWinCells cells (about 19k - 27.6k cells) are covered completely: every cell is inside the window of
// every segment, nothing is ever served from global memory and no sorting is needed.
int move_window_size(int n_points) { return n_points >= 2 ? kWinCells : 0; }

int launch_move_window(espic_ctx* ctx, MoveArgs& a, int periodic) {
    const size_t smem = (size_t)kWinNodes * 8;
    static int segs_per_cta = 0;  // process-wide tuning knob (an environment variable), not device state
    if (!segs_per_cta) {
        // measured (production-like 65536-cell step, ms): 1 -> 1.261, 2 -> 1.253, 4 -> 1.272, 8 -> 1.366: every segment
        // costs a window search, a table build, a flush and a barrier at which the CTA waits for its slowest warp
        // (~5 us); with one segment per CTA the split is static and the two wall segments set the pace.  Tried and
        // dropped (A/B builds, same run): table build / flush with all 28 loads of a thread in flight (the arrays go
        // to local memory at 64 registers: 1.38 vs 1.29 at 4 segments), a one-warp window search with aggregated
        // histogram atomics (no change), fetching the next ticket ahead (no change)
        segs_per_cta = 2;
        if (const char* e = getenv("ESPIC_WINDOW_SEGMENTS")) {  // segments per CTA
            const int s = atoi(e);
            if (s >= 1 && s <= 64) segs_per_cta = s;
        }
    }
    if (!ctx->win_configured) {  // per context: the opt-in belongs to this device's copy of the kernels
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_move_win<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_move_win<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->win_configured = true;
    }
    a.n_segments = ctx->sm_count * segs_per_cta;
    a.seg_counter = const_cast<int*>(a.flags_global) + 1;
    ctx->mover_grid = ctx->sm_count;
    ctx->mover_block = kWinThreads;
    ctx->mover_ranges = a.n_segments * (kWinThreads / 32);
    const bool timed = ctx->timing_on && ctx->ev_used + 2 <= ctx->ev_cap;
    if (timed) cudaEventRecord(ctx->ev[ctx->ev_used], ctx->stream);
    if (periodic) k_move_win<true><<<ctx->sm_count, kWinThreads, smem, ctx->stream>>>(a);
    else k_move_win<false><<<ctx->sm_count, kWinThreads, smem, ctx->stream>>>(a);
    if (timed) {
        cudaEventRecord(ctx->ev[ctx->ev_used + 1], ctx->stream);
        ctx->ev_used += 2;
    }
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_move_win launch");
}

}  // namespace espic
