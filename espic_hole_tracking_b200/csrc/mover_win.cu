// K1w: the mover for grids whose tables do not fit in shared memory (more than ~19k points, e.g.
// the 65536-cell hole-tracking runs).  Same arithmetic as k_move (mover.cu), same results bit
// for bit; what changes is where the grid tables live.
//
// The store is cut into segments of 32 consecutive ranges (one per warp of a CTA).  CTAs take
// segments from a device counter.  For each segment the CTA
//   1. samples 1024 positions of the segment, histograms their cells into 64 buckets and picks
//      the circular window of kWinCells cells (start on a bucket boundary) that holds the most
//      samples, centre-weighted;
//   2. builds the window's tables in shared memory: kick per cell (copied from the global kick
//      table k_build_dv made) and ONE u32 plane of the fixed-point deposit -- a word that wraps
//      (every 256 full-weight deposits on a node) credits 2^32 to the global accumulator directly,
//      which buys a window half as wide again as two planes would allow (8 instead of 12 B/cell);
//   3. runs the fast slot update on its ranges.  A slot whose gather cell or deposit cell is
//      outside the window is served per lane from global memory (kick table load, 64-bit
//      reductions into ctx->acc) -- the warp keeps going, nothing is re-done;
//   4. flushes the planes into ctx->acc.
// The window is circular (cell index modulo n_cells, with the two end NODES kept apart), so a
// segment of freshly injected particles -- half at each wall -- or a periodic segment straddling
// the seam still fits one window.
//
// After espic_sort_particles() a segment's particles sit within a few hundred cells of each
// other and drift apart by |v| dt per step; the window leaves ~13700 cells of margin either side,
// so nearly all slots stay on the shared-memory path for many steps.  On an unsorted store about
// kWinCells/n_cells of the slots do, the rest behave like the all-global kernel.  Because the
// deposit accumulates integers the density does not depend on which path a slot took, nor on the
// window, the segment schedule or the particle order.
#include <cstdio>
#include <cstdlib>

#include "mover_common.cuh"

#ifndef ESPIC_VAR_WINCARRY
#define ESPIC_VAR_WINCARRY 1   // development switch: 1 = running-maximum carry test, 0 = one compare per word
                               // (measured: production-like 65536-cell step 1.26 vs 1.31 ms)
#endif
#ifndef ESPIC_VAR_WINWRAP
#define ESPIC_VAR_WINWRAP 0    // development switch: 1 = no-wrap windows take the one-subtraction mapping in an instantiation
                               // of their own (measured: the duplicated loop costs more in instruction fetch than the
                               // shorter mapping saves -- production-like 65536-cell step 1.43 vs 1.31 ms)
#endif

namespace espic {

constexpr int kWinCells = 27648;
constexpr int kWinNodes = kWinCells + 2;  // both nodes of every cell, plus one when the window wraps past the seam
constexpr int kWinBuckets = 64;
constexpr int kWinThreads = 1024;

// cell -> index into the window tables.  Cells at or above w_lo map to cell - w_lo; cells that
// wrapped (below w_lo) map one further so that node n_cells (right end) and node 0 (left end)
// stay distinct.  `inside` says whether the cell belongs to the window at all.
// WRAP = false: the window ends before the seam (w_lo + kWinCells <= n_cells, true for every segment
// but the ones next to a wall or straddling the periodic seam) and the mapping is one subtraction.
template <bool WRAP>
struct WinMap {
    int w_lo, n_cells;
    __device__ __forceinline__ int operator()(int cell, bool& inside) const {
        const int t = cell - w_lo;
        if (!WRAP) {
            inside = (unsigned)t < (unsigned)kWinCells;
            return t;
        }
        const int m = t >> 31;
        const int rel = t + (n_cells & m);
        inside = (unsigned)rel < (unsigned)kWinCells;
        return rel - m;
    }
    // table index -> grid node (inverse of the above for the left node of a cell)
    __device__ __forceinline__ int node_of(int j) const {
        const int seam = n_cells - w_lo;  // index of node n_cells when the window reaches it
        return j <= seam ? w_lo + j : j - seam - 1;
    }
};

template <bool WRAP>
struct KickWindow {
    const float* s_dv;
    const float* dv_global;
    WinMap<WRAP> map;
    __device__ __forceinline__ float operator()(int c) const {
        bool inside;
        const int j = map(c, inside);
        if (inside) return s_dv[j];
        return __ldg(dv_global + c);
    }
};

// Deposit of the slots of a quad named by `dep`: window cells through the shared-memory plane, the others
// straight into the global accumulators.  A word of the plane can only wrap when its old value lies within
// 2^24 of 2^32, so one running maximum over the returned values decides whether any carry has to be looked
// for at all (see deposit_quad_all).
template <bool WRAP>
__device__ __forceinline__ void deposit_quad_win(const Geom& g, const float* x, const int* cell, unsigned dep,
                                                 const WinMap<WRAP>& map, unsigned* s_lo, unsigned long long* acc) {
#if ESPIC_VAR_WINCARRY
    unsigned old[8], top = 0u, outside = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned w0, w1;
        node_weights_fixed(g, x[c], cell[c], w0, w1);
        bool inside;
        const int j = map(cell[c], inside);
        const bool on = (dep >> c) & 1u;
        old[2 * c] = old[2 * c + 1] = 0u;
        if (on && inside) {
            old[2 * c] = atomicAdd(s_lo + j, w0);
            old[2 * c + 1] = atomicAdd(s_lo + j + 1, w1);
        }
        top = max(top, max(old[2 * c], old[2 * c + 1]));
        if (on && !inside) outside |= 1u << c;
    }
    if (__any_sync(0xffffffffu, (top >= 0xFF000000u) | (outside != 0u))) {  // rare: weights and indices are made again
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned w0, w1;
            node_weights_fixed(g, x[c], cell[c], w0, w1);
            bool inside;
            const int j = map(cell[c], inside);
            // the word wrapped: its 2^32 goes straight to the global accumulator of that node
            if (old[2 * c] > ~w0) atomicAdd(acc + map.node_of(j), 1ull << 32);
            if (old[2 * c + 1] > ~w1) atomicAdd(acc + map.node_of(j + 1), 1ull << 32);
            if (outside & (1u << c)) {
                atomicAdd(acc + cell[c], (unsigned long long)w0);
                atomicAdd(acc + cell[c] + 1, (unsigned long long)w1);
            }
        }
    }
#else
    unsigned carries = 0, outside = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned w0, w1;
        node_weights_fixed(g, x[c], cell[c], w0, w1);
        bool inside;
        const int j = map(cell[c], inside);
        const bool on = (dep >> c) & 1u;
        if (on && inside) {
            const unsigned old0 = atomicAdd(s_lo + j, w0);
            const unsigned old1 = atomicAdd(s_lo + j + 1, w1);
            if (old0 > ~w0) carries |= 1u << (2 * c);
            if (old1 > ~w1) carries |= 2u << (2 * c);
        }
        if (on && !inside) outside |= 1u << c;
    }
    if (__any_sync(0xffffffffu, (carries | outside) != 0u)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned w0, w1;
            node_weights_fixed(g, x[c], cell[c], w0, w1);
            bool inside;
            const int j = map(cell[c], inside);
            if (carries & (1u << (2 * c))) atomicAdd(acc + map.node_of(j), 1ull << 32);
            if (carries & (2u << (2 * c))) atomicAdd(acc + map.node_of(j + 1), 1ull << 32);
            if (outside & (1u << c)) {
                atomicAdd(acc + cell[c], (unsigned long long)w0);
                atomicAdd(acc + cell[c] + 1, (unsigned long long)w1);
            }
        }
    }
#endif
}

// One quad through the window tables.  The common case -- every live slot of the warp gathers from and
// deposits into the window -- is straight-line code: four table loads, eight shared-memory atomics, no
// per-lane branch (ncu on the first version, which mapped and tested every slot on its own: 109
// instructions per particle, 5 branches and 5.5 convergence barriers among them).  A warp that holds a slot
// whose gather or deposit cell lies outside the window takes the per-lane path for that part, behind one
// vote per quad.  Slots that do not take part (dead on entry, removed) use table index 0 with zero weights.
template <bool PERIODIC, bool WRAP>
__device__ __forceinline__ unsigned quad_fast_win(const MoveArgs& a, const Geom& g, Quad& p, int q, int n_slots,
                                                  const KickWindow<WRAP>& kick, unsigned* s_lo) {
    const unsigned last = (unsigned)g.last_cell;
    const WinMap<WRAP>& map = kick.map;
    bool rare = false;
    int c0[4], j0[4];
    unsigned alive = 0, in0 = 0;
    float x0[4];
    // ---- entry: liveness, cell, window index (ref infMagSim_c.c:146-165)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        bool ok;
        if (PERIODIC) {
            const float a0 = __fsub_rn(p.x[c], g.z_lo);
            x0[c] = __fadd_rn(a0, g.z_lo);
            const float b0 = __fsub_rn(x0[c], g.z_lo);
            rare |= !(a0 >= 0.f);
            rare |= !(b0 < g.cell_span);
            c0[c] = (int)min((unsigned)__float2int_rz(quotient_dz<true>(g, b0)), last);
            ok = true;
        } else {
            x0[c] = p.x[c];
            ok = (x0[c] > g.lo_edge) && (x0[c] < g.hi_edge);
            const int cr = __float2int_rz(quotient_dz<true>(g, __fsub_rn(x0[c], g.z_lo)));
            c0[c] = (int)min((unsigned)cr, last);
            rare |= ok && (c0[c] != cr);
        }
        alive |= ok ? (1u << c) : 0u;
        bool inside;
        const int j = map(c0[c], inside);
        j0[c] = (inside && ok) ? j : 0;
        in0 |= (inside || !ok) ? (1u << c) : 0u;
    }
    // ---- gather (ref :166-168): one table load per slot
    float kv[4];
    if (__all_sync(0xffffffffu, in0 == 0xFu)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) kv[c] = kick.s_dv[j0[c]];
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) kv[c] = (in0 & (1u << c)) ? kick.s_dv[j0[c]] : __ldg(kick.dv_global + c0[c]);
    }
    // ---- push, boundaries, exit cell (ref :166-189, :211-217)
    int c1[4], j1[4];
    unsigned dep = 0, rem = 0, in1 = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const bool ok0 = (alive >> c) & 1u;
        const float v1 = __fadd_rn(p.v[c], kv[c]);
        const float x1 = __fadd_rn(x0[c], __fmul_rn(v1, a.dt));
        bool deposits;
        if (PERIODIC) {
            const float a1 = __fsub_rn(x1, g.z_lo);
            rare |= !(a1 > g.fold_lo);
            rare |= !(a1 < g.fold_hi);
            const float off = __fsub_rn(a1, (a1 >= g.span) ? g.span : 0.f);
            const float x2 = __fadd_rn(off, (off >= 0.f) ? g.z_lo : g.z_hi);
            const float b2 = __fsub_rn(x2, g.z_lo);
            rare |= !(b2 < g.cell_span);
            c1[c] = (int)min((unsigned)__float2int_rz(quotient_dz<true>(g, b2)), last);
            p.x[c] = x2;
            p.v[c] = v1;
            deposits = true;
        } else {
            const bool ok1 = (x1 > g.lo_edge) && (x1 < g.hi_edge);
            const int cr = __float2int_rz(quotient_dz<true>(g, __fsub_rn(x1, g.z_lo)));
            c1[c] = (int)min((unsigned)cr, last);
            rare |= ok0 && ((ok1 && (c1[c] != cr)) || !(x1 == x1));
            const bool removed = ok0 ? !ok1 : (p.x[c] < g.live_thr);
            deposits = ok0 && ok1;
            p.x[c] = removed ? g.parked : (ok0 ? x1 : p.x[c]);
            p.v[c] = ok0 ? v1 : p.v[c];
            rem |= removed ? (1u << c) : 0u;
        }
        dep |= deposits ? (1u << c) : 0u;
        bool inside;
        const int j = map(c1[c], inside);
        j1[c] = (inside && deposits) ? j : 0;
        in1 |= (inside || !deposits) ? (1u << c) : 0u;
    }
    if (__any_sync(0xffffffffu, rare))
        return quad_generic<PERIODIC, false>(a, q, n_slots, a.dv_global, false, nullptr, nullptr);
    store_quad(a, q, p);
    // ---- deposit (ref :219-225)
    if (__all_sync(0xffffffffu, in1 == 0xFu)) {
        unsigned old[8], top = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned w0, w1;
            node_weights_fixed(g, p.x[c], c1[c], w0, w1);
            const bool on = (dep >> c) & 1u;
            w0 = on ? w0 : 0u;
            w1 = on ? w1 : 0u;
            old[2 * c] = atomicAdd(s_lo + j1[c], w0);
            old[2 * c + 1] = atomicAdd(s_lo + j1[c] + 1, w1);
            top = max(top, max(old[2 * c], old[2 * c + 1]));
        }
        if (__any_sync(0xffffffffu, top >= 0xFF000000u)) {  // a word of the plane may have wrapped: look, and
#pragma unroll                                             // credit its 2^32 to the global accumulator
            for (int c = 0; c < 4; ++c) {
                unsigned w0, w1;
                node_weights_fixed(g, p.x[c], c1[c], w0, w1);
                const bool on = (dep >> c) & 1u;
                if (on && old[2 * c] > ~w0) atomicAdd(a.acc + map.node_of(j1[c]), 1ull << 32);
                if (on && old[2 * c + 1] > ~w1) atomicAdd(a.acc + map.node_of(j1[c] + 1), 1ull << 32);
            }
        }
    } else {
        deposit_quad_win<WRAP>(g, p.x, c1, dep, map, s_lo, a.acc);
    }
    return rem;
}

// One warp's range of units of a segment through the window tables.
template <bool PERIODIC, bool WRAP>
__device__ __forceinline__ void win_range(const MoveArgs& a, const Geom& g, const Split& split, int range, int n_ranges,
                                          int n_slots, int n_quads, int lane, bool fast, bool has_obj,
                                          const float* s_dv, unsigned* s_lo, int w_lo, int n_cells) {
    const KickWindow<WRAP> kick{s_dv, a.dv_global, WinMap<WRAP>{w_lo, n_cells}};
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
            const unsigned rem0 = quad_fast_win<PERIODIC, WRAP>(a, g, p0, q0, n_slots, kick, s_lo);
            const unsigned rem1 = quad_fast_win<PERIODIC, WRAP>(a, g, p1, q1, n_slots, kick, s_lo);
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

template <bool PERIODIC>
__global__ void __launch_bounds__(kWinThreads, 1) k_move_win(const MoveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_dv = reinterpret_cast<float*>(smem_raw);
    unsigned* s_lo = reinterpret_cast<unsigned*>(s_dv + kWinNodes);
    __shared__ int s_bucket[kWinBuckets];
    __shared__ int s_score[kWinBuckets];
    __shared__ int s_seg, s_w_lo;

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
        const int seg = s_seg;
        if (seg >= a.n_segments) break;
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
            for (int i = 0; i < window_buckets; ++i)
                score += s_bucket[(tid + i) & (kWinBuckets - 1)] * min(i + 1, window_buckets - i);
            s_score[tid] = score;
        }
        __syncthreads();
        if (tid == 0) {
            int best = 0;
            for (int b = 1; b < kWinBuckets; ++b)
                if (s_score[b] > s_score[best]) best = b;
            s_w_lo = (n_cells <= kWinCells) ? 0 : best * bucket_cells;  // a window as large as the grid starts at cell 0
        }
        __syncthreads();
        const int w_lo = s_w_lo;
        const bool wraps = !ESPIC_VAR_WINWRAP || w_lo + kWinCells > n_cells;  // uniform
        const WinMap<true> map{w_lo, n_cells};          // the general mapping for table build and flush

        // 2. the window's tables
        for (int j = tid; j < kWinNodes; j += blockDim.x) {
            const int node = map.node_of(j);
            s_dv[j] = (node >= 0 && node < n_cells) ? __ldg(a.dv_global + node) : 0.f;
            s_lo[j] = 0u;
        }
        __syncthreads();

        // 3. this warp's range (the common no-wrap mapping and the general one are two instantiations;
        //    which one runs is uniform over the CTA)
        if (wraps) win_range<PERIODIC, true>(a, g, split, r0 + warp, n_ranges, n_slots, n_quads, lane, fast, has_obj, s_dv, s_lo, w_lo, n_cells);
        else win_range<PERIODIC, false>(a, g, split, r0 + warp, n_ranges, n_slots, n_quads, lane, fast, has_obj, s_dv, s_lo, w_lo, n_cells);
        __syncthreads();

        // 4. flush
        for (int j = tid; j < kWinNodes; j += blockDim.x) {
            const unsigned lo = s_lo[j];
            if (lo) atomicAdd(a.acc + map.node_of(j), (unsigned long long)lo);
        }
        __syncthreads();
    }
}

// Cells of the window for a grid whose whole-grid tables do not fit (the caller has checked that).  Grids of
// up to kWinCells cells (about 19k - 27.6k cells) are covered completely: every cell is inside the window of
// every segment, nothing is ever served from global memory and no sorting is needed.
int move_window_size(int n_points) { return n_points >= 2 ? kWinCells : 0; }

int launch_move_window(espic_ctx* ctx, MoveArgs& a, int periodic) {
    const size_t smem = (size_t)kWinNodes * 8;
    static int segs_per_cta = 0;  // process-wide tuning knob (an environment variable), not device state
    if (!segs_per_cta) {
        segs_per_cta = 2;  // measured (production-like 65536-cell step): 2 -> 1.231, 3 -> 1.250, 4 -> 1.280, 6 -> 1.346 ms
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
