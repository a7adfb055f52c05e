// K6: stable cell sort of a particle store.  Nothing in the reference sorts (its store is
// visited in slot order, infMagSim_c.c:142); this is the B200-side reordering the north star
// asks for: live particles are ordered by their cell index -- the same (int)((x-z_min)/dz) the
// mover computes on entry -- ties keep their slot order (stable), dead slots are dropped, the
// survivors are compacted to slots [0, n_live) and the free-slot stack is rebuilt the way the
// driver builds it (free slots listed in reverse so the lowest is popped first,
// ref infMagSim_script.py:486-489).  x, v and the injection tag move together.  The
// permutation equals numpy's stable argsort of the same keys bit for bit
// (tests/test_gpu_sort.py).
//
// Two LSD radix passes over (key, x, v, tag) with digits of ceil(bits/2) <= 9 bits.  A tile is
// 8 warps x 32 rows x 32 lanes = 8192 consecutive items; within a tile the order is (warp,
// row, lane).  Per pass: k_sort_hist counts digits per tile; k_sort_offsets turns the
// [digit][tile] counts into global output offsets (digit-major exclusive scan); k_sort_scatter
// ranks every item inside its tile with warp match + per-warp running counters (no atomics,
// deterministic) and moves the payload.  Items of one tile with equal digits land on
// consecutive addresses (runs of ~tile/bins items), which L2 write-combines.
#include "espic_internal.cuh"

namespace espic {

constexpr int kSortWarps = 8;
constexpr int kSortRows = 32;
constexpr int kSortTile = kSortWarps * kSortRows * 32;  // 8192
constexpr int kSortMaxBins = 512;

struct SortArrays {
    const int* key;
    const float* x;
    const float* v;
    const int* tag;
};
struct SortOut {
    int* key;
    float* x;
    float* v;
    int* tag;
};

// key of every slot: cell index of a live particle, n_cells for a dead slot (sorts last)
template <bool PERIODIC>
__global__ void k_sort_keys(const float* __restrict__ grid, int n_points, const float* __restrict__ px, int n_slots,
                            int* __restrict__ key, int* __restrict__ n_live) {
    const Geom g = make_geom(grid, n_points);
    int live = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        const float x = px[i];
        int k = n_points - 1;  // dead
        if (x < g.live_thr) {  // not parked (ref infMagSim_c.c:146 / :211)
            bool bad = false;
            k = cell_of<PERIODIC, false>(g, PERIODIC ? fold_periodic(g, x) : x, bad);
            ++live;
        }
        key[i] = k;
    }
    live = __reduce_add_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(n_live, live);
}

__global__ void __launch_bounds__(256) k_sort_hist(const int* __restrict__ key, int n, int shift, int bins, int n_tiles,
                                                   int* __restrict__ hist /* [bins][n_tiles] */) {
    __shared__ int s_cnt[kSortMaxBins];
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = threadIdx.x; b < bins; b += blockDim.x) s_cnt[b] = 0;
        __syncthreads();
        const int base = tile * kSortTile;
        for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) {
            const int idx = base + i;
            if (idx < n) atomicAdd(&s_cnt[(key[idx] >> shift) & (bins - 1)], 1);
        }
        __syncthreads();
        for (int b = threadIdx.x; b < bins; b += blockDim.x) hist[(size_t)b * n_tiles + tile] = s_cnt[b];
        __syncthreads();
    }
}

// Exclusive scan of the flattened [bins][n_tiles] table, in place.  One CTA per digit scans its
// row after a first kernel has produced the row totals; rows are short (n/8192 entries).
__global__ void __launch_bounds__(256) k_sort_row_totals(const int* __restrict__ hist, int n_tiles,
                                                         int* __restrict__ totals) {
    __shared__ int s_red[8];
    const int d = blockIdx.x;
    int sum = 0;
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) sum += hist[(size_t)d * n_tiles + t];
    sum = __reduce_add_sync(0xffffffffu, sum);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += s_red[w];
        totals[d] = tot;
    }
}

__global__ void __launch_bounds__(256) k_sort_offsets(int* __restrict__ hist, int n_tiles, int bins,
                                                      const int* __restrict__ totals) {
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    const int d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) {
        int base = 0;
        for (int b = 0; b < d; ++b) base += totals[b];  // bins <= 512: trivial
        s_carry = base;
    }
    __syncthreads();
    int* row = hist + (size_t)d * n_tiles;
    for (int t0 = 0; t0 < n_tiles; t0 += blockDim.x) {
        const int t = t0 + tid;
        const int c = (t < n_tiles) ? row[t] : 0;
        int incl = c;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += o;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        int before = 0;
        for (int ww = 0; ww < w; ++ww) before += s_warp[ww];
        const int carry = s_carry;
        if (t < n_tiles) row[t] = carry + before + incl - c;
        __syncthreads();
        if (tid == blockDim.x - 1) s_carry = carry + before + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortWarps * 32) k_sort_scatter(const SortArrays in, const SortOut out, int n,
                                                                  int shift, int bins, int n_tiles,
                                                                  const int* __restrict__ offsets /* [bins][n_tiles] */) {
    __shared__ int s_cnt[kSortWarps][kSortMaxBins];  // per-warp digit counters -> running output cursors
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = threadIdx.x; b < kSortWarps * bins; b += blockDim.x) s_cnt[b / bins][b % bins] = 0;
        __syncthreads();
        const int wbase = tile * kSortTile + w * (kSortRows * 32);
        // sweep 1: digits of this warp's 1024 items, row by row (rows are sequential, so the
        // warp-private counters need no atomics)
        for (int r = 0; r < kSortRows; ++r) {
            const int idx = wbase + r * 32 + lane;
            const bool ok = idx < n;
            const int d = ok ? ((in.key[idx] >> shift) & (bins - 1)) : -1;
            const unsigned same = __match_any_sync(0xffffffffu, d);
            if (ok && (same & lt) == 0u) s_cnt[w][d] += __popc(same);  // the lowest lane of each group
            __syncwarp();
        }
        __syncthreads();
        // warp cursors: offset of (digit, tile) + counts of the lower warps
        for (int b = threadIdx.x; b < bins; b += blockDim.x) {
            int run = offsets[(size_t)b * n_tiles + tile];
            for (int ww = 0; ww < kSortWarps; ++ww) {
                const int c = s_cnt[ww][b];
                s_cnt[ww][b] = run;
                run += c;
            }
        }
        __syncthreads();
        // sweep 2: same order again; rank inside the row from the match mask, then advance the cursor
        for (int r = 0; r < kSortRows; ++r) {
            const int idx = wbase + r * 32 + lane;
            const bool ok = idx < n;
            int k = 0;
            if (ok) k = in.key[idx];
            const int d = ok ? ((k >> shift) & (bins - 1)) : -1;
            const unsigned same = __match_any_sync(0xffffffffu, d);
            if (ok) {
                const int dst = s_cnt[w][d] + __popc(same & lt);
                out.key[dst] = k;
                out.x[dst] = in.x[idx];
                out.v[dst] = in.v[idx];
                out.tag[dst] = in.tag[idx];
            }
            __syncwarp();
            if (ok && (same & lt) == 0u) s_cnt[w][d] += __popc(same);
            __syncwarp();
        }
        __syncthreads();
    }
}

// Parks everything behind the live particles and rebuilds the free-slot stack.
__global__ void k_sort_finalize(const float* __restrict__ grid, int n_points, float* __restrict__ px,
                                float* __restrict__ pv, int* __restrict__ tag, int storage, int n_slots,
                                const int* __restrict__ n_live_ptr, int* __restrict__ empty_slots, int n_stack,
                                int* __restrict__ top, int* __restrict__ largest) {
    const Geom g = make_geom(grid, n_points);
    const int n_live = *n_live_ptr;
    const int n_free = storage - n_live;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(storage, n_stack); i += gridDim.x * blockDim.x) {
        if (i >= n_live && i < n_slots) {  // dead entries moved behind the live ones
            px[i] = g.parked;
            pv[i] = 0.f;
            tag[i] = 0;
        }
        if (i < n_stack) empty_slots[i] = (i < n_free) ? storage - 1 - i : -1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *top = n_free - 1;
        *largest = n_live - 1;
    }
}

int launch_sort(espic_ctx* ctx, const float* grid, int n_points, float* particles, int storage, int* tags,
                int* empty_slots, int n_stack, int* top, int* largest, int largest_index, int periodic) {
    const int n = largest_index + 1;
    if (n_points < 2 || storage < 1 || n > storage) return fail(ctx, ESPIC_E_ARG, "espic_sort_particles: bad sizes");
    int bits = 1;
    while ((1 << bits) < n_points) ++bits;  // keys are 0 .. n_points-1 (n_points-1 = dead)
    if (bits > 18) return fail(ctx, ESPIC_E_ARG, "espic_sort_particles: more than 2^18 grid points");
    const int digit_bits = (bits + 1) / 2;
    const int bins = 1 << digit_bits;
    const int n_tiles = (n + kSortTile - 1) / kSortTile;
    // scratch: key0, key1, x1, v1, tag1 (n each), hist [bins][n_tiles], totals [bins], n_live
    const size_t n_pad = ((size_t)(n > 0 ? n : 1) + 63) & ~(size_t)63;
    const size_t hist_ints = (size_t)bins * (n_tiles > 0 ? n_tiles : 1);
    const size_t bytes = 5 * n_pad * 4 + (hist_ints + kSortMaxBins + 64) * 4;
    int rc = ensure(ctx, ctx->sort_ws, bytes, "sort workspace");
    if (rc) return rc;
    int* key0 = (int*)ctx->sort_ws.ptr;
    int* key1 = key0 + n_pad;
    float* x1 = (float*)(key1 + n_pad);
    float* v1 = x1 + n_pad;
    int* tag1 = (int*)(v1 + n_pad);
    int* hist = tag1 + n_pad;
    int* totals = hist + hist_ints;
    int* n_live = totals + kSortMaxBins;
    float* px = particles;
    float* pv = particles + storage;
    cudaStream_t st = ctx->stream;
    ESPIC_CUDA(ctx, cudaMemsetAsync(n_live, 0, sizeof(int), st));
    if (n > 0) {
        const int kg = (n + 255) / 256 < 8 * ctx->sm_count ? (n + 255) / 256 : 8 * ctx->sm_count;
        if (periodic) k_sort_keys<true><<<kg, 256, 0, st>>>(grid, n_points, px, n, key0, n_live);
        else k_sort_keys<false><<<kg, 256, 0, st>>>(grid, n_points, px, n, key0, n_live);
        ctx->launches++;
        const int tg = n_tiles < 4 * ctx->sm_count ? n_tiles : 4 * ctx->sm_count;
        for (int pass = 0; pass < 2; ++pass) {
            const int shift = pass * digit_bits;
            SortArrays in = pass == 0 ? SortArrays{key0, px, pv, tags} : SortArrays{key1, x1, v1, tag1};
            SortOut out = pass == 0 ? SortOut{key1, x1, v1, tag1} : SortOut{key0, px, pv, tags};
            k_sort_hist<<<tg, 256, 0, st>>>(in.key, n, shift, bins, n_tiles, hist);
            k_sort_row_totals<<<bins, 256, 0, st>>>(hist, n_tiles, totals);
            k_sort_offsets<<<bins, 256, 0, st>>>(hist, n_tiles, bins, totals);
            k_sort_scatter<<<tg, kSortWarps * 32, 0, st>>>(in, out, n, shift, bins, n_tiles, hist);
            ctx->launches += 4;
        }
    }
    const int m = storage > n_stack ? storage : n_stack;
    const int fg = (m + 255) / 256 < 8 * ctx->sm_count ? (m + 255) / 256 : 8 * ctx->sm_count;
    k_sort_finalize<<<fg > 0 ? fg : 1, 256, 0, st>>>(grid, n_points, px, pv, tags, storage, n, n_live, empty_slots,
                                                    n_stack, top, largest);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "sort launch");
}

}  // namespace espic
