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
// One or two LSD radix passes over (x, v, tag) with digits of <= 9 bits; the key is recomputed
// from x where it is needed instead of being carried as a fourth array.  A tile is 8 warps x 32
// rows x 32 lanes = 8192 consecutive items; within a tile the order is (warp, row, lane).
// Per pass: k_sort_hist counts digits per tile; k_sort_offsets turns the [digit][tile] counts
// into global output offsets (digit-major exclusive scan); k_sort_scatter ranks every item inside
// its tile with warp votes + per-warp running counters (no atomics, deterministic), permutes the
// tile in shared memory and writes it out digit run by digit run, so the global stores of a run
// are consecutive addresses (scattered 4-byte stores are bound by L2 request rate, not bytes).
// With key_shift > 0 the key is the cell index >> key_shift ("coarse sort": all the windowed
// mover needs, and one pass instead of two for the 65536-cell grid).
#include "espic_internal.cuh"

namespace espic {

constexpr int kSortWarps = 8;
constexpr int kSortRows = 32;
constexpr int kSortTile = kSortWarps * kSortRows * 32;  // 8192
constexpr int kSortMaxBins = 512;
constexpr unsigned kSortNone = 0xFFFFu;

// key of a slot: (cell index >> key_shift) of a live particle -- the same (int)((x-z_min)/dz) the
// mover computes on entry -- or dead_key for a parked slot (sorts last)
// With lookahead != 0 the cell is that of x + v*lookahead (clamped into the grid, or folded when
// periodic): sorting by where a particle will be half-way to the next sort keeps a segment of the
// store together for twice as many steps.
template <bool PERIODIC>
__device__ __forceinline__ int sort_key(const Geom& g, float x, float v, float lookahead, int key_shift,
                                        int dead_key) {
    if (!(x < g.live_thr)) return dead_key;  // ref infMagSim_c.c:146 / :211
    bool bad = false;
    const float xp = (lookahead != 0.f) ? __fadd_rn(x, __fmul_rn(v, lookahead)) : x;
    const float xf = PERIODIC ? fold_periodic(g, xp) : xp;
    // same truncated quotient either way (espic_selftest_quotient, positions inside the domain); the
    // 3-operation form keeps the sort memory-bound.  A look-ahead position beyond a wall is outside
    // the range that self-test covers: division instruction there.
    const bool fast = g.fast_div && (PERIODIC || lookahead == 0.f);
    const int cell = fast ? cell_of<PERIODIC, true>(g, xf, bad) : cell_of<PERIODIC, false>(g, xf, bad);
    return cell >> key_shift;
}

// Lanes of the warp holding the same digit as this lane, from one ballot per digit bit (cost
// proportional to the digit width; __match_any_sync costs one round per DISTINCT value, i.e. 32
// rounds on unsorted data).  Lanes without an item (kSortNone) form their own group.
__device__ __forceinline__ unsigned same_digit_lanes(unsigned d, int digit_bits) {
    const bool none = d == kSortNone;
    const unsigned bn = __ballot_sync(0xffffffffu, none);
    unsigned peers = none ? bn : ~bn;
    for (int bit = 0; bit < digit_bits; ++bit) {
        const bool one = (d >> bit) & 1u;
        const unsigned b = __ballot_sync(0xffffffffu, one);
        peers &= one ? b : ~b;
    }
    return peers;
}

struct SortKeying {
    const float* grid;
    int n_points, key_shift, dead_key, shift, bins;
    float lookahead;
};

template <bool PERIODIC>
__global__ void __launch_bounds__(256) k_sort_hist(const SortKeying kk, const float* __restrict__ px,
                                                   const float* __restrict__ pv, int n,
                                                   int n_tiles, int* __restrict__ hist /* [bins][n_tiles] */,
                                                   int* __restrict__ n_live /* may be null */) {
    __shared__ int s_cnt[kSortMaxBins];
    const Geom g = make_geom(kk.grid, kk.n_points);
    int live = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = threadIdx.x; b < kk.bins; b += blockDim.x) s_cnt[b] = 0;
        __syncthreads();
        const int base = tile * kSortTile;
        for (int i0 = 0; i0 < kSortTile; i0 += 16 * 256) {  // 16 loads in flight per thread
            float xr[16], vr[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int idx = base + i0 + k * 256 + (int)threadIdx.x;
                xr[k] = idx < n ? __ldcs(px + idx) : 0.f;
                vr[k] = (idx < n && kk.lookahead != 0.f) ? __ldcs(pv + idx) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int idx = base + i0 + k * 256 + (int)threadIdx.x;
                if (idx < n) {
                    const int key = sort_key<PERIODIC>(g, xr[k], vr[k], kk.lookahead, kk.key_shift, kk.dead_key);
                    live += key != kk.dead_key;
                    atomicAdd(&s_cnt[(key >> kk.shift) & (kk.bins - 1)], 1);
                }
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < kk.bins; b += blockDim.x) hist[(size_t)b * n_tiles + tile] = s_cnt[b];
        __syncthreads();
    }
    if (n_live) {
        live = __reduce_add_sync(0xffffffffu, live);
        if ((threadIdx.x & 31) == 0 && live) atomicAdd(n_live, live);
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

struct SortArrays {
    const float* x;
    const float* v;
    const int* tag;  // may be null
};
struct SortOut {
    float* x;
    float* v;
    int* tag;
};

constexpr size_t kSortScatterSmem = (size_t)kSortWarps * kSortMaxBins * 4   // per-warp digit counters / cursors
                                    + (size_t)kSortTile * 2                  // digit, then sorted position, per item
                                    + (size_t)kSortTile * 2                  // rank among the warp's items of that digit
                                    + (size_t)kSortTile * 2                  // digit per sorted position
                                    + (size_t)kSortTile * 4                  // one payload array, permuted
                                    + (size_t)kSortMaxBins * 4 + 64;         // output address minus sorted position, per digit

template <bool PERIODIC>
__global__ void __launch_bounds__(kSortWarps * 32) k_sort_scatter(const SortKeying kk, const SortArrays in,
                                                                  const SortOut out, int n, int n_tiles,
                                                                  const int* __restrict__ offsets /* [bins][n_tiles] */) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    int* s_cnt = reinterpret_cast<int*>(sort_smem);                              // [warp][bin]
    unsigned short* s_pos = reinterpret_cast<unsigned short*>(s_cnt + kSortWarps * kSortMaxBins);  // [item]
    unsigned short* s_rank = s_pos + kSortTile;                                  // [item]
    unsigned short* s_bin = s_rank + kSortTile;                                  // [sorted position]
    unsigned* s_data = reinterpret_cast<unsigned*>(s_bin + kSortTile);           // [sorted position]
    int* s_delta = reinterpret_cast<int*>(s_data + kSortTile);                   // [bin]
    __shared__ int s_warp_sum[kSortWarps];

    const Geom g = make_geom(kk.grid, kk.n_points);
    const int bins = kk.bins;
    const int digit_bits = 31 - __clz(bins);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int* my_cnt = s_cnt + w * bins;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int base = tile * kSortTile;
        const int n_tile = min(kSortTile, n - base);
        for (int b = tid; b < kSortWarps * bins; b += blockDim.x) s_cnt[b] = 0;
        const int wbase = w * (kSortRows * 32);
        // this warp's 1024 positions, all loads in flight at once, kept for the permutation below
        float xr[kSortRows];
#pragma unroll
        for (int r = 0; r < kSortRows; ++r) {
            const int li = wbase + r * 32 + lane;
            xr[r] = (li < n_tile) ? __ldcs(in.x + base + li) : 0.f;
        }
        float vr[kSortRows];
#pragma unroll
        for (int r = 0; r < kSortRows; ++r) {
            const int li = wbase + r * 32 + lane;
            vr[r] = (li < n_tile && kk.lookahead != 0.f) ? __ldg(in.v + base + li) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kSortRows; ++r) {
            const int li = wbase + r * 32 + lane;
            unsigned d = kSortNone;
            if (li < n_tile) {
                const int key = sort_key<PERIODIC>(g, xr[r], vr[r], kk.lookahead, kk.key_shift, kk.dead_key);
                d = (unsigned)((key >> kk.shift) & (bins - 1));
            }
            s_pos[li] = (unsigned short)d;
        }
        __syncthreads();
        // ranking sweep: per-warp digit counts and every item's rank among the warp's items of
        // its digit, row by row (rows are sequential, so the warp-private counters need no atomics)
        for (int r = 0; r < kSortRows; ++r) {
            const int li = wbase + r * 32 + lane;
            const unsigned d = s_pos[li];
            const unsigned same = same_digit_lanes(d, digit_bits);
            int before = 0;
            if (d != kSortNone) before = my_cnt[d];
            __syncwarp();
            if (d != kSortNone && (same & lt) == 0u) my_cnt[d] = before + __popc(same);  // the lowest lane of each group
            __syncwarp();
            s_rank[li] = (unsigned short)(before + __popc(same & lt));
        }
        __syncthreads();
        // start of every digit's run inside the sorted tile (exclusive scan over the digits), the
        // warps' cursors inside the run, and where the run goes in the output
        {
            const int b0 = 2 * tid, b1 = b0 + 1;  // 256 threads x 2 digits = 512
            int t0 = 0, t1 = 0;
            for (int ww = 0; ww < kSortWarps; ++ww) {
                if (b0 < bins) t0 += s_cnt[ww * bins + b0];
                if (b1 < bins) t1 += s_cnt[ww * bins + b1];
            }
            int incl = t0 + t1;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += o;
            }
            if (lane == 31) s_warp_sum[w] = incl;
            __syncthreads();
            int before = 0;
            for (int ww = 0; ww < w; ++ww) before += s_warp_sum[ww];
            int run = before + incl - (t0 + t1);
            if (b0 < bins) {
                s_delta[b0] = offsets[(size_t)b0 * n_tiles + tile] - run;
                for (int ww = 0; ww < kSortWarps; ++ww) {
                    const int c = s_cnt[ww * bins + b0];
                    s_cnt[ww * bins + b0] = run;
                    run += c;
                }
            }
            if (b1 < bins) {
                s_delta[b1] = offsets[(size_t)b1 * n_tiles + tile] - run;
                for (int ww = 0; ww < kSortWarps; ++ww) {
                    const int c = s_cnt[ww * bins + b1];
                    s_cnt[ww * bins + b1] = run;
                    run += c;
                }
            }
        }
        __syncthreads();
        // sorted position of every item: start of the warp's share of the digit's run + rank
#pragma unroll 4
        for (int r = 0; r < kSortRows; ++r) {
            const int li = wbase + r * 32 + lane;
            const unsigned d = s_pos[li];
            if (d != kSortNone) {
                const unsigned pos = (unsigned)my_cnt[d] + s_rank[li];
                s_bin[pos] = (unsigned short)d;
                s_pos[li] = (unsigned short)pos;
            }
        }
        __syncthreads();
        // payload arrays, one at a time through shared memory: in at the sorted position, out in
        // sorted order (a digit's run is consecutive in the output)
#pragma unroll
        for (int r = 0; r < kSortRows; ++r) {
            const int li = wbase + r * 32 + lane;
            if (li < n_tile) s_data[s_pos[li]] = __float_as_uint(xr[r]);
        }
        __syncthreads();
#pragma unroll 8
        for (int j = tid; j < n_tile; j += kSortWarps * 32) __stcs(reinterpret_cast<unsigned*>(out.x) + s_delta[s_bin[j]] + j, s_data[j]);
        __syncthreads();
        for (int arr = 1; arr < 3; ++arr) {
            const unsigned* src = arr == 1 ? reinterpret_cast<const unsigned*>(in.v)
                                           : reinterpret_cast<const unsigned*>(in.tag);
            unsigned* dst = arr == 1 ? reinterpret_cast<unsigned*>(out.v) : reinterpret_cast<unsigned*>(out.tag);
            if (!src) continue;  // uniform (no tag array)
            unsigned val[kSortRows];
#pragma unroll
            for (int r = 0; r < kSortRows; ++r) {
                const int li = wbase + r * 32 + lane;
                val[r] = (li < n_tile) ? __ldcs(src + base + li) : 0u;
            }
#pragma unroll
            for (int r = 0; r < kSortRows; ++r) {
                const int li = wbase + r * 32 + lane;
                if (li < n_tile) s_data[s_pos[li]] = val[r];
            }
            __syncthreads();
#pragma unroll 8
            for (int j = tid; j < n_tile; j += kSortWarps * 32) __stcs(dst + s_delta[s_bin[j]] + j, s_data[j]);
            __syncthreads();
        }
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
            if (tag) tag[i] = 0;
        }
        if (i < n_stack) empty_slots[i] = (i < n_free) ? storage - 1 - i : -1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *top = n_free - 1;
        *largest = n_live - 1;
    }
}

int launch_sort(espic_ctx* ctx, const float* grid, int n_points, float* particles, int storage, int* tags,
                int* empty_slots, int n_stack, int* top, int* largest, int largest_index, int periodic,
                int key_shift, float lookahead) {
    const int n = largest_index + 1;
    if (n_points < 2 || storage < 1 || n > storage) return fail(ctx, ESPIC_E_ARG, "espic_sort_particles: bad sizes");
    if (key_shift < 0 || key_shift > 30) return fail(ctx, ESPIC_E_ARG, "espic_sort_particles: bad key_shift");
    const int dead_key = ((n_points - 2) >> key_shift) + 1;  // keys are 0 .. dead_key
    int bits = 1;
    while ((1 << bits) <= dead_key) ++bits;
    if (bits > 18) return fail(ctx, ESPIC_E_ARG, "espic_sort_particles: more than 2^18 distinct keys");
    const int passes = bits <= 9 ? 1 : 2;
    const int digit_bits = passes == 1 ? bits : (bits + 1) / 2;
    const int bins = 1 << digit_bits;
    const int n_tiles = (n + kSortTile - 1) / kSortTile;
    // scratch: x1, v1, tag1 (n each), hist [bins][n_tiles], totals [bins], n_live
    const size_t n_pad = ((size_t)(n > 0 ? n : 1) + 63) & ~(size_t)63;
    const size_t hist_ints = (size_t)bins * (n_tiles > 0 ? n_tiles : 1);
    const size_t bytes = 3 * n_pad * 4 + (hist_ints + kSortMaxBins + 64) * 4;
    int rc = ensure(ctx, ctx->sort_ws, bytes, "sort workspace");
    if (rc) return rc;
    float* x1 = (float*)ctx->sort_ws.ptr;
    float* v1 = x1 + n_pad;
    int* tag1 = (int*)(v1 + n_pad);
    int* hist = tag1 + n_pad;
    int* totals = hist + hist_ints;
    int* n_live = totals + kSortMaxBins;
    float* px = particles;
    float* pv = particles + storage;
    cudaStream_t st = ctx->stream;
    if (!ctx->sort_configured) {
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_sort_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kSortScatterSmem));
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_sort_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kSortScatterSmem));
        ctx->sort_configured = true;
    }
    ESPIC_CUDA(ctx, cudaMemsetAsync(n_live, 0, sizeof(int), st));
    if (n > 0) {
        const int hg = n_tiles < 8 * ctx->sm_count ? n_tiles : 8 * ctx->sm_count;
        const int sg = n_tiles < 2 * ctx->sm_count ? n_tiles : 2 * ctx->sm_count;
        for (int pass = 0; pass < passes; ++pass) {
            const SortKeying kk{grid, n_points, key_shift, dead_key, pass * digit_bits, bins, lookahead};
            const SortArrays in = pass == 0 ? SortArrays{px, pv, tags} : SortArrays{x1, v1, tags ? tag1 : nullptr};
            const SortOut out = pass == 0 ? SortOut{x1, v1, tag1} : SortOut{px, pv, tags};
            int* live_out = pass == 0 ? n_live : nullptr;
            if (periodic) k_sort_hist<true><<<hg, 256, 0, st>>>(kk, in.x, in.v, n, n_tiles, hist, live_out);
            else k_sort_hist<false><<<hg, 256, 0, st>>>(kk, in.x, in.v, n, n_tiles, hist, live_out);
            k_sort_row_totals<<<bins, 256, 0, st>>>(hist, n_tiles, totals);
            k_sort_offsets<<<bins, 256, 0, st>>>(hist, n_tiles, bins, totals);
            if (periodic) k_sort_scatter<true><<<sg, kSortWarps * 32, kSortScatterSmem, st>>>(kk, in, out, n, n_tiles, hist);
            else k_sort_scatter<false><<<sg, kSortWarps * 32, kSortScatterSmem, st>>>(kk, in, out, n, n_tiles, hist);
            ctx->launches += 4;
        }
        if (passes == 1) {  // the sorted arrays are in the scratch buffers
            ESPIC_CUDA(ctx, cudaMemcpyAsync(px, x1, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            ESPIC_CUDA(ctx, cudaMemcpyAsync(pv, v1, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            if (tags) ESPIC_CUDA(ctx, cudaMemcpyAsync(tags, tag1, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
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
