// K5: grid Poisson solve as one thread-block cluster.  Replaces poisson_solve_c + tridiagonal_solve_c
// (ref infMagSim_c.c:343-354, 356-541) and, fused in front of it, the density end-node fix /
// charge density of the driver loop (ref infMagSim_script.py:763-779, 807).
//
// The rows are built per grid node exactly as the reference builds them (float/double mix
// included).  The Thomas algorithm's sequential recurrences are evaluated as block-wide scans
// in fp64:
//   D[j] = diag[j] - (sup[j-1]*sub[j-1])/D[j-1]        Moebius map  -> scan of 2x2 matrices
//   y[j] = rhs[j]  - y[j-1]*(sub[j-1]/D[j-1])          affine map   -> scan of (A,B) pairs
//   z[j] = y[j]    - z[j+1]*(sup[j]/D[j+1])            affine map, reversed
// The kernel runs as a cluster of C <= 8 CTAs of 1024 threads acting as one team of Tn = C*1024
// threads (C = 1 up to 1024 points, 8 from 3585 points on).  Each thread owns a contiguous run of
// m = ceil(n/Tn) rows: it composes its run, the team scans the Tn composites (warp shuffles, then
// the CTA's 32 warp totals, then the C CTA totals, which every CTA writes into every other CTA's
// shared memory -- DSMEM -- ahead of one cluster barrier), and the thread re-walks its run from
// the exact incoming value.  Row j = T*m + i of team thread T is stored at [i*Tn + T] (coalesced);
// run-boundary neighbours are read from the neighbouring CTA's shared memory.  At 4097 points
// m = 1 (rows live in registers / one shared-memory plane), at 65537 points m = 9 instead of the
// 65 a single CTA would walk through one SM's L2 port.
// The elimination (D and the multipliers sub/D, sup/D, 1/D) depends only on the matrix, which
// is constant from step to step; it is kept in the context's workspace under a key the kernel
// re-validates on the device at every call, so a driver step costs two affine scans.
// Rounding differs from the reference's sequential sweep at the 1e-16 level (multiplier
// precomputed instead of (y*sub)/D; scan re-association); the potential is float.
//
// Periodic variant: bug-compatible with the reference: the second solve
// re-eliminates the already eliminated diagonal and v.w is taken as 0 (the reference reads
// one element past its arrays there; with glibc the product underflows to +0.0).
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>

#include "espic_internal.cuh"

namespace cg = cooperative_groups;

namespace espic {

constexpr int kSolveThreads = 1024;
constexpr int kSolveSmemRows = 16;  // the y plane lives in shared memory up to this many rows per thread (128 KB)

struct Mat2 {
    double a, b, c, d;
};
struct Aff {
    double A, B;
};

__device__ __forceinline__ Mat2 mat_identity() { return Mat2{1., 0., 0., 1.}; }
__device__ __forceinline__ Aff aff_identity() { return Aff{1., 0.}; }

// 2^-e for the binary exponent e of |m| (m finite, non-zero), built from the exponent bits
__device__ __forceinline__ double inv_pow2_of(double m) {
    const int hi = __double2hiint(m);
    const int biased = (hi >> 20) & 0x7ff;
    int inv = 2046 - biased;          // exponent field of 2^-(biased-1023)
    inv = min(max(inv, 1), 2046);
    return __hiloint2double(inv << 20, 0);
}

// later * earlier, rescaled by a power of two (the map is projective) so long products neither
// overflow nor vanish
__device__ __forceinline__ Mat2 compose(const Mat2& l, const Mat2& e) {
    Mat2 r;
    r.a = l.a * e.a + l.b * e.c;
    r.b = l.a * e.b + l.b * e.d;
    r.c = l.c * e.a + l.d * e.c;
    r.d = l.c * e.b + l.d * e.d;
    const double m = fmax(fmax(fabs(r.a), fabs(r.b)), fmax(fabs(r.c), fabs(r.d)));
    if (m > 0. && m < 1.7e308) {
        const double s = inv_pow2_of(m);
        r.a *= s;
        r.b *= s;
        r.c *= s;
        r.d *= s;
    }
    return r;
}
__device__ __forceinline__ Aff compose(const Aff& l, const Aff& e) {
    return Aff{l.A * e.A, l.A * e.B + l.B};
}

__device__ __forceinline__ Mat2 shfl_up(const Mat2& v, int d) {
    return Mat2{__shfl_up_sync(0xffffffffu, v.a, d), __shfl_up_sync(0xffffffffu, v.b, d),
                __shfl_up_sync(0xffffffffu, v.c, d), __shfl_up_sync(0xffffffffu, v.d, d)};
}
__device__ __forceinline__ Aff shfl_up(const Aff& v, int d) {
    return Aff{__shfl_up_sync(0xffffffffu, v.A, d), __shfl_up_sync(0xffffffffu, v.B, d)};
}

// Exclusive scan over the block in thread order; element t of the result is the composition
// of the elements of threads 0..t-1 (earlier applied first).
template <class T>
__device__ __forceinline__ T block_exclusive_scan(T mine, T identity, T* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T other = shfl_up(incl, d);
        if (lane >= d) incl = compose(incl, other);
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        T x = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T other = shfl_up(x, d);
            if (lane >= d) x = compose(x, other);
        }
        s_warp[lane] = x;
    }
    __syncthreads();
    T before_warp = (w > 0) ? s_warp[w - 1] : identity;
    T excl = shfl_up(incl, 1);
    if (lane == 0) excl = identity;
    T res = compose(excl, before_warp);
    __syncthreads();
    return res;
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = (lane < (int)(blockDim.x >> 5)) ? s_red[lane] : 0.;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) s_red[0] = x;
    }
    __syncthreads();
    const double r = s_red[0];
    __syncthreads();
    return r;
}

struct SolveArgs {
    const float* grid;
    const float* mask;
    const float* charge;   // input charge density (when ion == nullptr)
    float* ion;            // fused prepare: ion / electron densities (fixed in place) -> charge_out
    float* ele;
    float* charge_out;
    float* potential;
    double* ws;   // header + 10 planes of m*1024 doubles (see below)
    int* status;
    float debye, obj_pot, transparency;
    int boltzmann, periodic, n;
    int fix_i, fix_e;
    int fold_ends;  // fused prologue: periodic_particles (fold the end nodes) vs walls (double them)
    // multi-GPU exchange: the densities are the rank-ordered sum of every rank's published slot
    int world;
    unsigned expect;                          // publication counter value to wait for
    const float* peer_slot[ESPIC_MAX_RANKS];  // [2][n_pad] floats: ion row, electron row
    const unsigned* peer_flag[ESPIC_MAX_RANKS];
    int n_pad;
    unsigned long long timeout_ns;            // how long to wait for a peer's publication
    // fused deposit: when set, the density of that species is still in the movers' fixed-point
    // accumulators; the prologue converts it (the one float rounding per node k_finish would have
    // made), stores it in ion / ele and clears the accumulators for the next pushes
    unsigned long long* acc_i;
    unsigned long long* acc_e;
    double inv_scale;                         // 2^-24 / background_density
    StackFix fix;                             // see espic_internal.cuh; pushed[s] == nullptr: nothing to look at
};

// Practically never executed: the periodic movers of the previous step removed slots and pushed them in
// arrival order.  The new stretch of each stack is put into ascending slot order (what the reference's
// slot-ordered loop produces, ref infMagSim_c.c:214-217).  A handful of entries -- the normal case when the
// stretch exists at all -- is an insertion sort by one thread (fix_stack_small).  A long stretch (a run that
// blew up numerically can throw every particle out of a periodic box in one step) is sorted by the whole
// team, fix_stack_team below: quadratic work by one thread would look like a hung GPU.
constexpr int kFixSmall = 48;

__device__ __noinline__ void fix_stack_small(int* st, int lo, int hi) {
    for (int i = lo + 1; i <= hi; ++i) {
        const int v = st[i];
        int j = i - 1;
        while (j >= lo && st[j] > v) {
            st[j + 1] = st[j];
            --j;
        }
        st[j + 1] = v;
    }
}

#ifndef ESPIC_SOLVE_MAXTEAM
#define ESPIC_SOLVE_MAXTEAM 16  // CTAs per cluster at most.  8 is the portable size; 16 (non-portable opt-in, taken from 8193 points
                                // on; the launcher falls back when the device cannot co-schedule it) makes the 65537-point solve
                                // 9 us shorter: 65536-cell production step 1.030 -> 1.021 ms
#endif
constexpr int kMaxTeam = ESPIC_SOLVE_MAXTEAM;

// The CTAs of the cluster acting as one block of Tn threads.
struct Team {
    int rank, size;  // CTA rank in the cluster, CTAs in the cluster
    int T, Tn;       // team thread index, team thread count
};
__device__ __forceinline__ Team make_team() {
    cg::cluster_group c = cg::this_cluster();
    Team tm;
    tm.rank = (int)c.block_rank();
    tm.size = (int)c.num_blocks();
    tm.T = tm.rank * kSolveThreads + (int)threadIdx.x;
    tm.Tn = tm.size * kSolveThreads;
    return tm;
}
__device__ __forceinline__ void team_sync(const Team& tm) {
    if (tm.size > 1) cg::this_cluster().sync();
    else __syncthreads();
}
// Bitonic network with every compare-exchange ascending (the first sub-stage of a merge pairs i with its mirror
// image in the block), so the stretch can be padded to a power of two virtually: a position past the end holds
// +infinity and never moves.  Every team thread calls this with the same arguments; n log^2 n / Tn work each.
__device__ __noinline__ void fix_stack_team(const Team& tm, int* st, int lo, int hi) {
    const int n = hi - lo + 1;
    int P = 1;
    while (P < n) P <<= 1;
    team_sync(tm);
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool mirror = j == (k >> 1);
            for (int t = tm.T; t < (P >> 1); t += tm.Tn) {
                // the t-th pair of this sub-stage: i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = mirror ? (i ^ (k - 1)) : (i | j);
                if (l < n) {
                    const int a = __ldcg(st + lo + i), b = __ldcg(st + lo + l);
                    if (a > b) {
                        st[lo + i] = b;
                        st[lo + l] = a;
                    }
                }
            }
            __threadfence();
            team_sync(tm);
        }
    }
}

// All team threads, uniformly.  Returns after the stacks are in order and the counters are cleared.
__device__ __noinline__ void fix_stack_order(const Team& tm, const StackFix& f) {
    const int n0 = __ldcg(f.pushed[0]), n1 = __ldcg(f.pushed[1]);
    if ((n0 | n1) == 0) return;
    bool synced = false;
    for (int s = 0; s < 2; ++s) {
        const int n = s ? n1 : n0;
        if (n == 0) continue;
        const int top = __ldcg(f.top[s]);
        const int hi = min(top, f.largest_allowed_slot[s]);
        const int lo = max(top - n + 1, 0);
        if (hi - lo + 1 <= kFixSmall) {
            if (tm.T == 0) fix_stack_small(f.empty_slots[s], lo, hi);
        } else {
            fix_stack_team(tm, f.empty_slots[s], lo, hi);
            synced = true;
        }
    }
    if (synced) team_sync(tm);   // every thread has read the counters before they are cleared
    if (tm.T == 0) {
        *f.pushed[0] = 0;
        *f.pushed[1] = 0;
    }
}

template <class T>
__device__ __forceinline__ T* remote(T* p, int rank) {
    return cg::this_cluster().map_shared_rank(p, rank);
}

// Transposed run storage: element i of a team thread's run, in a global plane (stride Tn) or in
// this CTA's shared memory (stride 1024).
struct Plane {
    double* p;
    int stride, idx;
    __device__ __forceinline__ double& operator()(int i) const { return p[(size_t)i * stride + idx]; }
};

// Cluster-level exchange buffers (per CTA; every CTA holds a full copy, written by its peers).
// Two parities: exchange k+2 reuses the buffers of exchange k, which every CTA has finished
// reading before any CTA can pass the barrier of exchange k+1.
struct TeamShared {
    Mat2 mat_warp[32];
    Aff aff_warp[32];
    Mat2 mat_tot[2][kMaxTeam];
    Aff aff_tot[2][kMaxTeam];
    Aff aff_warps[2][32];           // the CTA's warp composites, in scan order (team_scan_aff)
    double bcast[2][2];
    double sum[2][kMaxTeam];
    double red[32];
};
__device__ __forceinline__ Mat2* warp_buf(TeamShared& sh, const Mat2&) { return sh.mat_warp; }
__device__ __forceinline__ Aff* warp_buf(TeamShared& sh, const Aff&) { return sh.aff_warp; }
__device__ __forceinline__ Mat2* tot_buf(TeamShared& sh, int par, const Mat2&) { return sh.mat_tot[par]; }
__device__ __forceinline__ Aff* tot_buf(TeamShared& sh, int par, const Aff&) { return sh.aff_tot[par]; }

// One value from one thread to every CTA of the team; visible after the next team exchange
// (team_exclusive_scan / team_barrier_exchange) of the same phase.
__device__ __forceinline__ void team_publish(const Team& tm, TeamShared& sh, int phase, int slot, double v) {
    for (int r = 0; r < tm.size; ++r) remote(&sh.bcast[phase & 1][0], r)[slot] = v;
}

// Exclusive scan over the team in team-thread order (reverse: CTAs in descending rank order; the
// caller mirrors the threads inside the CTA).  Advances `phase`.
template <class T>
__device__ __forceinline__ T team_exclusive_scan(const Team& tm, TeamShared& sh, T mine, T identity, int& phase,
                                                 bool reverse) {
    const T res = block_exclusive_scan(mine, identity, warp_buf(sh, mine));
    const int par = phase & 1;
    ++phase;
    if (tm.size == 1) {
        __syncthreads();  // publishes of this phase
        return res;
    }
    if (threadIdx.x == kSolveThreads - 1) {
        const T tot = compose(mine, res);
        for (int r = 0; r < tm.size; ++r) remote(tot_buf(sh, par, mine), r)[tm.rank] = tot;
    }
    cg::this_cluster().sync();
    T pre = identity;
    const T* tot = tot_buf(sh, par, mine);
    if (!reverse) {
        for (int r = 0; r < tm.rank; ++r) pre = compose(tot[r], pre);
    } else {
        for (int r = tm.size - 1; r > tm.rank; --r) pre = compose(tot[r], pre);
    }
    return compose(res, pre);
}

// Exclusive scan of affine maps over the team with one CTA barrier and one cluster barrier: every
// warp scans its lanes with shuffles and posts its composite; after the CTA barrier every warp
// scans the 32 posted composites for itself (no second barrier), one warp writes the CTA's total
// into every CTA's table (8 DSMEM stores per CTA -- remote stores are slow, ~50 ns each: posting all
// 256 warp composites to every CTA instead cost 7 us per scan), and after the cluster barrier the
// totals of the CTAs before this one are folded in.  reverse: the same in descending team-thread
// order (lanes, warps and CTAs mirrored), for the back substitution.  Advances `phase`.
__device__ __forceinline__ Aff team_scan_aff(const Team& tm, TeamShared& sh, Aff mine, int& phase, bool reverse) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int vlane = reverse ? 31 - lane : lane;  // positions in scan order
    const int vwarp = reverse ? 31 - w : w;
    const int vrank = reverse ? tm.size - 1 - tm.rank : tm.rank;
    const int par = phase & 1;
    ++phase;
    Aff incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int src = reverse ? lane + d : lane - d;
        const Aff other{__shfl_sync(0xffffffffu, incl.A, src & 31), __shfl_sync(0xffffffffu, incl.B, src & 31)};
        if (vlane >= d) incl = compose(incl, other);
    }
    if (vlane == 31) sh.aff_warps[par][vwarp] = incl;
    __syncthreads();
    Aff warps = sh.aff_warps[par][lane];  // inclusive scan of the 32 warp composites, redundantly per warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Aff other{__shfl_up_sync(0xffffffffu, warps.A, d), __shfl_up_sync(0xffffffffu, warps.B, d)};
        if (lane >= d) warps = compose(warps, other);
    }
    Aff before{__shfl_sync(0xffffffffu, warps.A, (vwarp - 1) & 31), __shfl_sync(0xffffffffu, warps.B, (vwarp - 1) & 31)};
    if (vwarp == 0) before = aff_identity();
    if (tm.size > 1) {
        const Aff cta{__shfl_sync(0xffffffffu, warps.A, 31), __shfl_sync(0xffffffffu, warps.B, 31)};
        if (w == 0 && lane < tm.size) remote(&sh.aff_tot[par][0], lane)[vrank] = cta;
        cg::this_cluster().sync();
        Aff ctas = aff_identity();  // the CTAs before this one, earliest applied first
        for (int r = 0; r < vrank; ++r) ctas = compose(sh.aff_tot[par][r], ctas);
        before = compose(before, ctas);
    }
    const int prev = reverse ? lane + 1 : lane - 1;
    Aff excl{__shfl_sync(0xffffffffu, incl.A, prev & 31), __shfl_sync(0xffffffffu, incl.B, prev & 31)};
    if (vlane == 0) excl = aff_identity();
    return compose(excl, before);
}

// Sum over the team, CTAs added in rank order (deterministic).  Advances `phase`.
__device__ __forceinline__ double team_sum(const Team& tm, TeamShared& sh, double v, int& phase) {
    const double mine = block_sum(v, sh.red);
    const int par = phase & 1;
    ++phase;
    if (tm.size == 1) return mine;
    if (threadIdx.x == 0)
        for (int r = 0; r < tm.size; ++r) remote(&sh.sum[par][0], r)[tm.rank] = mine;
    cg::this_cluster().sync();
    double tot = 0.;
    for (int r = 0; r < tm.size; ++r) tot += sh.sum[par][r];
    return tot;
}

// s_arr[t] as published by the previous / next team thread (run-boundary neighbours); the caller
// has written s_arr and passed team_sync().
__device__ __forceinline__ double prev_thread_value(const Team& tm, double* s_arr, double fallback) {
    const int t = threadIdx.x;
    if (t > 0) return s_arr[t - 1];
    return tm.rank > 0 ? remote(s_arr, tm.rank - 1)[kSolveThreads - 1] : fallback;
}
__device__ __forceinline__ double next_thread_value(const Team& tm, double* s_arr, double fallback) {
    const int t = threadIdx.x;
    if (t + 1 < kSolveThreads) return s_arr[t + 1];
    return tm.rank + 1 < tm.size ? remote(s_arr, tm.rank + 1)[0] : fallback;
}

constexpr int kHeaderDoubles = 16;
constexpr unsigned long long kFactorMagic = 0x45535049434c5531ull;  // "ESPICLU1"

// One row of the system exactly as the reference builds it (ref infMagSim_c.c:385-503):
// plain [1,-2,1] / Robin / periodic rows, the absorbing-object variants, and their blend by
// object_transparency.  low = the sub-diagonal element of ROW j (the reference's lower_diagonal[j-1]).
struct RowConsts {
    float dz, lam, rhs_scale;
    double robin_diag, robin_off, near_one, tr, op;
};

__device__ __forceinline__ RowConsts row_consts(const SolveArgs& a) {
    RowConsts k;
    const int n = a.n;
    k.dz = (__ldg(a.grid + n - 1) - __ldg(a.grid)) / (float)(n - 1);
    k.lam = a.debye;
    const float one = 1.f;
    const float h2 = k.dz * k.dz / k.lam / k.lam;
    const float h1 = k.dz / k.lam;
    k.robin_diag = (double)(h2 * one) / 2. + (double)h1;
    k.robin_off = (double)(h2 * one) / 2. - (double)h1;
    k.rhs_scale = -k.dz * k.dz / k.lam / k.lam;
    k.near_one = 1. - (double)1.e-5f;
    k.tr = (double)a.transparency;
    k.op = 1. - (double)a.transparency;
    return k;
}

__device__ __forceinline__ void build_row(const SolveArgs& a, const RowConsts& k, const float* charge, int j,
                                          double& low, double& dg, double& sup, double& rhs, int& bad_mask) {
    const int n = a.n;
    const float dz = k.dz, lam = k.lam;
    double p_low = 1., p_dg = -2., p_sup = 1.;
    double p_rhs = (j == n - 1) ? 0. : (double)charge[j];
    if (a.boltzmann) {
        const double ephi = exp((double)a.potential[j]);
        p_dg -= (double)(dz * dz) * ephi / (double)lam / (double)lam;
        p_rhs -= ephi * (1. - (double)a.potential[j]);
    }
    p_rhs *= (double)k.rhs_scale;
    if (j == 0) {
        p_dg = k.robin_diag;
        p_sup = k.robin_off;
        p_rhs = 0.;
    }
    if (j == n - 1) {
        if (a.periodic) {
            p_low = 0.;
            p_dg = 1.;
        } else {
            p_low = k.robin_off;
            p_dg = k.robin_diag;
        }
        p_rhs = 0.;
    }
    double o_low = p_low, o_dg = p_dg, o_sup = p_sup, o_rhs = p_rhs;
    if (k.op != 0. && j >= 1 && j <= n - 2) {  // the object rows only matter when they are blended in
        const float mk = a.mask[j], mp = a.mask[j - 1];
        if (mk > 0.f) {
            if ((double)mk < k.near_one) {
                if (a.mask[j + 1] > 0.f) {
                    const float zl = (float)((1. - (double)mk) * (double)dz);
                    o_low = (double)(zl / dz);
                    o_dg = -1. - (double)(zl / dz);
                    o_sup = 0.;
                    o_rhs *= (double)(zl / dz);
                    o_rhs -= (double)a.obj_pot;
                } else if ((double)mp < k.near_one) {
                    const float zl = (float)((1. - (double)mp) * (double)dz);
                    const float zr = (float)((1. - (double)mk) * (double)dz);
                    o_low = -(1. - (double)(zl / dz));
                    o_dg = (double)(-(zl + zr) / dz);
                    o_sup = -(1. - (double)(zr / dz));
                    o_rhs = -2. * (double)a.obj_pot;
                } else if ((double)mp > k.near_one) {
                    const float zr = (float)((1. - (double)mk) * (double)dz);
                    o_low = 0.;
                    o_dg = -2. * (double)zr / (double)dz;
                    o_sup = -2. * (1. - (double)(zr / dz));
                    o_rhs = -2. * (double)a.obj_pot;
                } else {
                    bad_mask = 1;
                }
            } else if ((double)mk > k.near_one) {
                if ((double)mp < k.near_one) {
                    const float zl = (float)((1. - (double)mp) * (double)dz);
                    o_low = -2. * (1. - (double)(zl / dz));
                    o_dg = -2. * (double)zl / (double)dz;
                    o_sup = 0.;
                    o_rhs = -2. * (double)a.obj_pot;
                } else if ((double)mp > k.near_one) {
                    o_rhs = 0.;
                } else {
                    bad_mask = 1;
                }
            } else {
                bad_mask = 1;
            }
        } else if (mp > 0.f) {
            const float zr = (float)((1. - (double)mp) * (double)dz);
            o_low = 0.;
            o_dg = -1. - (double)(zr / dz);
            o_sup = (double)(zr / dz);
            o_rhs *= (double)(zr / dz);
            o_rhs -= (double)a.obj_pot;
        }
    }
    low = p_low * k.tr + o_low * k.op;
    dg = p_dg * k.tr + o_dg * k.op;
    sup = p_sup * k.tr + o_sup * k.op;
    rhs = p_rhs * k.tr + o_rhs * k.op;
}

// D[0] = dg[0]; D[j] = dg[j] - (sup[j-1]*low[j])/D[j-1].
// Runs are m rows per team thread: thread T owns rows T*m .. T*m+m-1 (clipped to n).
__device__ __forceinline__ void eliminate_diagonal(const Team& tm, TeamShared& sh, int& phase, const Plane& dg,
                                                   const Plane& low, const Plane& sup, const Plane& D, int n, int m,
                                                   double* s_edge) {
    const int t = threadIdx.x;
    const int j0 = tm.T * m;
    // sup[j0-1] lives in the previous thread's run: publish each run's last sup
    s_edge[t] = (j0 + m - 1 < n) ? sup(m - 1) : 0.;
    team_sync(tm);
    const double sup_before = prev_thread_value(tm, s_edge, 0.);
    Mat2 comp = mat_identity();
    for (int i = 0; i < m; ++i) {
        const int j = j0 + i;
        if (j >= n || j == 0) continue;
        const double sp = (i == 0) ? sup_before : sup(i - 1);
        comp = compose(Mat2{dg(i), -(sp * low(i)), 1., 0.}, comp);
    }
    // value entering the first run: D[0] = dg[0]
    if (tm.T == 0) team_publish(tm, sh, phase, 0, dg(0));
    const int par = phase & 1;
    const Mat2 pre = team_exclusive_scan(tm, sh, comp, mat_identity(), phase, false);
    const double d0 = sh.bcast[par][0];
    double d = (pre.a * d0 + pre.b) / (pre.c * d0 + pre.d);
    for (int i = 0; i < m; ++i) {
        const int j = j0 + i;
        if (j >= n) break;
        if (j == 0) {
            d = d0;
        } else {
            const double sp = (i == 0) ? sup_before : sup(i - 1);
            d = dg(i) - sp * low(i) / d;
        }
        D(i) = d;
    }
    team_sync(tm);
}

// Multiplier planes of one Thomas factorisation: fwd[j] = low[j]/D[j-1] (j>=1),
// bwd[j] = sup[j]/D[j+1] (j<=n-2), inv[j] = 1/D[j].
__device__ __forceinline__ void make_multipliers(const Team& tm, const Plane& low, const Plane& sup, const Plane& D,
                                                 const Plane& fwd, const Plane& bwd, const Plane& inv, int n, int m,
                                                 double* s_edge) {
    const int t = threadIdx.x, j0 = tm.T * m;
    // neighbours across run boundaries: last D of the previous run, first D of the next run
    s_edge[t] = (j0 + m - 1 < n) ? D(m - 1) : 1.;
    s_edge[kSolveThreads + t] = (j0 < n) ? D(0) : 1.;
    team_sync(tm);
    const double D_before = prev_thread_value(tm, s_edge, 1.);
    const double D_after = next_thread_value(tm, s_edge + kSolveThreads, 1.);
    for (int i = 0; i < m; ++i) {
        const int j = j0 + i;
        if (j >= n) break;
        const double Dp = (i == 0) ? D_before : D(i - 1);
        const double Dn = (i == m - 1) ? D_after : D(i + 1);
        fwd(i) = (j >= 1) ? low(i) / Dp : 0.;
        bwd(i) = (j <= n - 2) ? sup(i) / Dn : 0.;
        inv(i) = 1. / D(i);
    }
    team_sync(tm);
}

// The sweeps below are written without data-dependent control flow and unrolled by 8 so that the
// plane loads of the next iterations are in flight while the recurrence of the current one runs.
// Plane element i < m exists for every thread, so rows past n may be loaded; they are masked
// out of the arithmetic and never stored.

// in place: y[j] -= y[j-1]*fwd[j], j = 1..n-1
__device__ __forceinline__ void forward_sweep(const Team& tm, TeamShared& sh, int& phase, const Plane& y,
                                              const Plane& fwd, int n, int m) {
    const int j0 = tm.T * m;
    Aff comp = aff_identity();
#pragma unroll 8
    for (int i = 0; i < m; ++i) {
        const int j = j0 + i;
        const double f = fwd(i), yy = y(i);
        const Aff next = compose(Aff{-f, yy}, comp);
        const bool on = (j < n) && (j != 0);
        comp.A = on ? next.A : comp.A;
        comp.B = on ? next.B : comp.B;
    }
    if (tm.T == 0) team_publish(tm, sh, phase, 0, y(0));
    const int par = phase & 1;
    const Aff pre = team_scan_aff(tm, sh, comp, phase, false);
    double prev = pre.A * sh.bcast[par][0] + pre.B;
#pragma unroll 8
    for (int i = 0; i < m; ++i) {
        const int j = j0 + i;
        const double f = fwd(i), yy = y(i);
        const double r = (j == 0) ? yy : yy - prev * f;
        if (j < n) {
            prev = r;
            y(i) = r;
        }
    }
}

// in place: z[j] -= z[j+1]*bwd[j], j = n-2..0.  The scan runs in REVERSE team-thread order.
__device__ __forceinline__ void backward_sweep(const Team& tm, TeamShared& sh, int& phase, const Plane& z,
                                               const Plane& bwd, int n, int m) {
    const int j0 = tm.T * m;
    Aff comp = aff_identity();
#pragma unroll 8
    for (int i = m - 1; i >= 0; --i) {
        const int j = j0 + i;
        const double b = bwd(i), zz = z(i);
        const Aff next = compose(Aff{-b, zz}, comp);
        const bool on = j < n - 1;  // row n-1 is the start value; rows past n do not exist
        comp.A = on ? next.A : comp.A;
        comp.B = on ? next.B : comp.B;
    }
    {
        const int Tl = (n - 1) / m, il = (n - 1) - Tl * m;  // owner of row n-1
        if (tm.T == Tl) team_publish(tm, sh, phase, 0, z(il));
    }
    const int par = phase & 1;
    const Aff pre = team_scan_aff(tm, sh, comp, phase, true);
    double next = pre.A * sh.bcast[par][0] + pre.B;
#pragma unroll 8
    for (int i = m - 1; i >= 0; --i) {
        const int j = j0 + i;
        const double b = bwd(i), zz = z(i);
        const double r = (j == n - 1) ? zz : zz - next * b;
        if (j <= n - 1) {
            next = r;
            z(i) = r;
        }
    }
}

// Workspace (doubles): [header 16] then planes of m*Tn: low, dg, sup, D, D2 (scratch of the
// factorisation), fwd, bwd, inv, wq (the factorisation itself, kept between calls) and y.
// The factorisation depends only on (grid ends, n, debye_length, periodic) while the object
// rows are blended out (object_transparency == 1) and the electrons are not Boltzmann -- every
// driver step of every configuration -- so it is computed on the first call, recorded under that
// key in the header, and later calls go straight to the two sweeps.  Both routes end in the
// same sweeps over the same planes, so a cached and an uncached call return identical bits.
struct FactorPlanes {
    Plane low, dg, sup, D, D2, fwd, bwd, inv, wq;
};

// First call for a matrix (and every call whose matrix is not cacheable): rows, eliminated diagonal,
// multiplier planes, periodic correction vector.  Kept out of line so that the per-step path of
// k_poisson stays short and contiguous in the instruction cache.
__device__ __noinline__ void factorise(const SolveArgs& a, const RowConsts& k, const Team& tm, TeamShared& sh,
                                       int& phase, const float* charge, const FactorPlanes& P, double* hdr, int m,
                                       bool cacheable, double* s_edge, int& bad_mask) {
    const int n = a.n, tid = threadIdx.x;
    const Plane &low = P.low, &dg = P.dg, &sup = P.sup, &D = P.D, &D2 = P.D2, &fwd = P.fwd, &bwd = P.bwd, &inv = P.inv,
                &wq = P.wq;
    // ---- factorisation: rows (ref :385-503), eliminated diagonal (ref :343-349), multipliers
    for (int i = 0; i < m; ++i) {
        const int j = tm.T * m + i;
        if (j >= n) break;
        double r_low, r_dg, r_sup, r_rhs;
        build_row(a, k, charge, j, r_low, r_dg, r_sup, r_rhs, bad_mask);
        low(i) = r_low;
        dg(i) = r_dg;
        sup(i) = r_sup;
    }
    __syncthreads();
    eliminate_diagonal(tm, sh, phase, dg, low, sup, D, n, m, s_edge);
    make_multipliers(tm, low, sup, D, fwd, bwd, inv, n, m, s_edge);
    if (a.periodic) {  // ref :504-507: the second solve re-eliminates the eliminated diagonal (quirk 1)
        eliminate_diagonal(tm, sh, phase, D, low, sup, D2, n, m, s_edge);
        // w = "solve" of u = e_{n-1} on that matrix: the forward sweep leaves u unchanged, the
        // backward sweep and the final division act.  dg / low are free now: scratch for them.
        const Plane bwd2 = dg, wz = low;
        {
            const int j0 = tm.T * m;
            s_edge[kSolveThreads + tid] = (j0 < n) ? D2(0) : 1.;
            team_sync(tm);
            const double D_after = next_thread_value(tm, s_edge + kSolveThreads, 1.);
            for (int i = 0; i < m; ++i) {
                const int j = j0 + i;
                if (j >= n) break;
                const double Dn = (i == m - 1) ? D_after : D2(i + 1);
                const double sp = sup(i);
                bwd2(i) = (j <= n - 2) ? sp / Dn : 0.;
            }
            for (int i = 0; i < m; ++i) {
                const int j = j0 + i;
                if (j < n) wz(i) = (j == n - 1) ? 1. : 0.;
            }
            team_sync(tm);
        }
        backward_sweep(tm, sh, phase, wz, bwd2, n, m);
        for (int i = 0; i < m; ++i)
            if (tm.T * m + i < n) wq(i) = wz(i) / D2(i);
        __syncthreads();
    }
    if (tm.T == 0 && cacheable) {
        hdr[1] = (double)__ldg(a.grid);
        hdr[2] = (double)__ldg(a.grid + n - 1);
        hdr[3] = (double)a.debye;
        hdr[4] = (double)n;
        hdr[5] = (double)a.periodic;
        hdr[6] = (double)tm.size;
        __threadfence();
        reinterpret_cast<unsigned long long*>(hdr)[0] = kFactorMagic;
    } else if (tm.T == 0) {
        reinterpret_cast<unsigned long long*>(hdr)[0] = 0ull;  // the planes no longer match any key
    }
}

// Right-hand side through the full row builder (object rows blended in, Boltzmann electrons).
__device__ __noinline__ void rhs_general(const SolveArgs& a, const RowConsts& k, const Team& tm, const float* charge,
                                         const Plane& y, int m, int& bad_mask) {
    for (int i = 0; i < m; ++i) {
        const int j = min(tm.T * m + i, a.n - 1);
        double r_low, r_dg, r_sup, r_rhs;
        build_row(a, k, charge, j, r_low, r_dg, r_sup, r_rhs, bad_mask);
        y(i) = r_rhs;
    }
}

// ONE_ROW: every team thread owns exactly one row (n <= Tn; the 4097-point grid on 8 CTAs).  The
// run loops collapse at compile time, which keeps the executed path short: the kernel starts with a
// cold instruction cache every step (the movers in between evict it), and the general variant,
// whose sweeps are unrolled for long runs, spent 40 % of its 23 us waiting for instructions.
template <bool ONE_ROW>
__global__ void __launch_bounds__(kSolveThreads) k_poisson(const __grid_constant__ SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];  // y plane when it fits (m <= kSolveSmemRows)
    __shared__ TeamShared sh;
    __shared__ __align__(16) double s_edge[2 * kSolveThreads];  // run-boundary values; reused as Aff[1024]
    __shared__ int s_cached;
    __shared__ float s_end[4];
    const Team tm = make_team();
    int phase = 0;
    const int n = a.n, tid = threadIdx.x;
    const int m = ONE_ROW ? 1 : (n + tm.Tn - 1) / tm.Tn;
    const size_t plane = (size_t)m * tm.Tn;
    double* hdr = a.ws;
    double* base = a.ws + kHeaderDoubles;
    const Plane low{base, tm.Tn, tm.T}, dg{base + plane, tm.Tn, tm.T}, sup{base + 2 * plane, tm.Tn, tm.T},
        D{base + 3 * plane, tm.Tn, tm.T}, D2{base + 4 * plane, tm.Tn, tm.T}, fwd{base + 5 * plane, tm.Tn, tm.T},
        bwd{base + 6 * plane, tm.Tn, tm.T}, inv{base + 7 * plane, tm.Tn, tm.T}, wq{base + 8 * plane, tm.Tn, tm.T};
    const Plane y = (m <= kSolveSmemRows) ? Plane{reinterpret_cast<double*>(smem_raw), kSolveThreads, tid}
                                          : Plane{base + 9 * plane, tm.Tn, tm.T};

    // The factorisation planes were last touched a whole step ago and the movers have streamed
    // gigabytes through L2 since: start pulling this thread's elements back in now, so that the
    // sweeps below meet L2 hits instead of one DRAM round trip per dependent load.
    if ((tid & 15) == 0) {  // one request per 128-byte line
        for (int i = 0; i < m; ++i) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(&fwd(i)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(&bwd(i)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(&inv(i)));
            if (a.periodic) asm volatile("prefetch.global.L2 [%0];" ::"l"(&wq(i)));
        }
    }

    // ---- multi-GPU: the all-reduce of the two density grids (ref infMagSim_script.py:763, 773),
    // done by every rank for itself over NVLink peer memory.  Ranks are summed in rank order,
    // so all ranks obtain bit-identical densities and therefore bit-identical fields.
    int timed_out = 0;  // CTA-uniform after the barrier below
    if (a.world > 1) {
        int late = 0;
        if (tid < a.world) {
            // wait until peer `tid` has published its densities for this step (system-scope acquire)
            const unsigned* flag = a.peer_flag[tid];
            unsigned long long t0 = 0ull;
            unsigned seen;
            for (unsigned spins = 0;; ++spins) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
                if ((int)(seen - a.expect) >= 0) break;
                if ((spins & 63u) == 63u) {  // the wall clock is only consulted once the wait is long
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0ull) t0 = now;
                    if (now - t0 > a.timeout_ns) {
                        // a peer never published (it died, or lags by more than the configured bound):
                        // FATAL.  The status bit is raised and this CTA's share of the potential is
                        // poisoned with NaN below -- nothing is solved on stale or half-written slots.
                        atomicOr(a.status, ESPIC_ST_EXCHANGE_TIMEOUT);
                        late = 1;
                        break;
                    }
                }
                __nanosleep(200);
            }
        }
        timed_out = __syncthreads_or(late);
        for (int j = tm.T; j < n; j += tm.Tn) {
            float si = 0.f, se = 0.f;
            for (int r = 0; r < a.world; ++r) {
                const float* slot = a.peer_slot[r];
                si = __fadd_rn(si, __ldcv(slot + j));
                se = __fadd_rn(se, __ldcv(slot + a.n_pad + j));
            }
            a.ion[j] = si;
            a.ele[j] = se;
        }
        team_sync(tm);  // the cluster barrier releases / acquires the global writes above
    }

    // ---- fused driver prologue: end-node fix + charge density (ref infMagSim_script.py:763-779, 807)
    const RowConsts k = row_consts(a);
    const bool cacheable = (a.transparency == 1.f) && !a.boltzmann;
    if (tid == 0) {
        if (a.ion) {  // every CTA works the fixed end values out for itself; CTA 0 stores them
            float i0 = a.acc_i ? (float)((double)__ldcg(a.acc_i) * a.inv_scale) : a.ion[0];
            float i1 = a.acc_i ? (float)((double)__ldcg(a.acc_i + n - 1) * a.inv_scale) : a.ion[n - 1];
            float e0 = a.acc_e ? (float)((double)__ldcg(a.acc_e) * a.inv_scale) : a.ele[0];
            float e1 = a.acc_e ? (float)((double)__ldcg(a.acc_e + n - 1) * a.inv_scale) : a.ele[n - 1];
            if (a.fold_ends) {
                if (a.fix_i) { i0 = __fadd_rn(i0, i1); i1 = i0; }
                if (a.fix_e) { e0 = __fadd_rn(e0, e1); e1 = e0; }
            } else {
                if (a.fix_i) { i0 = __fmul_rn(i0, 2.f); i1 = __fmul_rn(i1, 2.f); }
                if (a.fix_e) { e0 = __fmul_rn(e0, 2.f); e1 = __fmul_rn(e1, 2.f); }
            }
            s_end[0] = i0; s_end[1] = i1; s_end[2] = e0; s_end[3] = e1;
        }
        const unsigned long long* h = reinterpret_cast<const unsigned long long*>(hdr);
        s_cached = cacheable && h[0] == kFactorMagic && hdr[1] == (double)__ldg(a.grid) &&
                   hdr[2] == (double)__ldg(a.grid + n - 1) && hdr[3] == (double)a.debye && hdr[4] == (double)n &&
                   hdr[5] == (double)a.periodic && hdr[6] == (double)tm.size;
    }
    team_sync(tm);  // every CTA has read the raw end values and the workspace key
    const float* charge = a.charge;
    if (a.ion) {
        if (tm.T == 0) {
            a.ion[0] = s_end[0]; a.ion[n - 1] = s_end[1]; a.ele[0] = s_end[2]; a.ele[n - 1] = s_end[3];
        }
        // Fused deposit still in the accumulators.  One row per thread: the row's owner converts it in place
        // (coalesced as it is).  Several rows per thread (the transposed layout, j = T*m + i): a conversion
        // pass of its own with consecutive threads on consecutive nodes, then one more barrier -- walking the
        // 8-byte accumulators with a stride of m rows made the 65537-point solve twice as slow.
        bool pending_i = a.acc_i != nullptr, pending_e = a.acc_e != nullptr;
        if (!ONE_ROW && (pending_i || pending_e)) {
            for (int j = tm.T; j < n; j += tm.Tn) {
                const bool interior = j != 0 && j != n - 1;
                if (pending_i) {
                    const float x = (float)((double)__ldcg(a.acc_i + j) * a.inv_scale);
                    a.acc_i[j] = 0ull;
                    if (interior) a.ion[j] = x;
                }
                if (pending_e) {
                    const float y = (float)((double)__ldcg(a.acc_e + j) * a.inv_scale);
                    a.acc_e[j] = 0ull;
                    if (interior) a.ele[j] = y;
                }
            }
            team_sync(tm);
            pending_i = pending_e = false;
        }
        // each thread makes the charge of its own rows: nobody else reads them in this kernel
        for (int i = 0; i < m; ++i) {
            const int j = tm.T * m + i;
            if (j >= n) break;
            float x, yv;
            if (pending_i) {  // every CTA has read the raw end nodes before the barrier above: the words can go
                x = (float)((double)__ldcg(a.acc_i + j) * a.inv_scale);
                a.acc_i[j] = 0ull;
                if (j != 0 && j != n - 1) a.ion[j] = x;
            } else {
                x = a.ion[j];
            }
            if (pending_e) {
                yv = (float)((double)__ldcg(a.acc_e + j) * a.inv_scale);
                a.acc_e[j] = 0ull;
                if (j != 0 && j != n - 1) a.ele[j] = yv;
            } else {
                yv = a.ele[j];
            }
            x = (j == 0) ? s_end[0] : (j == n - 1) ? s_end[1] : x;
            yv = (j == 0) ? s_end[2] : (j == n - 1) ? s_end[3] : yv;
            a.charge_out[j] = __fsub_rn(x, yv);
        }
        charge = a.charge_out;
    }
    int bad_mask = 0;

    if (!s_cached)
        factorise(a, k, tm, sh, phase, charge, FactorPlanes{low, dg, sup, D, D2, fwd, bwd, inv, wq}, hdr, m, cacheable,
                  s_edge, bad_mask);

    // ---- solve: right-hand side, forward and backward sweeps (ref :350-353), output (ref :518-525)
    if (cacheable) {
        // transparency 1, no Boltzmann term: the blend rhs*1 + rhs*0 of build_row is the plain row's
        // right-hand side bit for bit (ref :392-395, :409, :428)
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const int j = min(tm.T * m + i, n - 1);  // rows past n: a copy of the last row, never used
            y(i) = (j == 0 || j == n - 1) ? 0. : (double)charge[j] * (double)k.rhs_scale;
        }
    } else {
        rhs_general(a, k, tm, charge, y, m, bad_mask);
    }
    if (bad_mask) atomicOr(a.status, ESPIC_ST_BAD_OBJECT_MASK);
    __syncthreads();
    forward_sweep(tm, sh, phase, y, fwd, n, m);
    backward_sweep(tm, sh, phase, y, bwd, n, m);
    if (a.periodic) {  // ref :508-514, :518-523
        if (tm.T == 0) team_publish(tm, sh, phase, 1, y(0) * inv(0));
        const int par = phase & 1;
        // (the sum below is the team exchange that makes the published value visible; it needs the
        // value itself, so the publication gets its own exchange first)
        (void)team_sum(tm, sh, 0., phase);
        const double v_dot_result = -sh.bcast[par][1];
        const double v_dot_w = 0.;  // quirk (2)
        double part = 0.;
        for (int i = 0; i < m; ++i) {
            if (tm.T * m + i >= n) break;
            const double r = y(i) * inv(i) - v_dot_result / (1. + v_dot_w) * wq(i);
            y(i) = r;
            part += r;
        }
        const double mean = team_sum(tm, sh, part, phase) / (double)n;
        for (int i = 0; i < m; ++i) {
            const int j = tm.T * m + i;
            if (j < n) a.potential[j] = (float)(y(i) - mean);
        }
    } else {
#pragma unroll 8
        for (int i = 0; i < m; ++i) {
            const int j = tm.T * m + i;
            const double r = y(i) * inv(i);
            if (j < n) a.potential[j] = (float)r;
        }
    }
    // stacks whose newest entries the periodic movers pushed in arrival order (practically never): put them in the
    // reference's ascending order; checked last, off the critical path of the solve
    if (a.fix.pushed[0]) fix_stack_order(tm, a.fix);
    if (timed_out) {  // exchange timeout: every barrier above was still taken; the result is withdrawn
        for (int i = 0; i < m; ++i) {
            const int j = tm.T * m + i;
            if (j < n) a.potential[j] = __int_as_float(0x7fc00000);
        }
    }
}

// ---- tridiagonal_solve_c (ref infMagSim_c.c:343-354) on device arrays -------------------------
// The reference's generic Thomas routine: b and d are modified in place, x = d/b.  Its three
// recurrences are strictly sequential and a caller may rely on the exact bits of b and d, so one
// thread walks them in the reference's operation order ((d*a)/b, (c*a)/b: separately rounded
// multiply, divide and subtract -- the library is built without FMA contraction) while the CTA
// stages the rows through shared memory in coalesced chunks.  Not on the step path (the field
// solve has its own scan kernels); a 65537-row system takes a few milliseconds.
constexpr int kTriChunk = 1024;

__global__ void __launch_bounds__(256) k_tridiagonal(const double* __restrict__ a, double* __restrict__ b,
                                                     const double* __restrict__ c, double* __restrict__ d,
                                                     double* __restrict__ x, int n) {
    __shared__ double sa[kTriChunk], sc[kTriChunk], sb[kTriChunk], sd[kTriChunk];
    __shared__ double s_carry[2];
    const int tid = threadIdx.x;
    // forward elimination, rows r = 1 .. n-1 (ref :346-349)
    if (tid == 0) { s_carry[0] = b[0]; s_carry[1] = d[0]; }
    for (int r0 = 1; r0 < n; r0 += kTriChunk) {
        const int len = min(kTriChunk, n - r0);
        for (int i = tid; i < len; i += blockDim.x) {
            sa[i] = a[r0 + i - 1];
            sc[i] = c[r0 + i - 1];
            sb[i] = b[r0 + i];
            sd[i] = d[r0 + i];
        }
        __syncthreads();
        if (tid == 0) {
            double bp = s_carry[0], dp = s_carry[1];
            for (int i = 0; i < len; ++i) {
                const double dn = __dsub_rn(sd[i], __ddiv_rn(__dmul_rn(dp, sa[i]), bp));
                const double bn = __dsub_rn(sb[i], __ddiv_rn(__dmul_rn(sc[i], sa[i]), bp));
                sd[i] = dn;
                sb[i] = bn;
                dp = dn;
                bp = bn;
            }
            s_carry[0] = bp;
            s_carry[1] = dp;
        }
        __syncthreads();
        for (int i = tid; i < len; i += blockDim.x) {
            b[r0 + i] = sb[i];
            d[r0 + i] = sd[i];
        }
        __syncthreads();
    }
    // back substitution, rows j = n-2 .. 0 (ref :350-351): d[j] -= d[j+1]*c[j]/b[j+1]
    if (tid == 0) s_carry[1] = d[n - 1];
    __syncthreads();
    for (int hi = n - 1; hi > 0; hi -= kTriChunk) {  // rows [lo, hi)
        const int lo = max(0, hi - kTriChunk), len = hi - lo;
        for (int i = tid; i < len; i += blockDim.x) {
            sc[i] = c[lo + i];
            sb[i] = b[lo + i + 1];
            sd[i] = d[lo + i];
        }
        __syncthreads();
        if (tid == 0) {
            double dn = s_carry[1];
            for (int i = len - 1; i >= 0; --i) {
                dn = __dsub_rn(sd[i], __ddiv_rn(__dmul_rn(dn, sc[i]), sb[i]));
                sd[i] = dn;
            }
            s_carry[1] = dn;
        }
        __syncthreads();
        for (int i = tid; i < len; i += blockDim.x) d[lo + i] = sd[i];
        __syncthreads();
    }
    for (int j = tid; j < n; j += blockDim.x) x[j] = __ddiv_rn(d[j], b[j]);  // ref :352-353
}

int launch_tridiagonal(espic_ctx* ctx, double* a, double* b, double* c, double* d, double* x, int n) {
    if (n < 1) return fail(ctx, ESPIC_E_ARG, "tridiagonal_solve needs n >= 1");
    k_tridiagonal<<<1, 256, 0, ctx->stream>>>(a, b, c, d, x, n);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_tridiagonal launch");
}

__global__ void k_prepare_charge(float* __restrict__ ion, float* __restrict__ ele, float* __restrict__ charge,
                                 int n, int periodic, int fix_i, int fix_e) {
    // one CTA; the two end nodes first (ref infMagSim_script.py:764-769, 774-779)
    __shared__ float s_end[4];
    if (threadIdx.x == 0) {
        float i0 = ion[0], i1 = ion[n - 1], e0 = ele[0], e1 = ele[n - 1];
        if (periodic) {
            if (fix_i) { i0 = __fadd_rn(i0, i1); i1 = i0; }
            if (fix_e) { e0 = __fadd_rn(e0, e1); e1 = e0; }
        } else {
            if (fix_i) { i0 = __fmul_rn(i0, 2.f); i1 = __fmul_rn(i1, 2.f); }
            if (fix_e) { e0 = __fmul_rn(e0, 2.f); e1 = __fmul_rn(e1, 2.f); }
        }
        s_end[0] = i0; s_end[1] = i1; s_end[2] = e0; s_end[3] = e1;
        ion[0] = i0; ion[n - 1] = i1; ele[0] = e0; ele[n - 1] = e1;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float a = (j == 0) ? s_end[0] : (j == n - 1) ? s_end[1] : ion[j];
        float b = (j == 0) ? s_end[2] : (j == n - 1) ? s_end[3] : ele[j];
        charge[j] = __fsub_rn(a, b);  // ref :807
    }
}

// CTAs in the solving cluster for n grid points: the smallest power of two (up to the portable
// cluster size) that gives every point its own thread, i.e. one row per thread up to 8192 points.
static int team_size_for(int n) {
    int c = 1;
    while (c < kMaxTeam && c * kSolveThreads < n) c *= 2;
    return c;
}

static int launch_poisson_impl(espic_ctx* ctx, SolveArgs& a) {
    const int n = a.n;
    if (n < 3) return fail(ctx, ESPIC_E_ARG, "poisson_solve needs n_points >= 3");
    int& max_team = ctx->solve_max_team;
    if (!ctx->solve_configured) {  // static (19 KB) + dynamic (up to 128 KB) exceeds the 48 KB default; per context/device
        max_team = kMaxTeam;
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_poisson<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSolveSmemRows * kSolveThreads * 8));
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_poisson<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSolveSmemRows * kSolveThreads * 8));
        if (kMaxTeam > 8) {
            ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_poisson<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_poisson<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        if (const char* e = getenv("ESPIC_SOLVE_TEAM")) {  // tuning knob: largest cluster to use
            const int t = atoi(e);
            if (t >= 1 && t <= kMaxTeam && (t & (t - 1)) == 0) max_team = t;
        }
        // the largest cluster this device can co-schedule with the biggest shared-memory footprint
        while (max_team > 1) {
            cudaLaunchConfig_t probe = {};
            probe.gridDim = dim3(max_team);
            probe.blockDim = dim3(kSolveThreads);
            probe.dynamicSmemBytes = kSolveSmemRows * kSolveThreads * 8;
            cudaLaunchAttribute pa[1];
            pa[0].id = cudaLaunchAttributeClusterDimension;
            pa[0].val.clusterDim.x = max_team;
            pa[0].val.clusterDim.y = 1;
            pa[0].val.clusterDim.z = 1;
            probe.attrs = pa;
            probe.numAttrs = 1;
            int clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&clusters, k_poisson<false>, &probe) == cudaSuccess && clusters >= 1) break;
            cudaGetLastError();
            max_team /= 2;
        }
        ctx->solve_configured = true;
    }
    int team = team_size_for(n);
    if (team > max_team) team = max_team;
    const int tn = team * kSolveThreads;
    const int m = (n + tn - 1) / tn;
    const size_t plane = (size_t)m * tn;
    int rc = ensure(ctx, ctx->solve_ws, (kHeaderDoubles + 10 * plane) * sizeof(double), "solve workspace");
    if (rc) return rc;
    a.ws = (double*)ctx->solve_ws.ptr;
    a.status = ctx->status;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(team);
    cfg.blockDim = dim3(kSolveThreads);
    cfg.dynamicSmemBytes = (m <= kSolveSmemRows) ? (size_t)m * kSolveThreads * sizeof(double) : 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (m == 1) ESPIC_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_poisson<true>, a));
    else ESPIC_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_poisson<false>, a));
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_poisson launch");
}

int launch_poisson(espic_ctx* ctx, const float* grid, const float* mask, const float* charge, float debye,
                   float* potential, float obj_pot, float transparency, int boltzmann, int periodic,
                   int n_points) {
    SolveArgs a{};
    a.grid = grid;
    a.mask = mask;
    a.charge = charge;
    a.potential = potential;
    a.debye = debye;
    a.obj_pot = obj_pot;
    a.transparency = transparency;
    a.boltzmann = boltzmann;
    a.periodic = periodic;
    a.n = n_points;
    return launch_poisson_impl(ctx, a);
}

int launch_field_solve(espic_ctx* ctx, const float* grid, const float* mask, float* ion, float* ele,
                       float* charge_out, int periodic_particles, int fix_i, int fix_e, float debye,
                       float* potential, float obj_pot, float transparency, int boltzmann, int periodic_potential,
                       int n_points, unsigned long long* acc_i, unsigned long long* acc_e, float bg,
                       const StackFix* fix) {
    SolveArgs a{};
    a.grid = grid;
    a.mask = mask;
    a.ion = ion;
    a.ele = ele;
    a.charge_out = charge_out;
    a.potential = potential;
    a.debye = debye;
    a.obj_pot = obj_pot;
    a.transparency = transparency;
    a.boltzmann = boltzmann;
    a.periodic = periodic_potential;
    a.fix_i = fix_i;
    a.fix_e = fix_e;
    a.n = n_points;
    a.fold_ends = periodic_particles;  // the driver ties the two flags (infMagSim_script.py:146); they are independent here
    a.acc_i = acc_i;
    a.acc_e = acc_e;
    a.inv_scale = 1.0 / ((double)kWeightOne * (double)bg);
    if (fix) a.fix = *fix;
    return launch_poisson_impl(ctx, a);
}

// ---- multi-GPU exchange buffers ---------------------------------------------------------------
// layout of one rank's buffer: [parity 0 slot][parity 1 slot][flags: one 128-byte line per parity]
static size_t xch_bytes(const espic_ctx::Exchange& x) { return 2 * x.slot_floats * sizeof(float) + 2 * 128; }
static unsigned* xch_flag(void* base, const espic_ctx::Exchange& x, int parity) {
    return reinterpret_cast<unsigned*>(static_cast<char*>(base) + 2 * x.slot_floats * sizeof(float) + parity * 128);
}
static float* xch_slot(void* base, const espic_ctx::Exchange& x, int parity) {
    return static_cast<float*>(base) + (size_t)parity * x.slot_floats;
}

int exchange_create(espic_ctx* ctx, int n_points, int world, int rank, void* handle_out) {
    if (world < 1 || world > ESPIC_MAX_RANKS || rank < 0 || rank >= world || n_points < 3 || !handle_out)
        return fail(ctx, ESPIC_E_ARG, "espic_exchange_create: bad arguments (world %d rank %d)", world, rank);
    if (ctx->xch.local) return fail(ctx, ESPIC_E_ARG, "espic_exchange_create: already created");
    auto& x = ctx->xch;
    x.world = world;
    x.rank = rank;
    x.n_points = n_points;
    x.slot_floats = 2 * (size_t)((n_points + 31) & ~31);
    ESPIC_CUDA(ctx, cudaMalloc(&x.local, xch_bytes(x)));
    ESPIC_CUDA(ctx, cudaMemset(x.local, 0, xch_bytes(x)));
    // counters start "one step in the past" so that step 0 has to be published like any other
    const unsigned minus_one = 0xffffffffu;
    for (int p = 0; p < 2; ++p)
        ESPIC_CUDA(ctx, cudaMemcpy(xch_flag(x.local, x, p), &minus_one, sizeof(unsigned), cudaMemcpyHostToDevice));
    cudaIpcMemHandle_t h;
    ESPIC_CUDA(ctx, cudaIpcGetMemHandle(&h, x.local));
    static_assert(sizeof(cudaIpcMemHandle_t) == ESPIC_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    x.peer[rank] = x.local;
    return 0;
}

int exchange_open(espic_ctx* ctx, const void* handles) {
    auto& x = ctx->xch;
    if (!x.local || !handles) return fail(ctx, ESPIC_E_ARG, "espic_exchange_open: create first");
    for (int r = 0; r < x.world; ++r) {
        if (r == x.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * ESPIC_IPC_HANDLE_BYTES, sizeof(h));
        ESPIC_CUDA(ctx, cudaIpcOpenMemHandle(&x.peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    x.opened = true;
    return 0;
}

// One CTA.  Fused deposit: the rows still sitting in the movers' fixed-point accumulators are converted
// into this rank's exchange slot first (and the accumulators cleared for the next pushes).
__global__ void __launch_bounds__(1024) k_publish(unsigned* flag, unsigned value, unsigned long long* acc_i,
                                                  unsigned long long* acc_e, float* slot, int n, int n_pad,
                                                  double inv_scale) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        if (acc_i) {
            slot[j] = (float)((double)acc_i[j] * inv_scale);
            acc_i[j] = 0ull;
        }
        if (acc_e) {
            slot[n_pad + j] = (float)((double)acc_e[j] * inv_scale);
            acc_e[j] = 0ull;
        }
    }
    // everything enqueued before this kernel on the stream (pushes, injections) has completed;
    // make it -- and the rows written above -- visible system-wide, then publish
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

int exchange_publish(espic_ctx* ctx, int parity, int step, unsigned long long* acc_i, unsigned long long* acc_e,
                     float bg) {
    auto& x = ctx->xch;
    if (!x.local) return fail(ctx, ESPIC_E_ARG, "espic_exchange_publish: no exchange");
    const bool convert = acc_i || acc_e;
    k_publish<<<1, convert ? 1024 : 32, 0, ctx->stream>>>(xch_flag(x.local, x, parity & 1), (unsigned)step, acc_i, acc_e,
                                                       xch_slot(x.local, x, parity & 1), convert ? x.n_points : 0,
                                                       (int)(x.slot_floats / 2),
                                                       convert ? 1.0 / ((double)kWeightOne * (double)bg) : 0.);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_publish launch");
}

int launch_field_solve_exchange(espic_ctx* ctx, int parity, int step, const float* grid, const float* mask,
                                float* ion, float* ele, float* charge_out, int periodic_particles, int fix_i,
                                int fix_e, float debye, float* potential, float obj_pot, float transparency,
                                int boltzmann, int periodic_potential, int n_points, const StackFix* fix) {
    auto& x = ctx->xch;
    if (!x.opened && x.world > 1) return fail(ctx, ESPIC_E_ARG, "espic_field_solve_exchange: exchange not opened");
    if (!x.local || n_points != x.n_points) return fail(ctx, ESPIC_E_ARG, "espic_field_solve_exchange: n_points mismatch");
    SolveArgs a{};
    a.grid = grid;
    a.mask = mask;
    a.ion = ion;
    a.ele = ele;
    a.charge_out = charge_out;
    a.potential = potential;
    a.debye = debye;
    a.obj_pot = obj_pot;
    a.transparency = transparency;
    a.boltzmann = boltzmann;
    a.periodic = periodic_potential;
    a.fix_i = fix_i;
    a.fix_e = fix_e;
    a.n = n_points;
    a.fold_ends = periodic_particles;
    a.world = x.world;
    if (fix) a.fix = *fix;
    a.expect = (unsigned)step;
    a.timeout_ns = (unsigned long long)(ctx->exchange_timeout_s * 1e9);
    a.n_pad = (int)(x.slot_floats / 2);
    for (int r = 0; r < x.world; ++r) {
        a.peer_slot[r] = xch_slot(x.peer[r], x, parity & 1);
        a.peer_flag[r] = xch_flag(x.peer[r], x, parity & 1);
    }
    if (x.world == 1) {  // single rank: the slot is the density itself
        ESPIC_CUDA(ctx, cudaMemcpyAsync(ion, a.peer_slot[0], (size_t)n_points * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        ESPIC_CUDA(ctx, cudaMemcpyAsync(ele, a.peer_slot[0] + a.n_pad, (size_t)n_points * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return launch_poisson_impl(ctx, a);
}

int launch_prepare_charge(espic_ctx* ctx, float* ion, float* ele, float* charge, int n, int periodic,
                          int fix_i, int fix_e) {
    if (n < 2) return fail(ctx, ESPIC_E_ARG, "n_points must be >= 2");
    k_prepare_charge<<<1, 1024, 0, ctx->stream>>>(ion, ele, charge, n, periodic, fix_i, fix_e);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_prepare_charge launch");
}

}  // namespace espic
