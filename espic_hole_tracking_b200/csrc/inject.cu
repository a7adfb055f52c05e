// K3: wall injection.  Replaces the body of inject_particles (ref infMagSim_cython.py:230-285)
// and provides the device-side source of its two random inputs: the flux-weighted drifting
// Maxwellian rejection sampler draw_velocities_c (ref infMagSim_c.c:560-666) and the uniform
// partial-step fractions the driver's host sampler supplies (ref infMagSim_script.py:321).
//
// The reference pops one slot per injected particle from the LIFO stack, in order.  Here
// particle l takes stack entry top-l directly, which is the same assignment as long as no
// particle has a "bad left_index" (ref :272-274: such a particle keeps its slot on the
// stack and the next one overwrites it).  That case needs |partial_dt*v| larger than the
// whole domain and is unreachable with physical inputs, but it is still honoured: a first
// kernel counts bad particles and, if there are any, the commit kernel replays the
// reference's sequential loop on one thread instead of the parallel path.
//
// Random numbers: the reference's generator is a process-global, strictly sequential MT19937
// whose consumption per particle depends on the rejection loop; its stream cannot be split
// across threads.  The device sampler therefore uses the counter-based Philox4x32-10
// generator (one independent counter per output element) with the same float conversion
// (ref :101) and the same acceptance test in double (ref :614, :654).  Parity of the
// injection body itself is exact when both sides are fed the same velocities and fractions
// (tests/test_gpu_parity.py::test_injection_given_inputs, tests/test_golden.py); parity of the sampler
// is statistical.
#include <cmath>

#include "espic_internal.cuh"

namespace espic {

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------------
struct Philox {
    uint32_t key[2];
    __device__ __forceinline__ static void round(uint32_t c[4], const uint32_t k[2]) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t out[4]) const {
        uint32_t c[4] = {c0, c1, c2, c3};
        uint32_t k[2] = {key[0], key[1]};
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// ref infMagSim_c.c:101: (float)y / (unsigned long)0xffffffff -- the divisor becomes the
// float 2^32, and the quotient may round up to exactly 1.0f.
__device__ __forceinline__ float u32_to_unit_float(uint32_t y) { return (float)y / 4294967296.0f; }

__global__ void k_uniforms(int n, float* __restrict__ out, unsigned long long seed, unsigned long long call) {
    Philox ph{{(uint32_t)seed, (uint32_t)(seed >> 32)}};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; 4 * i < n; i += gridDim.x * blockDim.x) {
        uint32_t r[4];
        ph((uint32_t)i, 0u, (uint32_t)call, (uint32_t)(call >> 32) ^ 0x554e4946u, r);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (4 * i + c < n) out[4 * i + c] = u32_to_unit_float(r[c]);
    }
}

// One thread per velocity; same side choice, window and acceptance rule as the reference
// (ref infMagSim_c.c:576-578, 584, 608-614, 648-654).  The reference evaluates the acceptance
// test log(u) <= log(|v|/|v_max|) + ((v_max+v_d)^2 - (v+v_d)^2)/(2 v_th^2) in double; the stream of
// uniforms differs from the reference's anyway (see the header), so the test is evaluated in
// float with the fast log: the accepted distribution changes by O(1e-6) relative.
// The constants of one draw_velocities call (ref infMagSim_c.c:576-578, 608-610, 648-650), worked
// out once on the host in double as the reference does: in the kernel they cost every thread a
// thousand double-precision instructions (pow, exp, two erf) -- 20 us of fixed latency per call.
struct SamplerConsts {
    float v_d, inv_two_vt2, side_split;
    float lo[2], width[2], mode_term[2], log_abs_mode[2];
};

static SamplerConsts sampler_consts(float v_th, float v_d) {
    SamplerConsts c;
    const float half_window = 5.f;
    const float beta = (float)((double)v_d / (sqrt(2.) * (double)v_th));
    const double b = (double)beta;
    const double gauss = 1. / sqrt(3.14159265358979323846) * exp(-pow(b, 2.));
    const float ratio = (float)((gauss - b * (1. - erf(b))) / (gauss + b * (1. + erf(b))));
    const double vd = (double)v_d, vt = (double)v_th;
    const double disc = sqrt(pow(vd, 2.) + 4. * pow(vt, 2.));
    c.v_d = v_d;
    c.inv_two_vt2 = (float)(1. / (2. * pow(vt, 2.)));
    c.side_split = 1.f / (1.f + ratio);
    float mode[2];
    mode[0] = (float)((-vd + disc) / 2.);   // enters at z_min moving right (ref :608-610)
    c.lo[0] = (float)fmax((double)(mode[0] - half_window * v_th), 0.);
    c.width[0] = (mode[0] + half_window * v_th) - c.lo[0];
    mode[1] = (float)((-vd - disc) / 2.);   // enters at z_max moving left (ref :648-650)
    c.lo[1] = mode[1] - half_window * v_th;
    c.width[1] = (float)fmin((double)(mode[1] + half_window * v_th), 0.) - c.lo[1];
    for (int s = 0; s < 2; ++s) {
        c.mode_term[s] = (mode[s] + v_d) * (mode[s] + v_d);
        c.log_abs_mode[s] = logf(fabsf(mode[s]));
    }
    return c;
}

// One velocity (element i of sampler call `call`).
__device__ __forceinline__ float draw_velocity(const Philox& ph, int i, unsigned long long call, const SamplerConsts& k,
                                               int* status) {
    const float v_d = k.v_d, inv_two_vt2 = k.inv_two_vt2, side_split = k.side_split;
    uint32_t r[4];
    uint32_t block = 0;
    ph((uint32_t)i, block++, (uint32_t)call, (uint32_t)(call >> 32) ^ 0x56454c4fu, r);
    const int s = (u32_to_unit_float(r[0]) - side_split >= 0.f) ? 0 : 1;   // ref :584-585
    const float lo_s = s ? k.lo[1] : k.lo[0], width_s = s ? k.width[1] : k.width[0];
    const float mode_term_s = s ? k.mode_term[1] : k.mode_term[0];
    const float log_abs_mode_s = s ? k.log_abs_mode[1] : k.log_abs_mode[0];
    int next = 2;  // first trial uses r[2], r[3] of block 0; later blocks give two trials each
    float v = 0.f;
    for (int tries = 1;; ++tries) {
        if (next >= 4) {
            ph((uint32_t)i, block++, (uint32_t)call, (uint32_t)(call >> 32) ^ 0x56454c4fu, r);
            next = 0;
        }
        const float u_v = u32_to_unit_float(r[next]);
        const float u_a = u32_to_unit_float(r[next + 1]);
        next += 2;
        v = __fadd_rn(__fmul_rn(u_v, width_s), lo_s);  // ref :611-612 / :651-652
        const float vs = v + v_d;
        const float bound = (__logf(fabsf(v)) - log_abs_mode_s) + (mode_term_s - vs * vs) * inv_two_vt2;
        if (__logf(u_a) <= bound) break;                 // ref :614 / :654
        // ref :589-607: after 1000 rejections the reference prints a warning and keeps
        // looping; we raise a status bit, and give up after 1e6 so a NaN parameter cannot
        // hang the GPU
        if (tries == 1000) atomicOr(status, ESPIC_ST_SAMPLER_STALL);
        if (tries >= 1000000) break;
    }
    return v;
}

// Element i of uniform call `call`: the same value k_uniforms writes to out[i].
__device__ __forceinline__ float uniform_at(const Philox& ph, int i, unsigned long long call) {
    uint32_t r[4];
    ph((uint32_t)(i >> 2), 0u, (uint32_t)call, (uint32_t)(call >> 32) ^ 0x554e4946u, r);
    return u32_to_unit_float(r[i & 3]);
}

__global__ void k_draw_velocities(int n, const SamplerConsts k, float* __restrict__ out, unsigned long long seed,
                                  unsigned long long call, int* status) {
    Philox ph{{(uint32_t)seed, (uint32_t)(seed >> 32)}};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = draw_velocity(ph, i, call, k, status);
}

// ---- injection body --------------------------------------------------------------------------
struct InjectArgs {
    const float* grid;
    const float* vel;
    const float* uni;
    float* px;
    float* pv;
    int* hist;
    int* empty_slots;
    int* top;
    int* largest;
    unsigned long long* acc;
    float* density;
    int* status;
    int* ws;  // [0] bad count, [1] stack top after a sequential replay, [2] a deposit went past the wall tiles
    unsigned* ticket;  // last-CTA-done counter of the commit kernel
    // Slot assignment by wall (sorted stores, see k_inject_commit): left_count[c] = particles of chunk c (of
    // kSampleThreads consecutive particles) that enter at the left wall, written by k_sample_classify
    int* left_count;
    int by_side;
    // Replayed step graphs (espic_pic_run): what changes from step to step is read from a device control block
    // instead of the argument block -- ctl[0] = step index (the tag), ctl[1] = index of the step in the run's count
    // arrays, ctl[2..3] = the sampler's call counter (64 bit), counts of species s at ctl[16 + s*ctl_stride + i].
    // n_inject is then the upper bound the launch was sized for.  The commit kernel of the species injected
    // last (ctl_advance) moves the block on to the next step.
    int* ctl;
    int ctl_species, ctl_stride, ctl_advance;
    int n_inject, n_points, k;
    float dt, a_b;
    double inv_scale;
};

struct InjectStep {
    int n_inject, k;
    unsigned long long call_uni, call_vel;
};
// The per-step values: from the argument block, or from the control block of a replayed graph.  A species with
// nothing to inject in a step draws nothing (the stepwise path skips its two sampler calls), so the electrons'
// calls come after the ions' only if there were ions.
__device__ __forceinline__ InjectStep inject_step(const InjectArgs& a, unsigned long long call_uni, unsigned long long call_vel) {
    InjectStep st{a.n_inject, a.k, call_uni, call_vel};
    if (a.ctl) {
        const int i = __ldcg(a.ctl + 1);
        st.k = __ldcg(a.ctl);
        st.n_inject = min(__ldcg(a.ctl + 16 + a.ctl_species * a.ctl_stride + i), a.n_inject);
        unsigned long long base = __ldcg(reinterpret_cast<const unsigned long long*>(a.ctl + 2));
        if (a.ctl_species > 0 && __ldcg(a.ctl + 16 + i) > 0) base += 2ull;
        st.call_uni = base;
        st.call_vel = base + 1ull;
    }
    return st;
}

struct Injected {
    float x, v;
    int cell;
    bool bad;
};

// ref infMagSim_cython.py:246, 262-271 -- float arithmetic throughout
__device__ __forceinline__ Injected place_injected(const Geom& g, float vel, float u, float dt, float a_b) {
    Injected r;
    const float edge = 2.e-6f;  // ref :238
    const float part_dt = __fmul_rn(dt, u);
    r.v = __fsub_rn(vel, __fmul_rn(a_b, part_dt));
    const float wall = (vel < 0.f) ? __fsub_rn(g.z_hi, edge) : __fadd_rn(g.z_lo, edge);
    r.x = __fadd_rn(wall, __fmul_rn(part_dt, r.v));
    r.cell = (int)__fdiv_rn(__fsub_rn(r.x, g.z_lo), g.dz);
    r.bad = (r.cell < 0) || (r.cell > g.last_cell);
    return r;
}

__global__ void k_inject_classify(const InjectArgs a) {
    const Geom g = make_geom(a.grid, a.n_points);
    int bad = 0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < a.n_inject; l += gridDim.x * blockDim.x)
        bad += place_injected(g, a.vel[l], a.uni[l], a.dt, a.a_b).bad ? 1 : 0;
    bad = __syncthreads_count(bad);
    if (threadIdx.x == 0 && bad) atomicAdd(a.ws, bad);
}

// The device sampler's two draws and the classification above in one launch: fractions and
// velocities are written out for the commit kernel, the bad-index count goes to ws[0].
constexpr int kSampleThreads = 128;

__device__ __forceinline__ void sample_classify_body(const InjectArgs& a, const SamplerConsts& k, float* __restrict__ uni_out,
                                                     float* __restrict__ vel_out, unsigned long long seed,
                                                     unsigned long long call_uni, unsigned long long call_vel, int bid, int nb) {
    const Geom g = make_geom(a.grid, a.n_points);
    Philox ph{{(uint32_t)seed, (uint32_t)(seed >> 32)}};
    int bad = 0;
    const InjectStep st = inject_step(a, call_uni, call_vel);
    call_uni = st.call_uni;
    call_vel = st.call_vel;
    const int n_chunks = (st.n_inject + kSampleThreads - 1) / kSampleThreads;
    for (int chunk = bid; chunk < n_chunks; chunk += nb) {  // uniform trip count per CTA
        const int l = chunk * kSampleThreads + threadIdx.x;
        int left = 0;
        if (l < st.n_inject) {
            const float u = uniform_at(ph, l, call_uni);
            const float v = draw_velocity(ph, l, call_vel, k, a.status);
            uni_out[l] = u;
            vel_out[l] = v;
            bad += place_injected(g, v, u, a.dt, a.a_b).bad ? 1 : 0;
            left = !(v < 0.f);  // enters at z_min (ref infMagSim_cython.py:263-268)
        }
        left = __syncthreads_count(left);
        if (threadIdx.x == 0 && a.left_count) a.left_count[chunk] = left;
    }
    bad = __syncthreads_count(bad);
    if (threadIdx.x == 0 && bad) atomicAdd(a.ws, bad);
}

__global__ void __launch_bounds__(kSampleThreads) k_sample_classify(const InjectArgs a, const SamplerConsts k,
                                                                    float* __restrict__ uni_out, float* __restrict__ vel_out,
                                                                    unsigned long long seed, unsigned long long call_uni,
                                                                    unsigned long long call_vel) {
    sample_classify_body(a, k, uni_out, vel_out, seed, call_uni, call_vel, (int)blockIdx.x, (int)gridDim.x);
}

// Both species of a step in one launch: the first `blocks` CTAs of the grid work for job 0 (the ions), the rest for
// job 1.  Each job has its own outputs, counters and sampler calls -- exactly what two launches would produce.
struct SampleJob {
    InjectArgs a;
    SamplerConsts k;
    float* uni;
    float* vel;
    unsigned long long call_uni, call_vel;
    int blocks;
};

__global__ void __launch_bounds__(kSampleThreads) k_sample_classify2(const __grid_constant__ SampleJob j0,
                                                                     const __grid_constant__ SampleJob j1,
                                                                     unsigned long long seed) {
    if ((int)blockIdx.x < j0.blocks)
        sample_classify_body(j0.a, j0.k, j0.uni, j0.vel, seed, j0.call_uni, j0.call_vel, (int)blockIdx.x, j0.blocks);
    else
        sample_classify_body(j1.a, j1.k, j1.uni, j1.vel, seed, j1.call_uni, j1.call_vel, (int)blockIdx.x - j0.blocks, j1.blocks);
}

__device__ __forceinline__ void deposit_injected(const Geom& g, const Injected& p, unsigned long long* acc) {
    const unsigned w0 = left_share_fixed(g, p.x, p.cell);
    atomicAdd(acc + p.cell, (unsigned long long)w0);
    atomicAdd(acc + p.cell + 1, (unsigned long long)(kWeightOne - w0));
}

// Injected particles land within |v|*dt of a wall, i.e. on a few hundred nodes: depositing them
// with global atomics serialises on those addresses.  Each CTA accumulates into a shared-memory
// copy of the two wall regions first (u32 fixed point, carry to the global accumulator when a
// word would overflow) and flushes the touched nodes once.
constexpr int kWallNodes = 2048;  // nodes kept per wall; anything farther in goes to global memory

__device__ __forceinline__ void deposit_injected_tiled(const Geom& g, const Injected& p, unsigned* s_wall,
                                                       unsigned long long* acc, int n_points, bool& interior) {
    const unsigned w0 = left_share_fixed(g, p.x, p.cell);
    const unsigned w[2] = {w0, kWeightOne - w0};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int node = p.cell + k;
        int local = -1;
        if (node < kWallNodes) local = node;
        else if (node >= n_points - kWallNodes) local = kWallNodes + (node - (n_points - kWallNodes));
        if (local >= 0 && n_points >= 2 * kWallNodes) {
            const unsigned old = atomicAdd(s_wall + local, w[k]);
            if (old > ~w[k]) atomicAdd(acc + node, 1ull << 32);  // the word wrapped: credit 2^32 globally
        } else {
            atomicAdd(acc + node, (unsigned long long)w[k]);
            interior = true;
        }
    }
}

// Density of the injected particles added to the caller's float density, accumulators re-zeroed,
// stack top updated: run by the last CTA of the commit kernel to finish (ref infMagSim_cython.py:281-285).
__device__ __forceinline__ void inject_finish(const InjectArgs& a, bool everywhere, int n_inject) {
    const int n = a.n_points;
    // injected particles sit within |v| dt of a wall: unless a deposit went further in, only the
    // two wall regions can hold anything
    const int span = a.density ? (everywhere ? n : 2 * kWallNodes) : 0;  // fused step path: the deposit stays in the accumulators
    for (int jj = threadIdx.x; jj < span; jj += blockDim.x) {
        const int j = everywhere ? jj : (jj < kWallNodes ? jj : (n - kWallNodes) + (jj - kWallNodes));
        const unsigned long long v = __ldcg(a.acc + j);
        if (v) {
            a.density[j] = __fadd_rn(a.density[j], (float)((double)v * a.inv_scale));
            a.acc[j] = 0ull;
        }
    }
    if (threadIdx.x == 0) {
        const int top0 = *a.top;
        if (a.ws[0] == 0) {
            const int n_ok = min(n_inject, top0 + 1);
            *a.top = top0 - (n_ok > 0 ? n_ok : 0);
        } else {
            *a.top = a.ws[1];
        }
        a.ws[0] = 0;
        a.ws[1] = 0;
        a.ws[2] = 0;
        *a.ticket = 0u;
        if (a.ctl && a.ctl_advance) {  // every CTA of this launch has read the block: on to the next step
            const int i = a.ctl[1];
            unsigned long long* calls = reinterpret_cast<unsigned long long*>(a.ctl + 2);
            *calls += (a.ctl[16 + i] > 0 ? 2ull : 0ull) + (a.ctl[16 + a.ctl_stride + i] > 0 ? 2ull : 0ull);
            a.ctl[0] += 1;
            a.ctl[1] = i + 1;
        }
    }
}

// Slot assignment.  The reference pops one stack entry per injected particle, in order: particle l takes entry
// top-l (by_side == 0: bit-identical slots).  On a store kept sorted by position (large grids, csrc/sort.cu)
// which of the free slots a new particle takes is physically irrelevant but decides whether it lands in a
// segment of the store whose shared-memory window covers its wall.  With by_side != 0 the mover's k_scatter has
// put the slots it freed this step on the stack alternately from the two ends of their ascending order, outermost
// on top (highest slot, lowest slot, second highest, ...): particles that enter at the right wall take the
// entries at even distance from the top, those that enter at the left wall the entries at odd distance -- the
// slots next to their own wall -- and when more particles left than enter, what stays on the stack are the
// innermost freed slots of both wall regions.  The same top n entries the reference would pop, each used once,
// deterministically (ranks from a scan, no atomics).  Without this half of all injected electrons sat 65536
// cells away from their segment until the next sort and were served from global memory; with a one-sided
// assignment (left from the bottom of the popped range, right from the top) a shortfall of left-wall slots
// -- (removed - injected)/2 per step while the population falls -- still put left-wall entrants into
// right-wall slots, and the two segments that collected them ran 6x longer than the others.
__device__ __forceinline__ void inject_commit_body(const InjectArgs& a, const int bid, const int nb) {
    __shared__ unsigned s_wall[2 * kWallNodes];
    __shared__ int s_last;
    __shared__ int s_scan[8], s_scan2[8];
    __shared__ int s_base_left, s_total_left;
    const Geom g = make_geom(a.grid, a.n_points);
    const int top0 = *a.top;
    const int n_bad = a.ws[0];
    const InjectStep st = inject_step(a, 0ull, 0ull);
    const int n_inject = st.n_inject;
    if (n_bad == 0) {
        for (int j = threadIdx.x; j < 2 * kWallNodes; j += blockDim.x) s_wall[j] = 0u;
        __syncthreads();
        int my_max = -1;
        bool interior = false;
        // contiguous runs of 256-particle rounds per CTA, so that the left-wall rank of a round is the rank
        // of the previous one plus its count
        const int n_rounds = (n_inject + 255) / 256;
        const int per_cta = (n_rounds + nb - 1) / nb;
        const int r0 = min(bid * per_cta, n_rounds), r1 = min(r0 + per_cta, n_rounds);
        int base_left = 0, total_left = 0;
        if (a.by_side && r0 < r1) {
            // particles entering at the left wall in front of this CTA's first round, and in all (two sampler chunks per round)
            int part = 0, all = 0;
            for (int c = threadIdx.x; c < 2 * n_rounds; c += blockDim.x) {
                const int v = (c * kSampleThreads < n_inject) ? a.left_count[c] : 0;
                all += v;
                part += c < 2 * r0 ? v : 0;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                part += __shfl_xor_sync(0xffffffffu, part, d);
                all += __shfl_xor_sync(0xffffffffu, all, d);
            }
            if ((threadIdx.x & 31) == 0) {
                s_scan[threadIdx.x >> 5] = part;
                s_scan2[threadIdx.x >> 5] = all;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int t = 0, u = 0;
                for (int w = 0; w < 8; ++w) {
                    t += s_scan[w];
                    u += s_scan2[w];
                }
                s_base_left = t;
                s_total_left = u;
            }
            __syncthreads();
            base_left = s_base_left;
            total_left = s_total_left;
            __syncthreads();
        }
        // by side: the mover's k_scatter left the freed slots on the stack alternately from the two ends of their
        // ascending order, outermost on top (mover.cu, interleave): entrants at the right wall take the entries at even
        // distance from the top, entrants at the left wall those at odd distance, as long as both sides have
        // entrants left (m of each); the surplus of the larger side continues below.  Exactly the top n_inject
        // entries are consumed, each once.
        const int m_pairs = min(total_left, n_inject - total_left);
        for (int r = r0; r < r1; ++r) {
            const int l = r * 256 + (int)threadIdx.x;
            const bool have = l < n_inject;
            const float vel = have ? a.vel[l] : 0.f;
            int sp = top0 - l;  // the reference's pop order
            if (a.by_side) {
                const bool left = have && !(vel < 0.f);
                const unsigned lanes = __ballot_sync(0xffffffffu, left);
                const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
                if (lane == 0) s_scan[w] = __popc(lanes);
                __syncthreads();
                int before = 0, round_left = 0;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) {
                    before += ww < w ? s_scan[ww] : 0;
                    round_left += s_scan[ww];
                }
                const int rank_left = base_left + before + __popc(lanes & ((1u << lane) - 1u));
                const int rank = left ? rank_left : l - rank_left;  // rank among the entrants of its own wall
                const int t = rank < m_pairs ? 2 * rank + (left ? 1 : 0) : m_pairs + rank;  // distance from the top
                sp = top0 - t;
                base_left += round_left;
                __syncthreads();
            }
            if (!have) continue;
            if (sp < 0) {  // ref :255-256 "no empty slots" (the reference then reads empty_slots[-1])
                atomicOr(a.status, ESPIC_ST_STACK_EMPTY);
                continue;
            }
            const int slot = a.empty_slots[sp];
            const Injected p = place_injected(g, vel, a.uni[l], a.dt, a.a_b);
            a.hist[slot] = st.k;
            a.pv[slot] = p.v;
            a.px[slot] = p.x;
            my_max = max(my_max, slot);
            deposit_injected_tiled(g, p, s_wall, a.acc, a.n_points, interior);
            a.empty_slots[sp] = -1;
        }
        if (interior) a.ws[2] = 1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, d));
        if ((threadIdx.x & 31) == 0 && my_max >= 0) atomicMax(a.largest, my_max);
        __syncthreads();
        for (int j = threadIdx.x; j < 2 * kWallNodes; j += blockDim.x) {
            const unsigned v = s_wall[j];
            if (v) {
                const int node = (j < kWallNodes) ? j : (a.n_points - kWallNodes) + (j - kWallNodes);
                atomicAdd(a.acc + node, (unsigned long long)v);
            }
        }
    } else if (bid == 0 && threadIdx.x == 0) {
        // replay of the reference's sequential loop (only when a bad left_index occurred)
        int top = top0, last = *a.largest;
        for (int l = 0; l < n_inject; ++l) {
            if (top < 0) {
                atomicOr(a.status, ESPIC_ST_STACK_EMPTY);
                break;
            }
            const int slot = a.empty_slots[top];
            const Injected p = place_injected(g, a.vel[l], a.uni[l], a.dt, a.a_b);
            a.hist[slot] = st.k;
            a.pv[slot] = p.v;
            a.px[slot] = p.x;
            last = max(last, slot);
            if (p.bad) {
                atomicOr(a.status, ESPIC_ST_BAD_INJECT_INDEX);
            } else {
                deposit_injected(g, p, a.acc);
                a.empty_slots[top] = -1;
                top -= 1;
            }
        }
        *a.largest = last;
        a.ws[1] = top;  // final stack top for inject_finish
        a.ws[2] = 1;
    }
    // the last CTA to get here finishes for all of them (the stack top may only change once every
    // CTA has read it, the accumulators only once every CTA has flushed)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == (unsigned)nb - 1u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const bool everywhere = n_bad != 0 || a.n_points < 2 * kWallNodes || *(volatile int*)(a.ws + 2) != 0;
        inject_finish(a, everywhere, n_inject);
    }
}

__global__ void __launch_bounds__(256) k_inject_commit(const __grid_constant__ InjectArgs a) {
    inject_commit_body(a, (int)blockIdx.x, (int)gridDim.x);
}

// Both species of a step in one launch (see k_sample_classify2): the first blocks0 CTAs commit job 0, the rest job 1,
// each with its own store, counters and ticket.  In a replayed graph the last CTA of the whole grid to finish moves
// the control block on to the next step (every CTA has read it by then).
__global__ void __launch_bounds__(256) k_inject_commit2(const __grid_constant__ InjectArgs a0, const __grid_constant__ InjectArgs a1,
                                                        int blocks0, unsigned* joint_ticket) {
    const bool second = (int)blockIdx.x >= blocks0;
    const InjectArgs& a = second ? a1 : a0;   // grid constants: their addresses can be taken
    inject_commit_body(a, second ? (int)blockIdx.x - blocks0 : (int)blockIdx.x, second ? (int)gridDim.x - blocks0 : blocks0);
    if (a0.ctl) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(joint_ticket, 1u) == gridDim.x - 1) {
            int* ctl = a0.ctl;
            const int i = ctl[1];
            unsigned long long* calls = reinterpret_cast<unsigned long long*>(ctl + 2);
            *calls += (ctl[16 + i] > 0 ? 2ull : 0ull) + (ctl[16 + a0.ctl_stride + i] > 0 ? 2ull : 0ull);
            ctl[0] += 1;
            ctl[1] = i + 1;
            *joint_ticket = 0u;
        }
    }
}

static unsigned long long rng_key(espic_ctx* ctx);

// Injection workspace: [16 counters of job 0][16 counters of job 1][left-wall counts per sampler chunk, job 0][..., job 1].
// The counter blocks sit at FIXED offsets: they are state that must be zero between calls (a stale bad-index count
// sends the commit kernel down the reference's sequential replay), so no layout of any call may place its per-chunk
// counts on top of them.
static size_t chunk_ints(int n_inject) { return ((size_t)n_inject / kSampleThreads + 2 + 15) & ~(size_t)15; }

static int fill_inject_args(espic_ctx* ctx, InjectArgs& a, int n_inject, const float* grid, int n_points, float dt,
                            float bg, float a_b, int job = 0, size_t left_offset_ints = 0) {
    if (n_points < 2) return fail(ctx, ESPIC_E_ARG, "n_points must be >= 2");
    int rc = ensure_acc(ctx, n_points);
    if (rc) return rc;
    rc = ensure(ctx, ctx->inject_ws, (32 + left_offset_ints + chunk_ints(n_inject)) * sizeof(int), "inject workspace");
    if (rc) return rc;
    a.grid = grid;
    a.acc = ctx->acc;
    a.status = ctx->status;
    a.ws = (int*)ctx->inject_ws.ptr + 16 * job;
    a.left_count = (int*)ctx->inject_ws.ptr + 32 + left_offset_ints;
    a.ticket = (unsigned*)(ctx->status + (job ? 13 : 8));
    a.n_inject = n_inject;
    a.n_points = n_points;
    a.dt = dt;
    a.a_b = a_b;
    a.inv_scale = 1.0 / ((double)kWeightOne * (double)bg);
    return 0;
}

int launch_inject(espic_ctx* ctx, int n_inject, const float* grid, int n_points, float dt, float bg,
                  const float* vel, const float* uni, float a_b, int k, int* hist, float* particles,
                  int storage, int* empty_slots, int* top, int* largest, float* density, bool classified,
                  unsigned long long* acc_fused, int by_side, const InjectCtl* ctl) {
    if (n_inject <= 0) return 0;
    InjectArgs a{};
    if (ctl) {
        a.ctl = ctl->ptr;
        a.ctl_species = ctl->species;
        a.ctl_stride = ctl->stride;
        a.ctl_advance = ctl->advance;
    }
    int rc = fill_inject_args(ctx, a, n_inject, grid, n_points, dt, bg, a_b);
    if (rc) return rc;
    if (acc_fused) {  // step path: add to the caller's accumulators (converted by the next field solve)
        a.acc = acc_fused;
        density = nullptr;
    }
    a.vel = vel;
    a.uni = uni;
    a.px = particles;
    a.pv = particles + storage;
    a.hist = hist;
    a.empty_slots = empty_slots;
    a.top = top;
    a.largest = largest;
    a.density = density;
    a.k = k;
    // assignment by wall needs the per-chunk counts the device sampler leaves behind (rng_sample_for_injection)
    a.by_side = (by_side && classified) ? 1 : 0;
    int grid_dim = (n_inject + 255) / 256;
    if (grid_dim > 4 * ctx->sm_count) grid_dim = 4 * ctx->sm_count;
    if (!classified) {
        k_inject_classify<<<grid_dim, 256, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    // one CTA per SM at most: every CTA zeroes and flushes its own 16 KB wall tile; the last one to
    // finish adds the injected density to the caller's and pops the stack
    k_inject_commit<<<grid_dim < ctx->sm_count ? grid_dim : ctx->sm_count, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "inject launch");
}

// Fractions + velocities for n_inject particles from the device sampler and their classification, one
// launch; follow with launch_inject(..., classified = true).  Same streams as rng_uniforms followed by
// rng_draw_velocities (ref infMagSim_cython.py:245-248: fractions first).
int rng_sample_for_injection(espic_ctx* ctx, int n_inject, float v_th, float v_d, float* vel, float* uni,
                             const float* grid, int n_points, float dt, float bg, float a_b,
                             const InjectCtl* ctl) {
    if (n_inject <= 0) return 0;
    InjectArgs a{};
    int rc = fill_inject_args(ctx, a, n_inject, grid, n_points, dt, bg, a_b);
    if (rc) return rc;
    const unsigned long long key = rng_key(ctx);
    unsigned long long call_uni = 0ull, call_vel = 0ull;
    if (ctl) {  // the kernel takes both from the control block; the caller keeps the host's counter in step
        a.ctl = ctl->ptr;
        a.ctl_species = ctl->species;
        a.ctl_stride = ctl->stride;
    } else {
        call_uni = ctx->rng_consumed++;
        call_vel = ctx->rng_consumed++;
    }
    int grid_dim = (n_inject + kSampleThreads - 1) / kSampleThreads;
    if (grid_dim > 8 * ctx->sm_count) grid_dim = 8 * ctx->sm_count;
    k_sample_classify<<<grid_dim, kSampleThreads, 0, ctx->stream>>>(a, sampler_consts(v_th, v_d), uni, vel, key, call_uni, call_vel);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_sample_classify launch");
}

// Sampling and commit of BOTH species of a step in two launches instead of four (espic_pic_step with the device
// sampler, both counts positive): same values, same slots, same deposit as the per-species calls -- each species
// keeps its own random inputs, counters and ticket; only the launches are shared.
int launch_inject_pair(espic_ctx* ctx, const InjectPairSpecies sp[2], const float* grid, int n_points, float dt, float bg,
                       float a_b, int k, int by_side, int* ctl, int ctl_stride) {
    const unsigned long long key = rng_key(ctx);
    // workspace: [species 0 block][species 1 block]; random inputs: [uni0 | vel0 | uni1 | vel1]
    const size_t left1 = chunk_ints(sp[0].n_inject);   // job 1's per-chunk counts sit behind job 0's
    size_t pad[2], off = 0;
    for (int s = 0; s < 2; ++s) pad[s] = ((size_t)sp[s].n_inject + 63) & ~(size_t)63;
    int rc = ensure(ctx, ctx->rng_ws, 2 * (pad[0] + pad[1]) * sizeof(float), "injection random inputs");
    if (rc) return rc;
    SampleJob job[2];
    InjectArgs ca[2];
    for (int s = 0; s < 2; ++s) {
        InjectArgs a{};
        rc = fill_inject_args(ctx, a, sp[s].n_inject, grid, n_points, dt, bg, a_b, s, s ? left1 : 0);
        if (rc) return rc;
        if (ctl) {
            a.ctl = ctl;
            a.ctl_species = s;
            a.ctl_stride = ctl_stride;
            a.ctl_advance = 0;   // the joint ticket of k_inject_commit2 advances the block
        }
        float* uni = (float*)ctx->rng_ws.ptr + off;
        float* vel = uni + pad[s];
        off += 2 * pad[s];
        job[s].a = a;
        job[s].k = sampler_consts(sp[s].v_th, sp[s].v_d);
        job[s].uni = uni;
        job[s].vel = vel;
        job[s].call_uni = job[s].call_vel = 0ull;
        if (!ctl) {  // same order as two calls: ions' fractions, ions' velocities, electrons' fractions, electrons' velocities
            job[s].call_uni = ctx->rng_consumed++;
            job[s].call_vel = ctx->rng_consumed++;
        }
        int blocks = (sp[s].n_inject + kSampleThreads - 1) / kSampleThreads;
        if (blocks > 8 * ctx->sm_count) blocks = 8 * ctx->sm_count;
        job[s].blocks = blocks;
        // commit arguments
        a.vel = vel;
        a.uni = uni;
        a.px = sp[s].particles;
        a.pv = sp[s].particles + sp[s].storage;
        a.hist = sp[s].hist;
        a.empty_slots = sp[s].empty_slots;
        a.top = sp[s].top;
        a.largest = sp[s].largest;
        a.density = sp[s].acc_fused ? nullptr : sp[s].density;
        if (sp[s].acc_fused) a.acc = sp[s].acc_fused;
        a.k = k;
        a.by_side = by_side ? 1 : 0;
        ca[s] = a;
    }
    // the two species' workspace blocks are fixed only now (the second fill may have grown the buffer)
    for (int s = 0; s < 2; ++s) {
        int* base = (int*)ctx->inject_ws.ptr;
        job[s].a.ws = ca[s].ws = base + 16 * s;
        job[s].a.left_count = ca[s].left_count = base + 32 + (s ? left1 : 0);
    }
    k_sample_classify2<<<job[0].blocks + job[1].blocks, kSampleThreads, 0, ctx->stream>>>(job[0], job[1], key);
    ctx->launches++;
    int cb[2];
    for (int s = 0; s < 2; ++s) {
        int g = (sp[s].n_inject + 255) / 256;
        cb[s] = g < ctx->sm_count ? g : ctx->sm_count;
    }
    k_inject_commit2<<<cb[0] + cb[1], 256, 0, ctx->stream>>>(ca[0], ca[1], cb[0], (unsigned*)(ctx->status + 14));
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "paired inject launch");
}

// ---- sampler front ends -----------------------------------------------------------------------
int rng_seed(espic_ctx* ctx, unsigned long seed) {
    ctx->rng_seeded = 1;
    ctx->rng_consumed = 0;
    // spread the (typically small) integer seed over the 64-bit Philox key
    unsigned long long z = (unsigned long long)seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    ctx->rng_key = z;
    return 0;
}

static unsigned long long rng_key(espic_ctx* ctx) {
    if (!ctx->rng_seeded) rng_seed(ctx, 4357ul);  // the reference's default seed (ref infMagSim_c.c:78-79)
    return ctx->rng_key;
}

int rng_draw_velocities(espic_ctx* ctx, int n, float v_th, float v_d, float* v_out) {
    if (n <= 0) return 0;
    const unsigned long long key = rng_key(ctx);
    int grid_dim = (n + 127) / 128;
    if (grid_dim > 8 * ctx->sm_count) grid_dim = 8 * ctx->sm_count;
    k_draw_velocities<<<grid_dim, 128, 0, ctx->stream>>>(n, sampler_consts(v_th, v_d), v_out, key, ctx->rng_consumed++,
                                                         ctx->status);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_draw_velocities launch");
}

int rng_uniforms(espic_ctx* ctx, int n, float* out) {
    if (n <= 0) return 0;
    const unsigned long long key = rng_key(ctx);
    int grid_dim = (n / 4 + 255) / 256 + 1;
    if (grid_dim > 8 * ctx->sm_count) grid_dim = 8 * ctx->sm_count;
    k_uniforms<<<grid_dim, 256, 0, ctx->stream>>>(n, out, key, ctx->rng_consumed++);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_uniforms launch");
}

}  // namespace espic
