// Internal declarations shared by the translation units of libespic_b200.so.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true
//        -prec-sqrt=true -ftz=false.  -fmad=false matters: the reference is compiled for
//        baseline x86-64 (no FMA contraction), so "v += a*dt" is a rounded multiply followed
//        by a rounded add; wherever a fused operation is wanted it is written as fmaf().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/espic_b200.h"

namespace espic {

constexpr int kWeightBits = 24;  // deposit weights are fixed point with 24 fractional bits
constexpr unsigned kWeightOne = 1u << kWeightBits;

// Geometry every kernel derives from the grid array the way the reference does
// (ref infMagSim_c.c:126-130).
struct Geom {
    float z_lo, z_hi, dz, span;
    float parked;    // inactive-slot marker 2*z_max
    float live_thr;  // smallest float >= 0.99*parked: (double)x < 0.99*parked  <=>  x < live_thr
    float lo_edge, hi_edge;  // z_lo + 1e-6f, z_hi - 1e-6f
    float rcp_dz;    // RN(1/dz), for the correctly rounded quotient below
    float half_span, near_span;  // span/2 and 1.49*span: one-compare test for "within a period"
    float q_shift;   // trunc(z_lo/dz): cell index -> quotient estimate of x/dz (deposit weights)
    float rcp_dz24;  // rcp_dz * 2^24 (exact scaling): remainder -> 24-bit fixed-point share
    float fold_lo, fold_hi;  // periodic fast path: x1 - z_lo must lie strictly inside (fold_lo, fold_hi)
    float cell_span;  // smallest offset b <= span with (int)(b/dz) > n_points-2: offsets below it index a real cell
    int last_cell;   // n_points - 2
    bool fast_div;   // the 3-operation quotient is safe for this dz (normal exponent range)
};

__device__ __forceinline__ Geom make_geom(const float* __restrict__ grid, int n_points) {
    Geom g;
    g.z_lo = __ldg(grid);
    g.z_hi = __ldg(grid + n_points - 1);
    g.dz = __fdiv_rn(__fsub_rn(g.z_hi, g.z_lo), (float)(n_points - 1));
    g.span = __fsub_rn(g.z_hi, g.z_lo);
    g.parked = (float)(2. * (double)g.z_hi);
    g.live_thr = __double2float_ru(0.99 * (double)g.parked);
    g.lo_edge = __fadd_rn(g.z_lo, 1.e-6f);
    g.hi_edge = __fsub_rn(g.z_hi, 1.e-6f);
    g.rcp_dz = __frcp_rn(g.dz);
    g.half_span = 0.5f * g.span;
    g.near_span = 1.49f * g.span;
    g.q_shift = truncf(g.z_lo * g.rcp_dz);
    g.rcp_dz24 = g.rcp_dz * 16777216.0f;
    // within one period either side of the domain AND safely below the parked threshold
    // (a = fl(x - z_lo) is within half an ulp of x - z_lo; 1e-3*span is ample slack)
    g.fold_lo = -0.99f * g.span;
    g.fold_hi = fminf(1.99f * g.span, (g.live_thr - g.z_lo) - 1e-3f * g.span);
    g.last_cell = n_points - 2;
    // dz is a rounded quotient, so the last float(s) below span can already divide to n_cells
    // (e.g. 37 cells on [-3, 3]); the reference then takes cell 0 (periodic) or reports a bad index
    g.cell_span = g.span;
    for (int k = 0; k < 4; ++k) {
        const float below = __uint_as_float(__float_as_uint(g.cell_span) - 1u);
        if (!(below > 0.f) || __float2int_rz(__fdiv_rn(below, g.dz)) <= g.last_cell) break;
        g.cell_span = below;
    }
    // exponent window in which neither dz, 1/dz, the quotient nor the residual can leave the
    // normal range for |x - z_lo| <= span
    g.fast_div = (g.dz > 1e-18f) && (g.dz < 1e18f) && (g.span < 1e18f) &&
                 (g.live_thr - g.z_lo > 1.002f * g.span);  // parked marker well outside the domain
    return g;
}

static __device__ __noinline__ float fmod_slow(float a, float b) { return fmodf(a, b); }

// RN(a/dz).  FAST: bit-identical to IEEE division in three operations (Markstein):
// q0 = RN(a*y), r = a - q0*dz exactly (one FMA), q = RN(q0 + r*y), with y = RN(1/dz).
// Correct whenever no intermediate leaves the normal range: kernels take the FAST
// instantiation only when Geom::fast_div holds, and every call site passes |a| <= span
// (positions are folded or bounds-checked first).  For a below the normal range the last place
// may differ, which cannot change the truncation (0).  espic_selftest_quotient() checks this
// exhaustively on the device against the division instruction for a given grid.
template <bool FAST>
__device__ __forceinline__ float quotient_dz(const Geom& g, float a) {
    if (!FAST) return __fdiv_rn(a, g.dz);
    const float q0 = __fmul_rn(a, g.rcp_dz);
    const float r = fmaf(-q0, g.dz, a);
    return fmaf(r, g.rcp_dz, q0);
}

// fmodf(x - z_lo, span) followed by the reference's re-basing (ref infMagSim_c.c:148-153).
// Within one period either side the remainder is the dividend itself or one exact
// (Sterbenz) subtraction; anything farther out (or NaN) takes the library routine, exact too.
__device__ __forceinline__ float fold_periodic(const Geom& g, float x) {
    const float a = __fsub_rn(x, g.z_lo);
    float off = __fsub_rn(a, (a >= g.span) ? g.span : 0.f);
    if (!(fabsf(a - g.half_span) < g.near_span)) off = fmod_slow(a, g.span);  // rare
    return __fadd_rn(off, (off >= 0.f) ? g.z_lo : g.z_hi);
}

// (int)((x - z_lo)/dz) with IEEE division and truncation (ref infMagSim_c.c:161-163), clamped
// into [0, n_points-2] (the reference prints "bad left_index" and then indexes out of bounds);
// `bad` accumulates whether clamping changed anything.
template <bool PERIODIC, bool FAST>
__device__ __forceinline__ int cell_of(const Geom& g, float x, bool& bad) {
    int c = __float2int_rz(quotient_dz<FAST>(g, __fsub_rn(x, g.z_lo)));
    if (PERIODIC) c = (c > g.last_cell) ? 0 : c;
    const int cc = min(max(c, 0), g.last_cell);
    bad |= (cc != c);
    return cc;
}

// Share of the particle given to its LEFT node (ref infMagSim_c.c:219-223): fmodf(x,dz)/dz
// for x>0, 1+fmodf(x,dz)/dz otherwise, as 24-bit fixed point in [0, 2^24].
// The remainder is one FMA against a quotient estimate derived from the cell index (exact when
// the estimate is the true truncated quotient; otherwise off by whole cells, which the
// floor below removes), the division by dz a multiplication by RN(1/dz).  Those liberties are
// <= 1 ulp and only touch the density, which is tolerance-checked, never the particle state.
// Particles exactly on a node keep the reference's assignment (x>0: share 0, x<=0: share 1).
__device__ __forceinline__ unsigned left_share_fixed(const Geom& g, float x, int cell) {
    const float q = (float)cell + g.q_shift;
    const float s = __fmul_rn(fmaf(-q, g.dz, x), g.rcp_dz);  // fmod(x,dz)/dz up to whole cells
    const bool pos = x > 0.f;
    float t = pos ? s : -s;
    t = __fsub_rn(t, floorf(t));                             // in [0,1)
    const float share = pos ? t : __fsub_rn(1.0f, t);
    return __float2uint_rn(__fmul_rn(share, (float)kWeightOne));
}

// Same share, computed with one float->int conversion: floor(s*2^24) masked to 24 bits is the
// fractional part for any whole-cell offset of the quotient estimate.  Returns both node
// weights (w_left + w_right == 2^24 exactly).
__device__ __forceinline__ void node_weights_fixed(const Geom& g, float x, int cell, unsigned& w_left,
                                                   unsigned& w_right) {
    const float q = (float)cell + g.q_shift;
    const float s24 = __fmul_rn(fmaf(-q, g.dz, x), g.rcp_dz24);  // 2^24 * fmod(x,dz)/dz up to whole cells
    const bool pos = x > 0.f;
    const unsigned t = (unsigned)__float2int_rd(pos ? s24 : -s24) & (kWeightOne - 1u);  // 2^24 * frac, in [0, 2^24)
    const unsigned c = kWeightOne - t;
    w_left = pos ? t : c;    // x>0: share = frac ; x<=0: share = 1 - frac(-s)  (ref infMagSim_c.c:219-223)
    w_right = pos ? c : t;
}

// ---- context -------------------------------------------------------------------------------
struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace espic

struct espic_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    long long launches = 0;

    // device scratch
    unsigned long long* acc = nullptr;  // fixed-point density accumulators, always left zeroed
    int acc_len = 0;
    int* status = nullptr;              // ESPIC_ST_* bits
    unsigned int* ticket = nullptr;     // last-block-done counters
    espic::Scratch staging;             // removed slot ids, one region per chunk
    espic::Scratch chunk_counts;        // removed count per range
    espic::Scratch range_offsets;       // their exclusive scan
    espic::Scratch dv_table;            // per-cell velocity kick for grids too large for smem
    espic::Scratch solve_ws;            // fp64 rows of the field solve
    espic::Scratch inject_ws;
    espic::Scratch rng_ws;              // MT19937 state + stream buffers
    espic::Scratch sort_ws;             // radix-sort ping-pong buffers and histograms
    espic::Scratch host_stage;          // device staging for the host-pointer entry points
    espic::Scratch filter_ws;           // block centres + moments of the hole-tracking filter
    espic::Scratch debug_ws;            // ESPIC_DEBUG_CTA: per-CTA cycle counts of the last windowed mover launch
    const float* filter_grid = nullptr; // grid whose end points are cached below
    int filter_n = 0;
    float filter_ends[2] = {0.f, 0.f};
    int* dev_int2 = nullptr;            // two device ints used by the host-pointer wrappers
    // the pipelined host-array mover: upload / compute / download streams and per-chunk events
    cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t* pipe_ev = nullptr;
    int pipe_ev_cap = 0;

    // multi-GPU density exchange (see solve.cu)
    struct Exchange {
        int world = 0, rank = 0, n_points = 0;
        size_t slot_floats = 0;         // floats per parity slot (2 * padded n_points)
        void* local = nullptr;          // this rank's buffer
        void* peer[ESPIC_MAX_RANKS] = {nullptr};  // all ranks' buffers as seen from here (peer[rank] == local)
        bool opened = false;
    } xch;

    // mover launch geometry
    int mover_grid = 0, mover_block = 0;
    int mover_ranges = 0;  // ranges of the work split of the last mover launch (see Split)

    // Per-context (hence per-device) launch configuration: the >48 KB shared-memory opt-in of a
    // kernel is a property of the device's copy of the function, so it is set once per context,
    // never once per process (a second context on another device needs its own).
    struct MoveCfg {
        size_t smem = (size_t)-1;  // dynamic shared memory the cached choice was made for
        int per_sm = 0, threads = 0;
    } move_cfg[8];                 // one per k_move instantiation
    bool win_configured = false;
    bool sort_configured = false;
    bool solve_configured = false;
    int solve_max_team = 0;
    size_t hist_smem_set = 0;
    double exchange_timeout_s = 30.;  // espic_exchange_set_timeout

    // espic_pic_run: one step captured as a CUDA graph and replayed (slot 0: without hole tracking,
    // slot 1: with); valid while the step arguments and the context's scratch buffers stay what they were
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        espic_step_args key;
        unsigned long long epoch = 0;
        int launches = 0;              // kernels per replay (for espic_launch_count)
    } step_graph[2];
    cudaStream_t capture_stream = nullptr;
    espic::Scratch run_ctl;                 // control block + injection counts of the stretch being replayed
    int ctl_host_k = -1;                    // host mirror of the control block's step index
    unsigned long long scratch_epoch = 0;   // bumped whenever a scratch buffer is (re)allocated
    int track_cursor_host = -1;             // host mirror of the device cursor (status word 12)

    // per-launch timing of the mover kernel
    int timing_on = 0;
    cudaEvent_t* ev = nullptr;
    int ev_cap = 0, ev_used = 0;

    // MT19937 stream position bookkeeping (host side mirror)
    // Philox key derived from the seed, and the number of sampler launches so far (each launch
    // uses a fresh counter plane, so successive calls continue "the stream")
    unsigned long long rng_key = 0;
    unsigned long long rng_consumed = 0;
    int rng_seeded = 0;
};

namespace espic {

int fail(espic_ctx* ctx, int code, const char* fmt, ...);
int check_cuda(espic_ctx* ctx, cudaError_t e, const char* what);
int ensure(espic_ctx* ctx, Scratch& s, size_t bytes, const char* what);
int ensure_acc(espic_ctx* ctx, int n_points);

// Makes the context's device current for the duration of an entry point and restores the caller's
// afterwards: a context may live on any device of the process, whatever device is current.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const espic_ctx* ctx) {
        if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device)
            switched = cudaSetDevice(ctx->device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define ESPIC_CUDA(ctx, call)                                      \
    do {                                                           \
        int _rc = ::espic::check_cuda((ctx), (call), #call);       \
        if (_rc) return _rc;                                       \
    } while (0)

// Stacks whose newest entries were pushed by direct_push movers (mover_common.cuh) in arrival order: the field
// solve sorts the last *pushed entries of each into ascending slot order and clears the counter.
struct StackFix {
    int* pushed[2];       // device counters: the word behind the last node of each species' accumulator plane
    int* empty_slots[2];
    const int* top[2];
    int largest_allowed_slot[2];
};

// launchers implemented in the other translation units
// Large grids keep the per-cell kick table in global memory (k_build_dv).  A step pushes two species on the same
// grid: the first call can build both tables in one launch (role 1, with the second species' parameters), the
// second call then finds its table ready (role 2).  Each species has its own flag pair behind the tables.
struct DvPlan {
    int role;              // 0: build this call's table only
    const float* phi_b;    // role 1: the second species' potential, charge-to-mass ratio, frame acceleration, dt
    float qm_b, a_b_b, dt_b;
};
int launch_move(espic_ctx* ctx, const float* grid, const float* mask, const float* phi, float dt,
                float qm, float bg, int largest_index, const int* largest_index_dev, float* particles,
                float* density, int* empty_slots, int* top, float a_b, int update_position,
                int periodic, int largest_allowed_slot, int n_points, int storage, int minimal = 0,
                int slot_offset = 0, int finish_density = 1, unsigned long long* acc_fused = nullptr,
                int* pushed = nullptr, int stack_interleave = 0, const DvPlan* dv_plan = nullptr);
int launch_tridiagonal(espic_ctx* ctx, double* a, double* b, double* c, double* d, double* x, int n);
int move_window_cells(espic_ctx* ctx, int n_points);  // 0: whole-grid tables fit in shared memory
int launch_selftest_quotient(espic_ctx* ctx, const float* grid, int n_points, unsigned long long* counts_out);
int launch_sort(espic_ctx* ctx, const float* grid, int n_points, float* particles, int storage, int* tags,
                int* empty_slots, int n_stack, int* top, int* largest, int largest_index, int periodic,
                int key_shift, float lookahead);
int launch_shift_velocities(espic_ctx* ctx, float* particles, int storage, int largest_index, double v_b);
int launch_poisson(espic_ctx* ctx, const float* grid, const float* mask, const float* charge,
                   float debye, float* potential, float obj_pot, float transparency, int boltzmann,
                   int periodic, int n_points);
int launch_field_solve(espic_ctx* ctx, const float* grid, const float* mask, float* ion, float* ele,
                       float* charge_out, int periodic_particles, int fix_i, int fix_e, float debye,
                       float* potential, float obj_pot, float transparency, int boltzmann, int periodic_potential,
                       int n_points, unsigned long long* acc_i = nullptr, unsigned long long* acc_e = nullptr,
                       float bg = 1.f, const StackFix* fix = nullptr);
int exchange_create(espic_ctx* ctx, int n_points, int world, int rank, void* handle_out);
int exchange_open(espic_ctx* ctx, const void* handles);
int exchange_publish(espic_ctx* ctx, int parity, int step, unsigned long long* acc_i = nullptr,
                     unsigned long long* acc_e = nullptr, float bg = 1.f);
int launch_field_solve_exchange(espic_ctx* ctx, int parity, int step, const float* grid, const float* mask,
                                float* ion, float* ele, float* charge_out, int periodic_particles, int fix_i,
                                int fix_e, float debye, float* potential, float obj_pot, float transparency,
                                int boltzmann, int periodic_potential, int n_points, const StackFix* fix = nullptr);
int launch_prepare_charge(espic_ctx* ctx, float* ion, float* ele, float* charge, int n, int periodic,
                          int fix_i, int fix_e);
// Device control block of a replayed step graph (espic_pic_run, see InjectArgs in inject.cu): the injection kernels
// read the step index, the counts and the sampler's call counter from it; n_inject is then an upper bound.
constexpr int kCtlMaxSteps = 4096;   // steps of one espic_pic_run call whose injection counts fit the control block
struct InjectCtl {
    int* ptr;
    int species, stride, advance;
};
int launch_inject(espic_ctx* ctx, int n_inject, const float* grid, int n_points, float dt, float bg,
                  const float* vel, const float* uni, float a_b, int k, int* hist, float* particles,
                  int storage, int* empty_slots, int* top, int* largest, float* density, bool classified = false,
                  unsigned long long* acc_fused = nullptr, int by_side = 0, const InjectCtl* ctl = nullptr);
struct InjectPairSpecies {
    int n_inject;
    float v_th, v_d;
    int* hist;
    float* particles;
    int storage;
    int* empty_slots;
    int* top;
    int* largest;
    float* density;
    unsigned long long* acc_fused;
};
int launch_inject_pair(espic_ctx* ctx, const InjectPairSpecies sp[2], const float* grid, int n_points, float dt, float bg,
                       float a_b, int k, int by_side, int* ctl, int ctl_stride);
int rng_sample_for_injection(espic_ctx* ctx, int n_inject, float v_th, float v_d, float* vel, float* uni,
                             const float* grid, int n_points, float dt, float bg, float a_b,
                             const InjectCtl* ctl = nullptr);
int launch_convert_density(espic_ctx* ctx, unsigned long long* acc, float* density, int n_points, float bg);
int launch_convert_density2(espic_ctx* ctx, unsigned long long* acc0, float* den0, unsigned long long* acc1, float* den1,
                            int n_points, float bg);
int rng_seed(espic_ctx* ctx, unsigned long seed);
int rng_draw_velocities(espic_ctx* ctx, int n, float v_th, float v_d, float* v_out);
int rng_uniforms(espic_ctx* ctx, int n, float* out);
int launch_field_filter(espic_ctx* ctx, int n_points, const float* grid, const float* data, float L,
                        int n_fine, const float* fine, float* res);
int launch_hole_position(espic_ctx* ctx, int n_points, const float* grid, const float* potential,
                         double dz, float L, int n_fine, const float* fine, float* scratch_res,
                         float* pos_out, int* cursor = nullptr);
int launch_histogram(espic_ctx* ctx, const float* xs, const float* ys, int n_data, const int* tags,
                     int tag_value, float x_min, float step_x, float y_min, float step_y,
                     int nx, int ny, int* hist);

}  // namespace espic
