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
    float inv_dz;    // only used for deposit weights (tolerance-checked, not bit-checked)
    int last_cell;   // n_points - 2
};

__device__ __forceinline__ Geom make_geom(const float* __restrict__ grid, int n_points) {
    Geom g;
    g.z_lo = __ldg(grid);
    g.z_hi = __ldg(grid + n_points - 1);
    g.dz = (g.z_hi - g.z_lo) / (float)(n_points - 1);
    g.span = g.z_hi - g.z_lo;
    g.parked = (float)(2. * (double)g.z_hi);
    g.live_thr = __double2float_ru(0.99 * (double)g.parked);
    g.lo_edge = g.z_lo + 1.e-6f;
    g.hi_edge = g.z_hi - 1.e-6f;
    g.inv_dz = 1.0f / g.dz;
    g.last_cell = n_points - 2;
    return g;
}

// fmodf(x - z_lo, span) followed by the reference's re-basing (ref infMagSim_c.c:148-153).
// The three branches in front are exact special cases of fmodf (identity inside one period,
// Sterbenz-exact subtraction one period above); everything else takes the library routine,
// which is exact as well.
__device__ __forceinline__ float fold_periodic(const Geom& g, float x) {
    float a = x - g.z_lo;
    float off;
    if (a >= 0.f && a < g.span) {
        off = a;
    } else if (a >= g.span && a < 2.f * g.span) {
        off = a - g.span;
    } else if (a < 0.f && a > -g.span) {
        off = a;
    } else {
        off = fmodf(a, g.span);
    }
    return (off >= 0.f) ? off + g.z_lo : off + g.z_hi;
}

// (int)((x - z_lo)/dz) with IEEE division and truncation (ref infMagSim_c.c:161-163).
// Returns the reference's index; `bad` is raised when it lies outside [0, n_points-2]
// (the reference prints and then indexes out of bounds; callers clamp).
template <bool PERIODIC>
__device__ __forceinline__ int cell_of(const Geom& g, float x, bool& bad) {
    int c = (int)((x - g.z_lo) / g.dz);
    if (PERIODIC && c > g.last_cell) c = 0;
    bad = (c < 0) || (c > g.last_cell);
    return c;
}

// Share of the particle given to its LEFT node (ref infMagSim_c.c:219-223): fmodf(x,dz)/dz
// for x>0, 1+fmodf(x,dz)/dz otherwise.  The remainder is formed exactly (one fused
// multiply-add against the truncated quotient, corrected by +-1 if the quotient estimate was
// off); the division by dz is a multiplication by 1/dz, a 1-ulp liberty that only touches
// the tolerance-checked density.  Returned as 24-bit fixed point in [0, 2^24].
__device__ __forceinline__ unsigned left_share_fixed(const Geom& g, float x) {
    float q = truncf(x * g.inv_dz);
    float r = fmaf(-q, g.dz, x);
    if (x > 0.f) {
        if (r < 0.f) r = fmaf(-(q - 1.f), g.dz, x);
        else if (r >= g.dz) r = fmaf(-(q + 1.f), g.dz, x);
    } else {
        if (r > 0.f) r = fmaf(-(q + 1.f), g.dz, x);
        else if (r <= -g.dz) r = fmaf(-(q - 1.f), g.dz, x);
    }
    float share = r * g.inv_dz;
    if (!(x > 0.f)) share = 1.0f + share;
    share = fminf(fmaxf(share, 0.f), 1.f);
    return __float2uint_rn(share * (float)kWeightOne);
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
    espic::Scratch chunk_counts;        // removed count per chunk (+ scanned offsets)
    espic::Scratch dv_table;            // per-cell velocity kick for grids too large for smem
    espic::Scratch solve_ws;            // fp64 rows of the field solve
    espic::Scratch inject_ws;
    espic::Scratch rng_ws;              // MT19937 state + stream buffers
    espic::Scratch host_stage;          // device staging for the host-pointer entry points
    int* dev_int2 = nullptr;            // two device ints used by the host-pointer wrappers

    // mover launch geometry
    int mover_grid = 0, mover_block = 0;

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

#define ESPIC_CUDA(ctx, call)                                      \
    do {                                                           \
        int _rc = ::espic::check_cuda((ctx), (call), #call);       \
        if (_rc) return _rc;                                       \
    } while (0)

// launchers implemented in the other translation units
int launch_move(espic_ctx* ctx, const float* grid, const float* mask, const float* phi, float dt,
                float qm, float bg, int largest_index, const int* largest_index_dev, float* particles,
                float* density, int* empty_slots, int* top, float a_b, int update_position,
                int periodic, int largest_allowed_slot, int n_points, int storage);
int launch_shift_velocities(espic_ctx* ctx, float* particles, int storage, int largest_index, double v_b);
int launch_poisson(espic_ctx* ctx, const float* grid, const float* mask, const float* charge,
                   float debye, float* potential, float obj_pot, float transparency, int boltzmann,
                   int periodic, int n_points);
int launch_prepare_charge(espic_ctx* ctx, float* ion, float* ele, float* charge, int n, int periodic,
                          int fix_i, int fix_e);
int launch_inject(espic_ctx* ctx, int n_inject, const float* grid, int n_points, float dt, float bg,
                  const float* vel, const float* uni, float a_b, int k, int* hist, float* particles,
                  int storage, int* empty_slots, int* top, int* largest, float* density);
int rng_seed(espic_ctx* ctx, unsigned long seed);
int rng_draw_velocities(espic_ctx* ctx, int n, float v_th, float v_d, float* v_out);
int rng_uniforms(espic_ctx* ctx, int n, float* out);
int launch_field_filter(espic_ctx* ctx, int n_points, const float* grid, const float* data, float L,
                        int n_fine, const float* fine, float* res);
int launch_hole_position(espic_ctx* ctx, int n_points, const float* grid, const float* potential,
                         double dz, float L, int n_fine, const float* fine, float* scratch_res,
                         float* pos_out);
int launch_histogram(espic_ctx* ctx, const float* xs, const float* ys, int n_data, const int* tags,
                     int tag_value, float x_min, float step_x, float y_min, float step_y,
                     int nx, int ny, int* hist);

}  // namespace espic
