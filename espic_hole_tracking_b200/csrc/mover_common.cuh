// Device code shared by the mover kernels (mover.cu: whole grid in shared memory or in global
// memory; mover_win.cu: a window of a large grid in shared memory): the slot update in its generic
// and fast forms, the fixed-point deposit, the ordered staging of removed slots and the static
// work split.
#pragma once

#include "espic_internal.cuh"

// development switches for A/B builds (scripts/build_variants.sh); the defaults are the shipped code
#ifndef ESPIC_VAR_CARRY
#define ESPIC_VAR_CARRY 0     // 1: one running maximum over the eight returned words; 0: one compare per word
                              // (measured on B200, 1e8 particles: 0.2937 vs 0.2920 ms per launch -- no gain in k_move)
#endif
#ifndef ESPIC_DEBUG_CYCLES
#define ESPIC_DEBUG_CYCLES 0  // 1: the windowed mover records clock cycles per segment and per CTA (espic_debug_cta_cycles)
#endif
#ifndef ESPIC_VAR_TMA
#define ESPIC_VAR_TMA 1       // 1: units reach the staging ring as 1 KB bulk copies (TMA) issued by one lane and completed
#endif                        //    on an mbarrier; 0: four 16-byte cp.async per lane (0.2868 -> 0.2797 ms per 1e8 particles)
#ifndef ESPIC_TMA_STAGES
#define ESPIC_TMA_STAGES 2
#endif
#ifndef ESPIC_VAR_STAGE_WALLS
#define ESPIC_VAR_STAGE_WALLS 0   // 1: the wall-bounded whole-grid kernel uses the staging ring too
#endif
#ifndef ESPIC_VAR_UNPACK
#define ESPIC_VAR_UNPACK 0    // 1: the second quad is read from the ring after the first has been pushed (no spill in
                              // the loop, but the exposed LDS latency costs more: 0.3058 vs 0.2920 ms per launch)
#endif
#if ESPIC_VAR_TMA && ESPIC_VAR_UNPACK
#error "ESPIC_VAR_UNPACK belongs to the cp.async ring (ESPIC_VAR_TMA=0)"
#endif
#ifndef ESPIC_VAR_EPILOGUE
#define ESPIC_VAR_EPILOGUE 0  // 1: last-CTA removal epilogue compiled into the round-robin periodic kernel (the extra
                              // code at the end of the kernel perturbs the main loop: 0.3038 vs 0.2920 ms per launch)
#endif

namespace espic {

#ifndef ESPIC_MOVE_BOUNDS
#define ESPIC_MOVE_BOUNDS 896   // CTA width of k_move: 28 warps per SM at 72 registers per thread.  Measured (B200, 1e8 particles,
                                // 4096 cells, ms per launch, periodic / walls): 512 -> 0.288 / 0.320, 640 -> 0.294 / 0.294,
                                // 768 -> 0.295 / 0.285, 896 -> 0.287 / 0.279, 1024 (64 registers) -> 0.294 / 0.291
#endif
constexpr int kTmaStages = ESPIC_VAR_TMA ? ESPIC_TMA_STAGES : 2;
constexpr int kRingBytesPerWarp = 2048 * kTmaStages;   // one 2 KB unit (x, v of 256 slots) per stage
#ifndef ESPIC_MOVE_BOUNDS_STAGED
#define ESPIC_MOVE_BOUNDS_STAGED 512   // CTA width of the periodic whole-grid kernels, whose units arrive through the bulk-copy ring:
                                       // 16 warps at 108 registers, nothing spilled.  Measured (B200, 1e8 particles, 4096 cells, ms per
                                       // launch): 256 x2 -> 0.282, 384 -> 0.294, 512 -> 0.261, 640 -> 0.268, 768 -> 0.266, 896 -> 0.272
#endif
constexpr int kMoveThreads = ESPIC_MOVE_BOUNDS;  // one CTA per SM: one tile to zero and flush per SM
constexpr int kMoveThreadsStaged = ESPIC_MOVE_BOUNDS_STAGED;
template <bool PERIODIC, bool SMEM>
constexpr int move_threads() { return (PERIODIC && SMEM) ? kMoveThreadsStaged : kMoveThreads; }
constexpr int kUnitQuads = 64;                // one unit = two warp instructions of 32 quads
constexpr int kUnitSlots = 4 * kUnitQuads;    // 256 slots
constexpr int kSmemBytesPerPoint = 12;

struct MoveArgs {
    const float* grid;
    const float* mask;
    const float* phi;
    float* px;
    float* pv;
    int* staging;
    int* counts;
    int* total_removed;  // sum of counts[] (lets k_finish skip the scan when nothing left)
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
    int prefetch_units;  // L2 prefetch distance along a warp's stream, in 256-slot units
    int staged;          // 1: units are staged through shared memory with cp.async, one unit ahead
    int tile_bytes;      // size of the grid tables in shared memory (the staging ring follows them)
    int strided;         // experiment: units dealt round-robin to the warps instead of in contiguous ranges
    int* seg_counter;    // windowed kernel: next segment to hand out (reset by k_finish)
    int n_segments;      // windowed kernel: segments of 32 ranges each
    // fused step path, periodic variant with per-unit ranges: a removed slot -- in a periodic box that needs a
    // particle flung past 0.99*2*z_max in one step -- is pushed on the stack right away by the (out-of-line,
    // practically never executed) generic slot update, so a step needs no k_finish / k_scatter launch.  Slots
    // removed by one launch reach the stack in arrival order; the next field solve puts that stretch of the
    // stack into the ascending order the reference's sequential loop produces (solve.cu, StackFix).
    int direct_push;
    int* pushed;      // slots pushed this way since the last fix-up
    int epilogue;
    unsigned* ticket;
    int* top;
    int* empty_slots;
    int largest_allowed_slot;
    int slot_offset;
    long long* cta_cycles;  // development aid (ESPIC_DEBUG_CTA=1): per CTA, clock cycles from entry to exit
    int minimal;  // move_particles_c_minimal (ref infMagSim_c.c:231-341): no mask, and a slot that is out of
                  // bounds on entry is left alone even when it does not carry the parked marker (ref :333)
};

// The velocity kick of a cell, read straight from a table indexed by cell.
struct KickDirect {
    const float* dv;
    __device__ __forceinline__ float operator()(int c) const { return dv[c]; }
};

enum { ACT_NONE = 0, ACT_DEPOSIT = 1, ACT_REMOVE = 2 };

// ---------------------------------------------------------------------------------------------
// Generic slot update: the reference's loop body (ref infMagSim_c.c:143-217) in full generality
// (update_position on/off, absorbing object, any geometry, any position), everything except
// the deposit.  Bit-exact in x and v: each rounded operation of the reference is one rounded
// intrinsic here.  Used out of line for the tail of the store, for start-up calls
// (accumulate_density / initialize_mover) and whenever the fast path below meets a case it
// does not handle.
// ---------------------------------------------------------------------------------------------
template <bool PERIODIC>
__device__ __forceinline__ int push_generic(const Geom& g, float& x_io, float& v_io, const float* dv, float dt,
                                            bool update_position, bool has_obj, const float* __restrict__ grid,
                                            const float* __restrict__ mask, int& cell_out, bool& bad_any,
                                            bool minimal = false) {
    float x = x_io, v = v_io;
    bool alive;
    if (PERIODIC) {
        alive = x < g.live_thr;                       // ref :146
        if (alive) x = fold_periodic(g, x);           // ref :148-153
    } else {
        alive = (x > g.lo_edge) && (x < g.hi_edge);   // ref :156
    }
    const bool alive_at_entry = alive;
    int cell = 0;
    if (alive_at_entry) {
        bool bad = false;
        cell = cell_of<PERIODIC, false>(g, x, bad);   // ref :161-165
        v = __fadd_rn(v, dv[cell]);                   // ref :166-168 via the per-cell kick table
        if (update_position) {
            x = __fadd_rn(x, __fmul_rn(v, dt));       // ref :170
            if (PERIODIC) {
                alive = x < g.live_thr;               // ref :172
                if (alive) x = fold_periodic(g, x);   // ref :174-179
            } else {
                alive = (x > g.lo_edge) && (x < g.hi_edge);  // ref :182
            }
            if (alive) cell = cell_of<PERIODIC, false>(g, x, bad);  // ref :185-189
        }
        bad_any |= bad;
    }
    bool swallowed = false;
    if (has_obj && alive) {  // ref :200-210, evaluated in double there
        const float m0 = __ldg(mask + cell);
        if (m0 > 0.f) {
            const float m1 = __ldg(mask + cell + 1);
            const double covered = (m1 > 0.f) ? 1. - (double)m0 : (double)m0;
            swallowed = (double)x > (double)__ldg(grid + cell) + covered * (double)g.dz;
        }
    }
    cell_out = cell;
    int act = ACT_NONE;
    // ref :211-212; the minimal variant tests within_bounds_before_move alone (ref :333)
    if (alive_at_entry || (!minimal && x < g.live_thr)) act = (!alive || swallowed) ? ACT_REMOVE : ACT_DEPOSIT;
    x_io = (act == ACT_REMOVE) ? g.parked : x;
    v_io = v;
    return act;
}

// lo += w with the carry into hi taken from the value the atomic returns; the carry update is
// a predicated shared-memory reduction.
__device__ __forceinline__ void smem_add_carry(unsigned* lo, unsigned* hi, unsigned w) {
    const unsigned old = atomicAdd(lo, w);
    if (old > ~w) atomicAdd(hi, 1u);
}

template <bool SMEM, bool RESEED = false>
__device__ __forceinline__ void deposit_one(const Geom& g, float x, int cell, unsigned* s_lo,
                                            unsigned* s_hi, unsigned long long* acc) {
    // The weights depend on x alone (fmodf(x,dz)/dz, ref :219-223); the cell only seeds the quotient
    // estimate, which must be within 127 cells of x/dz.  The generic path (RESEED) also sees slots
    // whose deposit cell is not their position's cell -- the periodic overflow to cell 0 just below
    // z_max (ref :186-187) -- so it seeds from the position itself; same bits whenever the two agree.
    int seed = cell;
    if (RESEED) seed = min(max(__float2int_rz(quotient_dz<false>(g, __fsub_rn(x, g.z_lo))), 0), g.last_cell + 1);
    unsigned w0, w1;
    node_weights_fixed(g, x, seed, w0, w1);
    if (SMEM) {
        smem_add_carry(s_lo + cell, s_hi + cell, w0);
        smem_add_carry(s_lo + cell + 1, s_hi + cell + 1, w1);
    } else {
        atomicAdd(acc + cell, (unsigned long long)w0);
        atomicAdd(acc + cell + 1, (unsigned long long)w1);
    }
}

// Deposit of a whole quad without per-slot branches: the eight atomics are issued unconditionally.
// A word can only wrap when its old value lies within 2^24 of 2^32 (a weight is at most 2^24), so one
// running maximum over the eight returned values tells whether ANY of them might have carried; the exact
// per-word test and the carry into the high plane run behind one warp vote, practically never.
template <bool SMEM>
__device__ __forceinline__ void deposit_quad_all(const Geom& g, const float* x, const int* cell, unsigned* s_lo,
                                                 unsigned* s_hi, unsigned long long* acc) {
    if (SMEM) {
#if ESPIC_VAR_CARRY
        unsigned w[8], old[8], top = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            node_weights_fixed(g, x[c], cell[c], w[2 * c], w[2 * c + 1]);
            old[2 * c] = atomicAdd(s_lo + cell[c], w[2 * c]);
            old[2 * c + 1] = atomicAdd(s_lo + cell[c] + 1, w[2 * c + 1]);
            top = max(top, max(old[2 * c], old[2 * c + 1]));
        }
        if (__any_sync(0xffffffffu, top >= 0xFF000000u)) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (old[2 * c] > ~w[2 * c]) atomicAdd(s_hi + cell[c], 1u);
                if (old[2 * c + 1] > ~w[2 * c + 1]) atomicAdd(s_hi + cell[c] + 1, 1u);
            }
        }
#else
        unsigned carries = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned w0, w1;
            node_weights_fixed(g, x[c], cell[c], w0, w1);
            const unsigned old0 = atomicAdd(s_lo + cell[c], w0);
            const unsigned old1 = atomicAdd(s_lo + cell[c] + 1, w1);
            if (old0 > ~w0) carries |= 1u << (2 * c);
            if (old1 > ~w1) carries |= 2u << (2 * c);
        }
        if (__any_sync(0xffffffffu, carries != 0u)) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (carries & (1u << (2 * c))) atomicAdd(s_hi + cell[c], 1u);
                if (carries & (2u << (2 * c))) atomicAdd(s_hi + cell[c] + 1, 1u);
            }
        }
#endif
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned w0, w1;
            node_weights_fixed(g, x[c], cell[c], w0, w1);
            atomicAdd(acc + cell[c], (unsigned long long)w0);
            atomicAdd(acc + cell[c] + 1, (unsigned long long)w1);
        }
    }
}

// The same for a quad of which only the slots named by `dep` deposit: the others add zero weights to
// their (valid, clamped) cell, which changes nothing and needs no branch.
__device__ __forceinline__ void deposit_quad_masked(const Geom& g, const float* x, const int* cell, unsigned dep,
                                                    unsigned* s_lo, unsigned* s_hi) {
    unsigned carries = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned w0, w1;
        node_weights_fixed(g, x[c], cell[c], w0, w1);
        const bool on = (dep >> c) & 1u;
        w0 = on ? w0 : 0u;
        w1 = on ? w1 : 0u;
        const unsigned old0 = atomicAdd(s_lo + cell[c], w0);
        const unsigned old1 = atomicAdd(s_lo + cell[c] + 1, w1);
        if (old0 > ~w0) carries |= 1u << (2 * c);
        if (old1 > ~w1) carries |= 2u << (2 * c);
    }
    if (__any_sync(0xffffffffu, carries != 0u)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (carries & (1u << (2 * c))) atomicAdd(s_hi + cell[c], 1u);
            if (carries & (2u << (2 * c))) atomicAdd(s_hi + cell[c] + 1, 1u);
        }
    }
}

// One quad (4 consecutive slots, possibly extending past n_slots) through the generic update:
// scalar loads, push, deposit, scalar stores.  Returns the 4-bit mask of removed components.
template <bool PERIODIC, bool SMEM>
__device__ __noinline__ unsigned quad_generic(const MoveArgs& a, int q, int n_slots, const float* dv,
                                              bool has_obj, unsigned* s_lo, unsigned* s_hi) {
    const Geom g = make_geom(a.grid, a.n_points);
    const bool upd = a.update_position != 0;
    unsigned rem = 0;
    bool bad_any = false;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        const int s = 4 * q + c;
        if (s >= n_slots) break;
        float x = a.px[s], v = a.pv[s];
        int cell;
        const int act = push_generic<PERIODIC>(g, x, v, dv, a.dt, upd, has_obj, a.grid, a.mask, cell, bad_any,
                                               a.minimal != 0);
        a.px[s] = x;
        a.pv[s] = v;
        if (act == ACT_DEPOSIT) deposit_one<SMEM, true>(g, x, cell, s_lo, s_hi, a.acc);
        if (act == ACT_REMOVE && a.direct_push) {
            const int pos = atomicAdd(a.top, 1) + 1;
            if (pos > a.largest_allowed_slot) atomicOr(a.status, ESPIC_ST_STACK_OVERFLOW);
            else if (pos >= 0) a.empty_slots[pos] = s + a.slot_offset;
            atomicAdd(a.pushed, 1);
        } else {
            rem |= (act == ACT_REMOVE) ? (1u << c) : 0u;
        }
    }
    if (bad_any) atomicOr(a.status, ESPIC_ST_BAD_LEFT_INDEX);
    return rem;
}

// ---------------------------------------------------------------------------------------------
// Fast slot update: the production case (position update on, no absorbing object, dz in the
// normal exponent range), straight-line code.  Everything it does not cover raises `rare`
// instead of being handled: the caller then discards the quad's results and sends the quad
// through quad_generic.  Rare cases: a position outside one period of the domain (periodic),
// a folded position landing exactly on z_max, a cell index outside the grid, a NaN, and -- in
// periodic mode only, where it does not occur in practice -- a removal.
// On return x,v hold the updated slot (unchanged for dead slots), `cell` the deposit cell,
// and the result says whether the slot deposits; `removed` (non-periodic) whether it was parked.
// ---------------------------------------------------------------------------------------------
// Periodic flavour.  Covers a LIVE slot whose position is inside the domain on entry and moves by
// less than a period; everything else (dead slots included) raises `rare`.  Returns nothing:
// in the fast path every slot of the quad deposits.
template <class Kick>
__device__ __forceinline__ void push_fast_periodic(const Geom& g, float& x_io, float& v_io, const Kick& kick,
                                                   float dt, int& cell_out, bool& rare) {
    const unsigned last = (unsigned)g.last_cell;
    // ref :146-153 for a position already inside [z_lo, z_hi): fmodf is the identity
    const float a0 = __fsub_rn(x_io, g.z_lo);
    const float x0 = __fadd_rn(a0, g.z_lo);
    const float b0 = __fsub_rn(x0, g.z_lo);                                 // ref :161 (x - z_min)
    rare |= !(a0 >= 0.f);                 // would wrap from below (or NaN)
    rare |= !(b0 < g.cell_span);          // would wrap from above, folded onto z_max (or so close that the cell
                                          // index overflows: the reference takes cell 0 then), or a parked slot
    const int c0 = (int)min((unsigned)__float2int_rz(quotient_dz<true>(g, b0)), last);
    const float v1 = __fadd_rn(v_io, kick(c0));                              // ref :166-168
    const float x1 = __fadd_rn(x0, __fmul_rn(v1, dt));                      // ref :170
    // ref :172-179 within one period either side (and below the parked threshold, ref :172)
    const float a1 = __fsub_rn(x1, g.z_lo);
    rare |= !(a1 > g.fold_lo);
    rare |= !(a1 < g.fold_hi);
    const float off = __fsub_rn(a1, (a1 >= g.span) ? g.span : 0.f);
    const float x2 = __fadd_rn(off, (off >= 0.f) ? g.z_lo : g.z_hi);
    const float b2 = __fsub_rn(x2, g.z_lo);                                 // ref :185
    rare |= !(b2 < g.cell_span);          // folded onto z_max or within rounding of it: the reference takes cell 0
    cell_out = (int)min((unsigned)__float2int_rz(quotient_dz<true>(g, b2)), last);
    x_io = x2;
    v_io = v1;
}

// Wall flavour: dead slots are common (removed particles wait for re-injection), so liveness
// is handled with selects.  Returns whether the slot deposits; `removed` whether it was parked.
template <class Kick>
__device__ __forceinline__ bool push_fast_walls(const Geom& g, float& x_io, float& v_io, const Kick& kick, float dt,
                                                int& cell_out, bool& removed, bool& rare) {
    const float x_in = x_io, v_in = v_io;
    const unsigned last = (unsigned)g.last_cell;
    const bool alive0 = (x_in > g.lo_edge) && (x_in < g.hi_edge);        // ref :156
    const int c0r = __float2int_rz(quotient_dz<true>(g, __fsub_rn(x_in, g.z_lo)));  // ref :161
    const int c0 = (int)min((unsigned)c0r, last);
    const float v1 = __fadd_rn(v_in, kick(c0));                          // ref :166-168
    const float x1 = __fadd_rn(x_in, __fmul_rn(v1, dt));                // ref :170
    const bool alive1 = (x1 > g.lo_edge) && (x1 < g.hi_edge);           // ref :182
    const int c1r = __float2int_rz(quotient_dz<true>(g, __fsub_rn(x1, g.z_lo)));    // ref :185
    const int c1 = (int)min((unsigned)c1r, last);
    rare |= alive0 && ((c0 != c0r) || (alive1 && (c1 != c1r)) || !(x1 == x1));
    cell_out = c1;
    // ref :211-217: alive slots that left the domain, and out-of-bounds slots that do not
    // carry the parked marker, are parked and pushed on the stack
    removed = alive0 ? !alive1 : (x_in < g.live_thr);
    x_io = removed ? g.parked : (alive0 ? x1 : x_in);
    v_io = alive0 ? v1 : v_in;
    return alive0 && alive1;
}

struct Quad {
    float x[4], v[4];
};

__device__ __forceinline__ void load_quad(const MoveArgs& a, int q, Quad& p) {
    const float4 xx = __ldcs(reinterpret_cast<const float4*>(a.px) + q);
    const float4 vv = __ldcs(reinterpret_cast<const float4*>(a.pv) + q);
    p.x[0] = xx.x; p.x[1] = xx.y; p.x[2] = xx.z; p.x[3] = xx.w;
    p.v[0] = vv.x; p.v[1] = vv.y; p.v[2] = vv.z; p.v[3] = vv.w;
}

__device__ __forceinline__ void unpack_quad(const float4 xx, const float4 vv, Quad& p) {
    p.x[0] = xx.x; p.x[1] = xx.y; p.x[2] = xx.z; p.x[3] = xx.w;
    p.v[0] = vv.x; p.v[1] = vv.y; p.v[2] = vv.z; p.v[3] = vv.w;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 1-D bulk copies (TMA) completed on an mbarrier: one lane moves a whole warp-unit.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void store_quad(const MoveArgs& a, int q, const Quad& p) {
    __stcs(reinterpret_cast<float4*>(a.px) + q, make_float4(p.x[0], p.x[1], p.x[2], p.x[3]));
    __stcs(reinterpret_cast<float4*>(a.pv) + q, make_float4(p.v[0], p.v[1], p.v[2], p.v[3]));
}

// Fast path for one complete quad held in registers: push the four slots; if the warp met no
// rare case, deposit and store; otherwise redo this quad generically from memory.
template <bool PERIODIC, bool SMEM>
__device__ __forceinline__ unsigned quad_fast(const MoveArgs& a, const Geom& g, Quad& p, int q, int n_slots,
                                              const float* dv, unsigned* s_lo, unsigned* s_hi) {
    int cell[4];
    bool rare = false;
    if (PERIODIC) {
#pragma unroll
        for (int c = 0; c < 4; ++c) push_fast_periodic(g, p.x[c], p.v[c], KickDirect{dv}, a.dt, cell[c], rare);
        if (__any_sync(0xffffffffu, rare))  // uniform, practically never taken once the run is under way
            return quad_generic<PERIODIC, SMEM>(a, q, n_slots, dv, false, s_lo, s_hi);
        store_quad(a, q, p);
        deposit_quad_all<SMEM>(g, p.x, cell, s_lo, s_hi, a.acc);
        return 0u;
    } else {
        unsigned dep = 0, rem = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            bool removed;
            const bool d = push_fast_walls(g, p.x[c], p.v[c], KickDirect{dv}, a.dt, cell[c], removed, rare);
            dep |= d ? (1u << c) : 0u;
            rem |= removed ? (1u << c) : 0u;
        }
        if (__any_sync(0xffffffffu, rare)) return quad_generic<PERIODIC, SMEM>(a, q, n_slots, dv, false, s_lo, s_hi);
        store_quad(a, q, p);
        if (__all_sync(0xffffffffu, dep == 0xFu)) {  // uniform: the whole warp deposits
            deposit_quad_all<SMEM>(g, p.x, cell, s_lo, s_hi, a.acc);
        } else if (SMEM) {
            // some slot in the warp is dead or has just left (half of all warp iterations at 0.5 %
            // dead slots): same branch-free deposit, the slots that do not deposit add zero weights
            deposit_quad_masked(g, p.x, cell, dep, s_lo, s_hi);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (dep & (1u << c)) deposit_one<SMEM>(g, p.x[c], cell[c], s_lo, s_hi, a.acc);
        }
        return rem;
    }
}

// Ordered compaction of the removed slot ids of one warp iteration (ascending slot order):
// lanes hold consecutive quads, so a lane's ids follow all ids of lower lanes.
__device__ __forceinline__ void stage_removed(unsigned rem, int q, int lane, int* stage, int& removed) {
    if (__ballot_sync(0xffffffffu, rem != 0u)) {  // uniform
        const unsigned b0 = __ballot_sync(0xffffffffu, rem & 1u);
        const unsigned b1 = __ballot_sync(0xffffffffu, rem & 2u);
        const unsigned b2 = __ballot_sync(0xffffffffu, rem & 4u);
        const unsigned b3 = __ballot_sync(0xffffffffu, rem & 8u);
        const unsigned lt = (1u << lane) - 1u;
        int pos = removed + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (rem & (1u << c)) stage[pos++] = 4 * q + c;
        removed += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
    }
}

// Work split: the complete units [0, n_units) are dealt to the W warps of the grid as contiguous,
// ascending ranges whose lengths differ by at most one unit (static and balanced to 256
// slots); the leftover slots behind the last complete unit form one more range owned by the
// last warp.  Range c stages its removed slot ids at staging + first_slot(c).
struct Split {
    int n_units, base, extra;
    __host__ __device__ Split(int n_slots, int n_warps) {
        n_units = (n_slots >> 2) / kUnitQuads;
        base = n_units / n_warps;
        extra = n_units % n_warps;
    }
    __host__ __device__ int first_unit(int w) const { return w * base + (w < extra ? w : extra); }
    __host__ __device__ int unit_count(int w) const { return base + (w < extra ? 1 : 0); }
};

// mover_win.cu
int move_window_size(int n_points);  // window cells, 0 when the windowed kernel does not apply
int launch_move_window(espic_ctx* ctx, MoveArgs& a, int periodic);

}  // namespace espic
