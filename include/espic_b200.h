/*
 * espic_b200.h -- C ABI of libespic_b200.so, the sm_100a implementation of the per-timestep
 * hot path of ctzhou/ESPIC_hole_tracking's 1-D infinite-B electrostatic PIC.
 *
 * Two families of entry points:
 *
 *  (A) espic_*  -- DEVICE-pointer operators on an explicit context (stream + scratch).
 *      Particle stores, grids and the free-slot stack stay resident in HBM between calls;
 *      the two in/out scalars of the reference (stack top, largest used index) are DEVICE
 *      ints so a timestep needs no host synchronisation.  This is what the Python operator
 *      layer (espic_hole_tracking_b200/operators.py) and the driver mirror call.
 *
 *  (B) the reference's own symbol names with HOST pointers (move_particles_c,
 *      poisson_solve_c, ...), same parameter lists as infMagSim_c.c, for relinking the
 *      reference's Cython extension against this library without touching the driver.
 *      Each call stages its arrays through the device (H2D, kernels, D2H).
 *
 * "ref:" citations are file:line in the reference repository.
 * All functions returning int return 0 on success or a negative ESPIC_E_* code; the message
 * is available from espic_last_error().  No entry point falls back to CPU arithmetic.
 */
#ifndef ESPIC_B200_H
#define ESPIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESPIC_ABI_VERSION 1

enum {
    ESPIC_OK = 0,
    ESPIC_E_CUDA = -1,     /* a CUDA runtime call failed */
    ESPIC_E_ARG = -2,      /* invalid argument */
    ESPIC_E_NOMEM = -3,    /* scratch allocation failed */
    ESPIC_E_NODEVICE = -4  /* no CUDA device / wrong architecture */
};

/* Bits of the device status word (espic_status): the reference only printf()s in these
 * situations and then indexes out of bounds; we record the event and stay in bounds. */
enum {
    ESPIC_ST_BAD_LEFT_INDEX = 1,   /* ref infMagSim_c.c:164-165,188-189 "bad left_index" */
    ESPIC_ST_STACK_OVERFLOW = 2,   /* ref infMagSim_c.c:215-216 "bad current_empty_slot" */
    ESPIC_ST_STACK_EMPTY = 4,      /* ref infMagSim_cython.py:255-256 "no empty slots" */
    ESPIC_ST_BAD_INJECT_INDEX = 8, /* ref infMagSim_cython.py:272-273 */
    ESPIC_ST_BAD_OBJECT_MASK = 16, /* ref infMagSim_c.c:468,482,485 */
    ESPIC_ST_SAMPLER_STALL = 32,   /* ref infMagSim_c.c:589-607 (1000 rejections) */
    ESPIC_ST_FILTER_GRID = 128,    /* hole-tracking filter: the grid is too far from uniform for the block
                                    * expansion (see csrc/diag.cu); the response may be off by > 1e-8 */
    ESPIC_ST_EXCHANGE_TIMEOUT = 64 /* FATAL: a peer never published its densities within the exchange
                                    * timeout (multi-GPU exchange).  The potential of that step is NaN. */
};

#define ESPIC_MAX_RANKS 16
#define ESPIC_IPC_HANDLE_BYTES 64

typedef struct espic_ctx espic_ctx;

/* ---- context ---------------------------------------------------------------------------- */
int espic_abi_version(void);
/* Contexts may be created on any device of the process, several per process; every entry point
 * makes its context's device current for the duration of the call and restores the caller's. */
int espic_create(espic_ctx **out, int device);
void espic_destroy(espic_ctx *ctx);
/* cudaStream_t as void*; NULL = the legacy default stream.  All work of this context is
 * enqueued on it; nothing synchronises unless stated. */
int espic_set_stream(espic_ctx *ctx, void *cuda_stream);
const char *espic_last_error(espic_ctx *ctx);
/* Reads and clears the status word (synchronises the stream). */
int espic_status(espic_ctx *ctx, int *status_out);
/* Multiprocessor count and the persistent grid the mover uses (for bench bookkeeping). */
int espic_device_info(espic_ctx *ctx, int *sm_count, int *mover_grid, int *mover_block);
/* Cells of the shared-memory window the mover keeps per segment of the store when the grid is
 * too large for whole-grid tables (0 when the whole grid fits, i.e. up to ~19k points).  Drivers
 * use it to decide how often to call espic_sort_particles_coarse; a window of at least n_points-1
 * cells covers the grid completely and needs no sorting. */
int espic_mover_window_cells(espic_ctx *ctx, int n_points);
/* Kernels launched on this context since creation (bench's gpu_launches). */
long long espic_launch_count(espic_ctx *ctx);
/* Development aid: in a library built with -DESPIC_DEBUG_CYCLES=1 and with ESPIC_DEBUG_CTA=1 in the environment,
 * per-segment and per-CTA clock cycles of the last windowed mover launch (up to n values copied to out; returns
 * how many, 0 in a regular build; synchronises). */
int espic_debug_cta_cycles(espic_ctx *ctx, long long *out, int n);
/* Per-launch CUDA-event timing of the mover kernel: enable, run, then read the mean launch
 * duration in ms and the launch count (synchronises). */
int espic_timing_enable(espic_ctx *ctx, int on);
int espic_timing_read(espic_ctx *ctx, double *mean_ms, int *launches, double *total_ms);

/* ---- (A) device-pointer operators -------------------------------------------------------- */

/* Replaces move_particles_c (ref infMagSim_c.c:121-125), same parameter order and meaning.
 * particles = float[2][particle_storage_length] (row 0 positions, row 1 velocities);
 * density (n_points floats) is OVERWRITTEN; removed slots are pushed on empty_slots in
 * ascending slot order exactly as the reference's sequential loop does; current_empty_slot
 * is a DEVICE int (stack top), updated in place. */
int espic_move_particles(espic_ctx *ctx, const float *grid, const float *object_mask,
                         const float *potential, float dt, float charge_to_mass,
                         float background_density, int largest_index, float *particles,
                         float *density, int *empty_slots, int *current_empty_slot, float a_b,
                         int update_position, int periodic_particles, int largest_allowed_slot,
                         int n_points, int particle_storage_length);

/* Same, but largest_index is read from a DEVICE int at launch time (the injector updates it
 * on the device); max_slots bounds the launch (use particle_storage_length). */
int espic_move_particles_dev(espic_ctx *ctx, const float *grid, const float *object_mask,
                             const float *potential, float dt, float charge_to_mass,
                             float background_density, const int *largest_index_dev,
                             float *particles, float *density, int *empty_slots,
                             int *current_empty_slot, float a_b, int update_position,
                             int periodic_particles, int largest_allowed_slot, int n_points,
                             int particle_storage_length);

/* K6 (no reference counterpart: the reference never reorders its store): stable sort of the
 * live particles of slots 0..largest_index by cell index (the mover's (int)((x-z_min)/dz), after
 * the periodic fold when periodic_particles), ties in slot order; dead slots are dropped, the
 * survivors compacted to slots [0, n_live), x / v / injection tag moved together, everything
 * behind parked, and the free-slot stack rebuilt as the driver builds it (ref
 * infMagSim_script.py:486-489).  current_empty_slot and largest_index_dev (device ints) are
 * rewritten.  The permutation equals a stable argsort of the keys bit for bit. */
int espic_sort_particles(espic_ctx *ctx, const float *grid, int n_points, float *particles,
                         int particle_storage_length, int *particles_hist, int *empty_slots,
                         int n_stack, int *current_empty_slot, int *largest_index_dev,
                         int largest_index, int periodic_particles);

/* The same with the key coarsened to (cell index >> key_shift): particles end up grouped by
 * blocks of 2^key_shift cells, in slot order inside a block.  That is all the large-grid mover
 * needs (it keeps a window of the grid in shared memory per segment of the store) and it is one
 * radix pass instead of two on a 65536-cell grid.  With lookahead != 0 the cell is that of
 * fl(x + fl(v*lookahead)), folded into the domain (periodic) or clamped to the end cells: sorted by
 * where the particles will be half-way to the next sort, a segment stays together twice as long.
 * particles_hist may be NULL.  key_shift = 0, lookahead = 0 is espic_sort_particles. */
int espic_sort_particles_coarse(espic_ctx *ctx, const float *grid, int n_points, float *particles,
                                int particle_storage_length, int *particles_hist, int *empty_slots,
                                int n_stack, int *current_empty_slot, int *largest_index_dev,
                                int largest_index, int periodic_particles, int key_shift,
                                float lookahead);

/* v[i] -= v_b for i <= largest_index: the tail of initialize_mover
 * (ref infMagSim_cython.py:174-176; the subtraction runs in double there). */
int espic_shift_velocities(espic_ctx *ctx, float *particles, int particle_storage_length,
                           int largest_index, double v_b);

/* Replaces poisson_solve_c (ref infMagSim_c.c:356-358), same parameters.  Single-CTA
 * kernel; rows are built exactly as the reference builds them (Robin ends, object rows,
 * transparency blend, Boltzmann term, the periodic Sherman-Morrison variant including its
 * two quirks) and the Thomas recurrences run as fp64 block scans. */
int espic_poisson_solve(espic_ctx *ctx, const float *grid, const float *object_mask,
                        const float *charge, float debye_length, float *potential,
                        float object_potential, float object_transparency,
                        int boltzmann_electrons, int periodic_potential, int n_points);

/* Density end-node fix + charge density of the driver loop
 * (ref infMagSim_script.py:763-779,807): periodic -> fold the end nodes, else double them
 * (each species only if its fix flag is set); charge = ion - electron.  In place. */
int espic_prepare_charge(espic_ctx *ctx, float *ion_density, float *electron_density,
                         float *charge, int n_points, int periodic, int fix_ions,
                         int fix_electrons);

/* espic_prepare_charge + espic_poisson_solve as ONE kernel launch: the field part of a driver
 * timestep (ref infMagSim_script.py:763-779, 807, 831-836). */
int espic_field_solve(espic_ctx *ctx, const float *grid, const float *object_mask,
                      float *ion_density, float *electron_density, float *charge,
                      int periodic_particles, int fix_ions, int fix_electrons, float debye_length,
                      float *potential, float object_potential, float object_transparency,
                      int boltzmann_electrons, int periodic_potential, int n_points);

/* One whole driver timestep minus injection as ONE host call (ref infMagSim_script.py:763-779,
 * 807, 831-836, 929-950): espic_field_solve on the two densities, then espic_move_particles_dev
 * for the ions and for the electrons.  Everything is enqueued on the context's stream; nothing
 * synchronises.  A species with dt == 0 is frozen the way the driver freezes it (mover called
 * with dt 0).  species.potential == NULL means "the solved potential". */
typedef struct {
    float *particles;            /* float[2][particle_storage_length] */
    int *empty_slots;            /* int[largest_allowed_slot + 1] */
    int *current_empty_slot;     /* device int */
    int *largest_index;          /* device int */
    float *density;              /* float[n_points], overwritten by the push (acc == NULL) */
    const float *potential;      /* potential felt by this species, or NULL */
    int particle_storage_length;
    int largest_allowed_slot;
    float charge_to_mass;
    float dt;
    unsigned long long *acc;     /* NULL, or u64[n_points + 1] fixed-point deposit accumulators owned by the caller
                                  * (zero before the first step; the word behind the last node is bookkeeping of
                                  * the library): the pushes and injections of the step ADD their
                                  * deposit here (weights scaled by 2^24) and leave `density` alone; the next
                                  * step's field solve converts it (one float rounding per node, the same the
                                  * un-fused path makes) -- no conversion pass, no epilogue launch on the step path */
} espic_species;

typedef struct {
    const float *grid;
    const float *object_mask;
    float *potential;
    float *charge;
    espic_species ions;
    espic_species electrons;
    int n_points;
    float background_density;
    float a_b;
    float debye_length;
    float object_potential;
    float object_transparency;
    int periodic_particles;      /* mover wrap + end-node fold (ref infMagSim_script.py:145, 764-766) */
    int periodic_potential;      /* solver variant (ref infMagSim_script.py:146 ties it to the above) */
    int fix_ions;
    int fix_electrons;
    int exchange_parity;         /* < 0: single-rank field solve on ions.density / electrons.density.
                                  * 0/1: multi-GPU -- the field solve sums every rank's exchange slot of
                                  * this parity (espic_field_solve_exchange) into reduced_ion/_electron,
                                  * while the pushes deposit into ions.density / electrons.density
                                  * (the caller points them at the NEXT slot) */
    int exchange_step;
    float *reduced_ion_density;  /* only read when exchange_parity >= 0 */
    float *reduced_electron_density;
    int ion_acc_pending;         /* single rank, species.acc != NULL: 1 = this step's density of that species is still */
    int electron_acc_pending;    /* in its accumulators (the previous step left it there): the solve converts it into
                                  * species.density and clears them; 0 = species.density already holds it */
    /* wall injection after the pushes (ref infMagSim_script.py:986-1019) with the device sampler:
     * counts come from the host (expected flux + stochastic rounding, ref :965-975, 1005-1009) */
    int n_inject_ions;
    int n_inject_electrons;
    float v_th_ions, v_d_ions;           /* sampler parameters, ref :997-1000 */
    float v_th_electrons, v_d_electrons; /* ref :1016-1019 */
    float inject_dt;                     /* the driver's dt (partial-step placement), independent of a frozen species */
    int step_index;                      /* k: written into the injection tags */
    int inject_by_side;                  /* 0: particle l takes stack entry top-l, the reference's pop order (ref
                                          * infMagSim_cython.py:254-257, 282-283).  1 (stores kept sorted by position,
                                          * wall-bounded runs): the movers of this step push the slots they free
                                          * alternately from the two ends of their ascending order, outermost on top,
                                          * and the injection gives entrants at the right / left wall the entries at
                                          * even / odd distance from the top -- the same top n entries, each used once,
                                          * so a new particle lands among its neighbours in the store.  Slot numbers
                                          * only: physically invisible. */
    int *ion_tags;                       /* int[storage] injection histories; may be NULL when nothing is injected */
    int *electron_tags;
    float *density_snapshot;             /* NULL, or float[2][n_points]: the reduced, end-fixed ion and electron
                                          * densities the field solve used, copied right after the solve -- what the
                                          * reference driver stores per step (ref :763-779, 906-910) -- for the step
                                          * modes in which the pushes overwrite those arrays (species.acc == NULL) */
    /* hole tracking after the solve (ref :896-897): position_out == NULL switches it off */
    const float *fine_mesh;
    float *filter_scratch;               /* n_fine floats */
    float *hole_position_out;            /* one float */
    int n_fine;
    double tracking_dz;
    float tracking_width;
} espic_step_args;

int espic_pic_step(espic_ctx *ctx, const espic_step_args *args);

/* n_steps timesteps of the reference's main loop (ref infMagSim_script.py:756-1027) from one host call, for
 * runs whose step is too short to hide an interpreter between the launches (BASELINE configs[0] and [4]:
 * 1e5..1e6 particles, where the reference itself spends ~1e5 steps).  Step i is espic_pic_step with *first,
 * except: step_index = first_step + i; the injection counts come from the two host arrays (NULL = none);
 * hole_position_out = hole_positions + step_index while track_from <= step_index < track_until (NULL
 * otherwise, ref :896-897); from the second step on both accumulator planes count as pending.  Stream
 * ordered, nothing synchronises.  Single rank only (exchange_parity must be < 0).
 * use_graph != 0: the step is captured once as a CUDA graph and replayed -- one graph launch instead of the step's
 * 3..11 kernel launches; the first step of a call that needs a new graph runs ungraphed to size the scratch
 * buffers.  What varies from step to step stays out of the graph's arguments: the tracking result is indexed by a
 * device cursor, and an injecting stretch (at most 4096 steps per call) has its counts copied to the device once, its
 * launches sized for the stretch's maxima, and the step index, counts and sampler position read by the injection
 * kernels from a device control block that the last of them advances.  Results are bit-identical to stepping.
 * The graphs are kept in the context and reused by later calls with the same arguments. */
typedef struct espic_run_args {
    int n_steps;
    int first_step;
    const int *n_inject_ions;       /* host, n_steps entries, or NULL */
    const int *n_inject_electrons;  /* host, n_steps entries, or NULL */
    float *hole_positions;          /* device, indexed by step_index, or NULL */
    int track_from, track_until;
    int use_graph;
} espic_run_args;

int espic_pic_run(espic_ctx *ctx, const espic_step_args *first, const espic_run_args *run);

/* ---- multi-GPU density exchange fused into the field solve (replaces comm.Allreduce of the two
 * density grids, ref infMagSim_script.py:763, 773) ------------------------------------------------
 * One process per GPU of one box.  Each rank owns a small symmetric buffer in its own HBM
 * (two parity slots of [2][n_points] floats + a publication counter), exported to its peers
 * through CUDA IPC and read by them over NVLink.  Per step a rank's movers deposit into the
 * current slot, espic_exchange_publish() makes it visible, and espic_field_solve_exchange()
 * -- one kernel -- waits for every peer's counter, sums the P slots in rank order (so every
 * rank computes bit-identical fields), applies the end-node fix, forms the charge density and
 * solves.  No NCCL call, no extra launch, no host synchronisation on the step path.
 *   espic_exchange_create : allocate the local buffer, return its IPC handle (64 bytes)
 *   espic_exchange_open   : map the peers' buffers; handles = world x 64 bytes in rank order
 *   espic_exchange_slot   : device pointer of this rank's [2][n_points] slot of a given parity
 */
int espic_exchange_create(espic_ctx *ctx, int n_points, int world, int rank, void *handle_out);
int espic_exchange_open(espic_ctx *ctx, const void *handles);
int espic_exchange_slot(espic_ctx *ctx, int parity, float **slot_out);
/* Marks slot `parity` of this rank as holding the densities that enter step `step` (0-based;
 * call after the pushes and injections that produced them).  Stream-ordered. */
int espic_exchange_publish(espic_ctx *ctx, int parity, int step);
/* The same for the fused deposit (espic_species.acc): the given accumulator planes (either may be NULL) are
 * converted into this rank's slot rows -- ion row / electron row -- and cleared, then the slot is published. */
int espic_exchange_publish_fused(espic_ctx *ctx, int parity, int step, unsigned long long *ion_acc,
                                 unsigned long long *electron_acc, float background_density);
/* density[j] = (float)(acc[j] * 2^-24 / background_density), acc cleared: what the field solve does with a
 * pending fused deposit; for drivers that need the float density between steps (checkpoints, a change of
 * step mode). */
int espic_convert_density(espic_ctx *ctx, unsigned long long *acc, float *density, int n_points,
                          float background_density);
/* espic_field_solve over the rank-ordered sum of all ranks' slots of `parity` for `step`.
 * ion_density / electron_density (local, n_points each) receive the reduced, end-fixed densities. */
int espic_field_solve_exchange(espic_ctx *ctx, int parity, int step, const float *grid,
                               const float *object_mask, float *ion_density,
                               float *electron_density, float *charge, int periodic_particles,
                               int fix_ions, int fix_electrons, float debye_length,
                               float *potential, float object_potential,
                               float object_transparency, int boltzmann_electrons,
                               int periodic_potential, int n_points);

/* How long the fused field solve waits for a peer's publication before it gives up (default 30 s).
 * A timeout is fatal: ESPIC_ST_EXCHANGE_TIMEOUT is raised and the step's potential is NaN. */
int espic_exchange_set_timeout(espic_ctx *ctx, double seconds);

/* Replaces the body of inject_particles (ref infMagSim_cython.py:230-285) with both random
 * inputs supplied as device arrays: velocities[n_inject] (signed; the sign picks the wall)
 * and uniform01[n_inject] (partial-step fractions).  particles_hist = int[storage] tags.
 * current_empty_slot and largest_index are DEVICE ints, updated in place. */
int espic_inject_particles(espic_ctx *ctx, int n_inject, const float *grid, int n_points,
                           float dt, float background_density, const float *velocities,
                           const float *uniform01, float a_b, int k, int *particles_hist,
                           float *particles, int particle_storage_length, int *empty_slots,
                           int *current_empty_slot, int *largest_index, float *density);

/* Device replacements of sgenrand / draw_velocities_c / genrand (ref infMagSim_c.c:54-65,
 * 560-666, 67-103): the same flux-weighted drifting-Maxwellian rejection sampler (window,
 * side choice and acceptance test as in the reference, float conversion of ref :101) driven
 * by the counter-based Philox4x32-10 generator instead of the reference's sequential,
 * process-global MT19937, whose data-dependent consumption cannot be split across threads.
 * Same distribution, different stream: parity of the sampler is statistical (SURVEY.md s.7
 * hard part 7).  The key and call counter live in the context, so successive calls continue
 * the stream.  v_out / out are device arrays. */
int espic_seed(espic_ctx *ctx, unsigned long seed);
int espic_draw_velocities(espic_ctx *ctx, int n, float v_th, float v_d, float *v_out);
int espic_genrand(espic_ctx *ctx, int n, float *out);
/* Position of the sampler's stream for checkpoint / resume: state3 = {Philox key, number of
 * sampler calls consumed so far, seeded flag}.  Restoring it makes a resumed run draw exactly
 * what the uninterrupted run would have drawn. */
int espic_rng_get_state(espic_ctx *ctx, unsigned long long *state3);
int espic_rng_set_state(espic_ctx *ctx, const unsigned long long *state3);

/* Exhaustive device-side check of the mover's 3-operation quotient RN((x-z_lo)/dz) (see
 * quotient_dz in csrc/espic_internal.cuh) against the IEEE division instruction, over EVERY
 * float a in [0, span] for this grid.  counts_out (host, 2 values): [0] = inputs whose
 * truncated quotient (the cell index) differs -- must be 0; [1] = normal-range inputs whose
 * quotient differs in any bit -- must be 0.  Synchronises. */
int espic_selftest_quotient(espic_ctx *ctx, const float *grid, int n_points,
                            unsigned long long *counts_out);

/* ---- "next" rows: per-step diagnostics ---------------------------------------------------- */
/* Replaces electric_field_filter_c (ref infMagSim_c.c:668-683). Device arrays. */
int espic_field_filter(espic_ctx *ctx, int n_points, const float *grid, const float *data,
                       float L, int n_fine, const float *fine_mesh, float *res);
/* Hole position of the driver (ref infMagSim_script.py:746-753): np.gradient(potential, dz)
 * (float32 central differences, one-sided at the ends), the filter above and
 * fine_mesh[argmax]; scratch_res (n_fine floats) receives the filter response and
 * position_out (one device float) the position.  dz is the driver's Python float. */
int espic_hole_position(espic_ctx *ctx, int n_points, const float *grid,
                        const float *potential, double dz, float L, int n_fine,
                        const float *fine_mesh, float *scratch_res, float *position_out);
/* Replaces histogram2d_uniform_grid_c (ref infMagSim_c.c:543-558) directly on device arrays
 * xs[n_data], ys[n_data] (e.g. the two rows of a particle store: parked slots sit at 2*z_max,
 * outside any histogram range, which is what the driver's "occupied" mask removes,
 * ref infMagSim_script.py:866-867).  If tags != NULL only entries with tags[i] == tag_value
 * are counted (ref :880-881).  Counts are ADDED to hist[n_bins_x*n_bins_y]. */
int espic_phase_space_histogram(espic_ctx *ctx, const float *xs, const float *ys, int n_data,
                                const int *tags, int tag_value, float x_min, float step_x,
                                float y_min, float step_y, int n_bins_x, int n_bins_y,
                                int *hist);

/* ---- (B) reference-named host-pointer drop-ins ------------------------------------------ */
void move_particles_c(float *grid, float *object_mask, float *potential, float dt,
                      float charge_to_mass, float background_density, int largest_index,
                      float *particles, float *density, int *empty_slots,
                      int *current_empty_slot, float a_b, int update_position,
                      int periodic_particles, int largest_allowed_slot, int n_points,
                      int particle_storage_length); /* ref infMagSim_c.c:121-125 */
void move_particles_c_minimal(float *grid, float *object_mask, float *potential, float dt,
                              float charge_to_mass, float background_density, int largest_index,
                              float *particles, float *density, int *empty_slots,
                              int *current_empty_slot, int update_position,
                              int periodic_particles, int largest_allowed_slot, int n_points,
                              int particle_storage_length); /* ref infMagSim_c.c:231-235 (declared
                              cdef extern at infMagSim_cython.py:26-30, never called by the driver) */
void poisson_solve_c(float *grid, float *object_mask, float *charge, float debye_length,
                     float *potential, float object_potential, float object_transparency,
                     int boltzmann_electrons, int periodic_potential,
                     int n_points); /* ref infMagSim_c.c:356-358 */
void tridiagonal_solve_c(double *a, double *b, double *c, double *d, double *x,
                         int n); /* ref infMagSim_c.c:343-354; also modifies b and d, as there */
void draw_velocities_c(int n, float v_th, float v_d, float *v_array); /* ref :560 */
void sgenrand(unsigned long seed);                                    /* ref :54-56 */
float genrand(void);                                                  /* ref :67-69 */
void electric_field_filter_c(int n_points, float *grid, float *data_array, float L,
                             int n_fine_mesh, float *fine_mesh, float *Res); /* ref :668 */
void histogram2d_uniform_grid_c(float *X, float *Y, float x_min, float step_x, float y_min,
                                float step_y, int n_bins_x, int n_bins_y, int n_data,
                                int *hist); /* ref :543-544 */
/* The host-array movers page-lock the caller's particle arrays on first use (copies at full PCIe
 * speed, upload / push / download pipelined in chunks) and keep them registered, because the
 * reference driver passes the same arrays at every step.  Call this before freeing or reallocating
 * those arrays. */
void espic_host_release(void);
/* New symbol (the reference's injection body is Cython, infMagSim_cython.py:230-285):
 * host-pointer form of espic_inject_particles. */
void inject_particles_c(int n_inject, float *grid, int n_points, float dt,
                        float background_density, float *velocities, float *uniform01,
                        float a_b, int k, int *particles_hist, float *particles,
                        int particle_storage_length, int *empty_slots,
                        int *current_empty_slot, int *largest_index, float *density);

#ifdef __cplusplus
}
#endif
#endif /* ESPIC_B200_H */
