// C ABI of libespic_b200.so (see include/espic_b200.h): context management, the espic_*
// device-pointer operators, and the reference-named host-pointer drop-ins.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "espic_internal.cuh"

namespace espic {

int fail(espic_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int check_cuda(espic_ctx* ctx, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail(ctx, ESPIC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// Grow-only device scratch; new memory is zero-filled on the context's stream.
int ensure(espic_ctx* ctx, Scratch& s, size_t bytes, const char* what) {
    if (bytes <= s.bytes) return 0;
    if (s.ptr) {
        // the old buffer may still be in use by work queued on the stream
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return check_cuda(ctx, e, "cudaStreamSynchronize");
        cudaFree(s.ptr);
        s.ptr = nullptr;
        s.bytes = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    ctx->scratch_epoch++;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, ESPIC_E_NOMEM, "cudaMalloc(%zu bytes) for %s: %s", want, what, cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(p, 0, want, ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(p);
        return check_cuda(ctx, e, "cudaMemsetAsync");
    }
    s.ptr = p;
    s.bytes = want;
    return 0;
}

int ensure_acc(espic_ctx* ctx, int n_points) {
    if (n_points <= ctx->acc_len) return 0;
    if (ctx->acc) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return check_cuda(ctx, e, "cudaStreamSynchronize");
        cudaFree(ctx->acc);
        ctx->acc = nullptr;
        ctx->acc_len = 0;
    }
    const int len = n_points + 64;
    ctx->scratch_epoch++;
    cudaError_t e = cudaMalloc((void**)&ctx->acc, (size_t)len * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, ESPIC_E_NOMEM, "cudaMalloc for density accumulators: %s", cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(ctx->acc, 0, (size_t)len * sizeof(unsigned long long), ctx->stream);
    if (e != cudaSuccess) return check_cuda(ctx, e, "cudaMemsetAsync");
    ctx->acc_len = len;
    return 0;
}

static void free_scratch(Scratch& s) {
    if (s.ptr) cudaFree(s.ptr);
    s.ptr = nullptr;
    s.bytes = 0;
}

}  // namespace espic

using namespace espic;

// carve a sub-buffer out of the staging allocation
struct Carver {
    char* base;
    size_t off = 0;
    template <class T>
    T* take(size_t n) {
        T* p = reinterpret_cast<T*>(base + off);
        off += ((n * sizeof(T) + 255) / 256) * 256;
        return p;
    }
};
static size_t padded(size_t bytes) { return ((bytes + 255) / 256) * 256; }

extern "C" {

int espic_abi_version(void) { return ESPIC_ABI_VERSION; }

int espic_create(espic_ctx** out, int device) {
    if (!out) return ESPIC_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return ESPIC_E_NODEVICE;
    }
    if (device < 0 || device >= count) return ESPIC_E_ARG;
    // the context's device is made current only while an entry point runs (DeviceGuard); the
    // caller's current device is left as it was
    struct Restore {
        int prev = -1;
        Restore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
        ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore;
    if (cudaSetDevice(device) != cudaSuccess) return ESPIC_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ESPIC_E_CUDA;
    if (prop.major < 10) {
        fprintf(stderr, "espic_b200: device %d is sm_%d%d; this library contains sm_100a code only\n", device,
                prop.major, prop.minor);
        return ESPIC_E_NODEVICE;
    }
    espic_ctx* ctx = new espic_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaMalloc((void**)&ctx->status, 64) != cudaSuccess || cudaMemset(ctx->status, 0, 64) != cudaSuccess ||
        cudaMalloc((void**)&ctx->dev_int2, 64) != cudaSuccess || cudaMemset(ctx->dev_int2, 0, 64) != cudaSuccess) {
        delete ctx;
        return ESPIC_E_NOMEM;
    }
    *out = ctx;
    return ESPIC_OK;
}

void espic_destroy(espic_ctx* ctx) {
    DeviceGuard guard(ctx);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->acc) cudaFree(ctx->acc);
    if (ctx->status) cudaFree(ctx->status);
    if (ctx->dev_int2) cudaFree(ctx->dev_int2);
    free_scratch(ctx->staging);
    free_scratch(ctx->chunk_counts);
    free_scratch(ctx->range_offsets);
    free_scratch(ctx->dv_table);
    free_scratch(ctx->solve_ws);
    free_scratch(ctx->inject_ws);
    free_scratch(ctx->rng_ws);
    free_scratch(ctx->sort_ws);
    free_scratch(ctx->host_stage);
    free_scratch(ctx->filter_ws);
    free_scratch(ctx->debug_ws);
    free_scratch(ctx->run_ctl);
    for (auto& g : ctx->step_graph)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
    for (int r = 0; r < ctx->xch.world; ++r)
        if (r != ctx->xch.rank && ctx->xch.peer[r]) cudaIpcCloseMemHandle(ctx->xch.peer[r]);
    if (ctx->xch.local) cudaFree(ctx->xch.local);
    if (ctx->ev) {
        for (int i = 0; i < ctx->ev_cap; ++i) cudaEventDestroy(ctx->ev[i]);
        free(ctx->ev);
    }
    for (int i = 0; i < ctx->pipe_ev_cap; ++i) cudaEventDestroy(ctx->pipe_ev[i]);
    free(ctx->pipe_ev);
    for (int i = 0; i < 3; ++i)
        if (ctx->pipe_stream[i]) cudaStreamDestroy(ctx->pipe_stream[i]);
    delete ctx;
}

int espic_set_stream(espic_ctx* ctx, void* cuda_stream) {
    if (!ctx) return ESPIC_E_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;
    return ESPIC_OK;
}

const char* espic_last_error(espic_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int espic_status(espic_ctx* ctx, int* status_out) {
    DeviceGuard guard(ctx);
    if (!ctx || !status_out) return ESPIC_E_ARG;
    int st = 0;
    ESPIC_CUDA(ctx, cudaMemcpyAsync(&st, ctx->status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ESPIC_CUDA(ctx, cudaMemsetAsync(ctx->status, 0, sizeof(int), ctx->stream));
    ESPIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *status_out = st;
    return ESPIC_OK;
}

int espic_device_info(espic_ctx* ctx, int* sm_count, int* mover_grid, int* mover_block) {
    if (!ctx) return ESPIC_E_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (mover_grid) *mover_grid = ctx->mover_grid;
    if (mover_block) *mover_block = ctx->mover_block;
    return ESPIC_OK;
}

int espic_mover_window_cells(espic_ctx* ctx, int n_points) {
    DeviceGuard guard(ctx);
    if (!ctx || n_points < 2) return 0;
    return move_window_cells(ctx, n_points);
}

long long espic_launch_count(espic_ctx* ctx) { return ctx ? ctx->launches : -1; }

// Development aid: with ESPIC_DEBUG_CTA=1 in the environment the windowed mover records, per CTA, the clock cycles
// from entry to exit of its last launch; copies up to n of them to `out` (synchronises).  Returns the count.
int espic_debug_cta_cycles(espic_ctx* ctx, long long* out, int n) {
    DeviceGuard guard(ctx);
    if (!ctx || !out || n < 1 || !ctx->debug_ws.ptr) return 0;
    if (n > 2048) n = 2048;
    if (cudaMemcpyAsync(out, ctx->debug_ws.ptr, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return 0;
    return n;
}

int espic_timing_enable(espic_ctx* ctx, int on) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (on && !ctx->ev) {
        ctx->ev_cap = 8192;
        ctx->ev = (cudaEvent_t*)calloc((size_t)ctx->ev_cap, sizeof(cudaEvent_t));
        if (!ctx->ev) return fail(ctx, ESPIC_E_NOMEM, "event array");
        for (int i = 0; i < ctx->ev_cap; ++i) ESPIC_CUDA(ctx, cudaEventCreate(&ctx->ev[i]));
    }
    ctx->timing_on = on ? 1 : 0;
    ctx->ev_used = 0;
    return ESPIC_OK;
}

int espic_timing_read(espic_ctx* ctx, double* mean_ms, int* launches, double* total_ms) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    ESPIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0.;
    const int n = ctx->ev_used / 2;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        ESPIC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]));
        tot += ms;
    }
    if (mean_ms) *mean_ms = n ? tot / n : 0.;
    if (launches) *launches = n;
    if (total_ms) *total_ms = tot;
    ctx->ev_used = 0;
    return ESPIC_OK;
}

// ---- (A) device-pointer operators -------------------------------------------------------------

int espic_move_particles(espic_ctx* ctx, const float* grid, const float* object_mask, const float* potential,
                         float dt, float charge_to_mass, float background_density, int largest_index,
                         float* particles, float* density, int* empty_slots, int* current_empty_slot,
                         float a_b, int update_position, int periodic_particles, int largest_allowed_slot,
                         int n_points, int particle_storage_length) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (!grid || !object_mask || !potential || !particles || !density || !empty_slots || !current_empty_slot)
        return fail(ctx, ESPIC_E_ARG, "espic_move_particles: null pointer argument");
    return launch_move(ctx, grid, object_mask, potential, dt, charge_to_mass, background_density, largest_index,
                       nullptr, particles, density, empty_slots, current_empty_slot, a_b, update_position,
                       periodic_particles, largest_allowed_slot, n_points, particle_storage_length);
}

int espic_move_particles_dev(espic_ctx* ctx, const float* grid, const float* object_mask, const float* potential,
                             float dt, float charge_to_mass, float background_density,
                             const int* largest_index_dev, float* particles, float* density, int* empty_slots,
                             int* current_empty_slot, float a_b, int update_position, int periodic_particles,
                             int largest_allowed_slot, int n_points, int particle_storage_length) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (!grid || !object_mask || !potential || !particles || !density || !empty_slots || !current_empty_slot ||
        !largest_index_dev)
        return fail(ctx, ESPIC_E_ARG, "espic_move_particles_dev: null pointer argument");
    return launch_move(ctx, grid, object_mask, potential, dt, charge_to_mass, background_density, -1,
                       largest_index_dev, particles, density, empty_slots, current_empty_slot, a_b,
                       update_position, periodic_particles, largest_allowed_slot, n_points,
                       particle_storage_length);
}

int espic_selftest_quotient(espic_ctx* ctx, const float* grid, int n_points, unsigned long long* counts_out) {
    DeviceGuard guard(ctx);
    if (!ctx || !grid || !counts_out || n_points < 2) return ESPIC_E_ARG;
    return launch_selftest_quotient(ctx, grid, n_points, counts_out);
}

int espic_sort_particles(espic_ctx* ctx, const float* grid, int n_points, float* particles,
                         int particle_storage_length, int* particles_hist, int* empty_slots, int n_stack,
                         int* current_empty_slot, int* largest_index_dev, int largest_index, int periodic_particles) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (!grid || !particles || !particles_hist || !empty_slots || !current_empty_slot || !largest_index_dev)
        return fail(ctx, ESPIC_E_ARG, "espic_sort_particles: null pointer argument");
    return launch_sort(ctx, grid, n_points, particles, particle_storage_length, particles_hist, empty_slots, n_stack,
                       current_empty_slot, largest_index_dev, largest_index, periodic_particles, 0, 0.f);
}

int espic_sort_particles_coarse(espic_ctx* ctx, const float* grid, int n_points, float* particles,
                                int particle_storage_length, int* particles_hist, int* empty_slots, int n_stack,
                                int* current_empty_slot, int* largest_index_dev, int largest_index,
                                int periodic_particles, int key_shift, float lookahead) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (!grid || !particles || !empty_slots || !current_empty_slot || !largest_index_dev)
        return fail(ctx, ESPIC_E_ARG, "espic_sort_particles_coarse: null pointer argument");
    return launch_sort(ctx, grid, n_points, particles, particle_storage_length, particles_hist, empty_slots, n_stack,
                       current_empty_slot, largest_index_dev, largest_index, periodic_particles, key_shift, lookahead);
}

int espic_shift_velocities(espic_ctx* ctx, float* particles, int particle_storage_length, int largest_index,
                           double v_b) {
    DeviceGuard guard(ctx);
    if (!ctx || !particles) return ESPIC_E_ARG;
    return launch_shift_velocities(ctx, particles, particle_storage_length, largest_index, v_b);
}

int espic_poisson_solve(espic_ctx* ctx, const float* grid, const float* object_mask, const float* charge,
                        float debye_length, float* potential, float object_potential, float object_transparency,
                        int boltzmann_electrons, int periodic_potential, int n_points) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (!grid || !object_mask || !charge || !potential)
        return fail(ctx, ESPIC_E_ARG, "espic_poisson_solve: null pointer argument");
    return launch_poisson(ctx, grid, object_mask, charge, debye_length, potential, object_potential,
                          object_transparency, boltzmann_electrons, periodic_potential, n_points);
}

int espic_prepare_charge(espic_ctx* ctx, float* ion_density, float* electron_density, float* charge, int n_points,
                         int periodic, int fix_ions, int fix_electrons) {
    DeviceGuard guard(ctx);
    if (!ctx || !ion_density || !electron_density || !charge) return ESPIC_E_ARG;
    return launch_prepare_charge(ctx, ion_density, electron_density, charge, n_points, periodic, fix_ions,
                                 fix_electrons);
}

int espic_field_solve(espic_ctx* ctx, const float* grid, const float* object_mask, float* ion_density,
                      float* electron_density, float* charge, int periodic_particles, int fix_ions,
                      int fix_electrons, float debye_length, float* potential, float object_potential,
                      float object_transparency, int boltzmann_electrons, int periodic_potential, int n_points) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (!grid || !object_mask || !ion_density || !electron_density || !charge || !potential)
        return fail(ctx, ESPIC_E_ARG, "espic_field_solve: null pointer argument");
    return launch_field_solve(ctx, grid, object_mask, ion_density, electron_density, charge, periodic_particles,
                              fix_ions, fix_electrons, debye_length, potential, object_potential,
                              object_transparency, boltzmann_electrons, periodic_potential, n_points);
}

int espic_exchange_create(espic_ctx* ctx, int n_points, int world, int rank, void* handle_out) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    return exchange_create(ctx, n_points, world, rank, handle_out);
}

int espic_exchange_open(espic_ctx* ctx, const void* handles) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    return exchange_open(ctx, handles);
}

int espic_exchange_slot(espic_ctx* ctx, int parity, float** slot_out) {
    DeviceGuard guard(ctx);
    if (!ctx || !slot_out || !ctx->xch.local) return ESPIC_E_ARG;
    *slot_out = static_cast<float*>(ctx->xch.local) + (size_t)(parity & 1) * ctx->xch.slot_floats;
    return ESPIC_OK;
}

int espic_exchange_publish(espic_ctx* ctx, int parity, int step) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    return exchange_publish(ctx, parity, step);
}

int espic_exchange_publish_fused(espic_ctx* ctx, int parity, int step, unsigned long long* ion_acc,
                                 unsigned long long* electron_acc, float background_density) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    return exchange_publish(ctx, parity, step, ion_acc, electron_acc, background_density);
}

int espic_convert_density(espic_ctx* ctx, unsigned long long* acc, float* density, int n_points,
                          float background_density) {
    DeviceGuard guard(ctx);
    if (!ctx || !acc || !density) return ESPIC_E_ARG;
    return launch_convert_density(ctx, acc, density, n_points, background_density);
}

int espic_field_solve_exchange(espic_ctx* ctx, int parity, int step, const float* grid, const float* object_mask,
                               float* ion_density, float* electron_density, float* charge, int periodic_particles,
                               int fix_ions, int fix_electrons, float debye_length, float* potential,
                               float object_potential, float object_transparency, int boltzmann_electrons,
                               int periodic_potential, int n_points) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (!grid || !object_mask || !ion_density || !electron_density || !charge || !potential)
        return fail(ctx, ESPIC_E_ARG, "espic_field_solve_exchange: null pointer argument");
    return launch_field_solve_exchange(ctx, parity, step, grid, object_mask, ion_density, electron_density, charge,
                                       periodic_particles, fix_ions, fix_electrons, debye_length, potential,
                                       object_potential, object_transparency, boltzmann_electrons,
                                       periodic_potential, n_points);
}

// One timestep on ctx->stream.  track_cursor: device int indexing hole_position_out (step graphs), or null.
static int pic_step_impl(espic_ctx* ctx, const espic_step_args* a, int* track_cursor, int* inject_ctl = nullptr) {
    if (!a->grid || !a->object_mask || !a->potential || !a->charge) return fail(ctx, ESPIC_E_ARG, "espic_pic_step: null grid pointer");
    const espic_species* sp[2] = {&a->ions, &a->electrons};
    for (int i = 0; i < 2; ++i)
        if (!sp[i]->particles || !sp[i]->empty_slots || !sp[i]->current_empty_slot || !sp[i]->largest_index ||
            !sp[i]->density)
            return fail(ctx, ESPIC_E_ARG, "espic_pic_step: null species pointer");
    int rc;
    unsigned long long* acc[2] = {a->ions.acc, a->electrons.acc};
    if ((acc[0] == nullptr) != (acc[1] == nullptr))
        return fail(ctx, ESPIC_E_ARG, "espic_pic_step: give both species an accumulator plane or neither");
    // removals pushed directly by the periodic movers (mover_common.cuh, direct_push): counted in the word behind
    // the last node of the species' plane, put into ascending order by the next solve
    StackFix fix{};
    int* pushed[2] = {nullptr, nullptr};
    if (acc[0]) {
        for (int i = 0; i < 2; ++i) {
            pushed[i] = reinterpret_cast<int*>(acc[i] + a->n_points);
            fix.pushed[i] = pushed[i];
            fix.empty_slots[i] = sp[i]->empty_slots;
            fix.top[i] = sp[i]->current_empty_slot;
            fix.largest_allowed_slot[i] = sp[i]->largest_allowed_slot;
        }
    }
    if (a->exchange_parity >= 0) {
        if (!a->reduced_ion_density || !a->reduced_electron_density)
            return fail(ctx, ESPIC_E_ARG, "espic_pic_step: exchange mode needs the reduced density buffers");
        rc = launch_field_solve_exchange(ctx, a->exchange_parity, a->exchange_step, a->grid, a->object_mask,
                                         a->reduced_ion_density, a->reduced_electron_density, a->charge,
                                         a->periodic_particles, a->fix_ions, a->fix_electrons, a->debye_length,
                                         a->potential, a->object_potential, a->object_transparency, 0,
                                         a->periodic_potential, a->n_points, acc[0] ? &fix : nullptr);
    } else {
        unsigned long long* pend[2] = {a->ion_acc_pending ? acc[0] : nullptr, a->electron_acc_pending ? acc[1] : nullptr};
        if (a->n_points > 8192) {
            // beyond 8192 points the solve walks several rows per thread on 8 CTAs; converting 2 x n_points
            // accumulators there costs more (65537 points: 37 -> 59 us) than one wide launch in front of it
            if ((rc = launch_convert_density2(ctx, pend[0], sp[0]->density, pend[1], sp[1]->density, a->n_points,
                                              a->background_density)))
                return rc;
            pend[0] = pend[1] = nullptr;
        }
        rc = launch_field_solve(ctx, a->grid, a->object_mask, a->ions.density, a->electrons.density, a->charge,
                                a->periodic_particles, a->fix_ions, a->fix_electrons, a->debye_length, a->potential,
                                a->object_potential, a->object_transparency, 0, a->periodic_potential, a->n_points,
                                pend[0], pend[1], a->background_density, acc[0] ? &fix : nullptr);
    }
    if (rc) return rc;
    if (a->density_snapshot) {  // the densities the solve used, before the pushes overwrite them
        const float* src[2] = {a->exchange_parity >= 0 ? a->reduced_ion_density : a->ions.density,
                               a->exchange_parity >= 0 ? a->reduced_electron_density : a->electrons.density};
        for (int i = 0; i < 2; ++i)
            ESPIC_CUDA(ctx, cudaMemcpyAsync(a->density_snapshot + (size_t)i * a->n_points, src[i],
                                            (size_t)a->n_points * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    for (int i = 0; i < 2; ++i) {
        const espic_species* s = sp[i];
        // large grids: the ions' call builds both species' kick tables, the electrons' call finds its own ready
        const espic_species* o = sp[1];
        const DvPlan plan{i == 0 ? 1 : 2, o->potential ? o->potential : a->potential, o->charge_to_mass, a->a_b, o->dt};
        rc = launch_move(ctx, a->grid, a->object_mask, s->potential ? s->potential : a->potential, s->dt,
                         s->charge_to_mass, a->background_density, -1, s->largest_index, s->particles, s->density,
                         s->empty_slots, s->current_empty_slot, a->a_b, 1, a->periodic_particles, s->largest_allowed_slot,
                         a->n_points, s->particle_storage_length, 0, 0, 1, acc[i], pushed[i],
                         (a->inject_by_side && !a->periodic_particles) ? 1 : 0, &plan);
        if (rc) return rc;
    }
    if (a->hole_position_out) {
        if (!a->fine_mesh || !a->filter_scratch) return fail(ctx, ESPIC_E_ARG, "espic_pic_step: tracking buffers missing");
        rc = launch_hole_position(ctx, a->n_points, a->grid, a->potential, a->tracking_dz, a->tracking_width,
                                  a->n_fine, a->fine_mesh, a->filter_scratch, a->hole_position_out, track_cursor);
        if (rc) return rc;
    }
    const int n_inj[2] = {a->n_inject_ions, a->n_inject_electrons};
    const float v_th[2] = {a->v_th_ions, a->v_th_electrons}, v_d[2] = {a->v_d_ions, a->v_d_electrons};
    int* tags[2] = {a->ion_tags, a->electron_tags};
    // inject_ctl (replayed step graphs): n_inject_* are the upper bounds of the stretch; the kernels take the step's
    // own counts, the tag and the sampler's call counter from the control block, and the species injected last
    // moves the block on to the next step
    const int last_injecting = n_inj[1] > 0 ? 1 : 0;
    if (n_inj[0] > 0 && n_inj[1] > 0 && tags[0] && tags[1] && acc[0] && acc[1]) {   // fused deposit: a plane per species
        const char* e = getenv("ESPIC_PAIRED_INJECT");   // tuning / test knob; 0 = one sample and one commit launch per species
        if (!e || atoi(e) != 0) {
            InjectPairSpecies ps[2];
            for (int i = 0; i < 2; ++i) {
                const espic_species* s = sp[i];
                ps[i] = InjectPairSpecies{n_inj[i], v_th[i], v_d[i], tags[i], s->particles, s->particle_storage_length,
                                          s->empty_slots, s->current_empty_slot, s->largest_index, s->density, acc[i]};
            }
            return launch_inject_pair(ctx, ps, a->grid, a->n_points, a->inject_dt, a->background_density, a->a_b,
                                      a->step_index, a->inject_by_side, inject_ctl, kCtlMaxSteps);
        }
    }
    for (int i = 0; i < 2; ++i) {
        if (n_inj[i] <= 0) continue;
        const InjectCtl ic{inject_ctl, i, kCtlMaxSteps, i == last_injecting ? 1 : 0};
        const InjectCtl* icp = inject_ctl ? &ic : nullptr;
        if (!tags[i]) return fail(ctx, ESPIC_E_ARG, "espic_pic_step: injection needs the tag array");
        rc = ensure(ctx, ctx->rng_ws, (size_t)2 * ((n_inj[i] + 63) & ~63) * sizeof(float), "injection random inputs");
        if (rc) return rc;
        float* uni = (float*)ctx->rng_ws.ptr;
        float* vel = uni + ((n_inj[i] + 63) & ~63);
        // same draw order as the reference wrapper: fractions first, then velocities (ref infMagSim_cython.py:245-248);
        // both draws and the bad-index classification in one launch, placement + deposit + stack pop in a second
        if ((rc = rng_sample_for_injection(ctx, n_inj[i], v_th[i], v_d[i], vel, uni, a->grid, a->n_points, a->inject_dt,
                                           a->background_density, a->a_b, icp)))
            return rc;
        const espic_species* s = sp[i];
        rc = launch_inject(ctx, n_inj[i], a->grid, a->n_points, a->inject_dt, a->background_density, vel, uni, a->a_b,
                           a->step_index, tags[i], s->particles, s->particle_storage_length, s->empty_slots,
                           s->current_empty_slot, s->largest_index, s->density, true, acc[i], a->inject_by_side, icp);
        if (rc) return rc;
    }
    return ESPIC_OK;
}

int espic_pic_step(espic_ctx* ctx, const espic_step_args* a) {
    DeviceGuard guard(ctx);
    if (!ctx || !a) return ESPIC_E_ARG;
    return pic_step_impl(ctx, a, nullptr);
}

__global__ void k_set_int(int* p, int v) { *p = v; }
__global__ void k_set_ctl(int* ctl, int k, int i, unsigned long long calls) {
    ctl[0] = k;
    ctl[1] = i;
    *reinterpret_cast<unsigned long long*>(ctl + 2) = calls;
}

static void drop_step_graph(espic_ctx::StepGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
}

// Captures one step with the given arguments on the side stream (nothing executes) and instantiates it.
static int capture_step(espic_ctx* ctx, const espic_step_args* a, int which, int* inject_ctl) {
    espic_ctx::StepGraph& g = ctx->step_graph[which];
    drop_step_graph(g);
    if (!ctx->capture_stream)
        ESPIC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking));
    cudaStream_t user = ctx->stream;
    const long long before = ctx->launches;
    const unsigned long long epoch = ctx->scratch_epoch;
    ESPIC_CUDA(ctx, cudaStreamBeginCapture(ctx->capture_stream, cudaStreamCaptureModeThreadLocal));
    ctx->stream = ctx->capture_stream;
    int rc = pic_step_impl(ctx, a, which ? ctx->status + 12 : nullptr, inject_ctl);
    ctx->stream = user;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->capture_stream, &graph);
    const int n_launch = (int)(ctx->launches - before);
    ctx->launches = before;   // nothing ran
    if (rc || e != cudaSuccess || !graph || ctx->scratch_epoch != epoch) {
        // a buffer grew while capturing (the graph may hold the old pointer) or the step refused its arguments
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc ? rc : fail(ctx, ESPIC_E_CUDA, "espic_pic_run: capturing the step failed (%s)",
                              e != cudaSuccess ? cudaGetErrorString(e) : "scratch reallocated");
    }
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        g.exec = nullptr;
        return check_cuda(ctx, e, "cudaGraphInstantiate");
    }
    memcpy(&g.key, a, sizeof *a);
    g.epoch = epoch;
    g.launches = n_launch;
    return ESPIC_OK;
}

int espic_pic_run(espic_ctx* ctx, const espic_step_args* first, const espic_run_args* r) {
    DeviceGuard guard(ctx);
    if (!ctx || !first || !r) return ESPIC_E_ARG;
    if (r->n_steps < 0) return fail(ctx, ESPIC_E_ARG, "espic_pic_run: negative step count");
    if (first->exchange_parity >= 0)
        return fail(ctx, ESPIC_E_ARG, "espic_pic_run: single-rank stepping only (the density exchange publishes between steps)");
    const bool fused = first->ions.acc != nullptr;
    int warmed[2] = {0, 0};
    bool graphs = r->use_graph != 0 && !ctx->timing_on;
    // injecting stretches: the counts go to the device once, the launches are sized for the stretch's maxima
    // (rounded up, so that the captured geometry survives from stretch to stretch)
    int nmax[2] = {0, 0};
    for (int i = 0; i < r->n_steps; ++i) {
        if (r->n_inject_ions && r->n_inject_ions[i] > nmax[0]) nmax[0] = r->n_inject_ions[i];
        if (r->n_inject_electrons && r->n_inject_electrons[i] > nmax[1]) nmax[1] = r->n_inject_electrons[i];
    }
    const bool injecting = nmax[0] > 0 || nmax[1] > 0;
    int* ctl = nullptr;
    bool ctl_set = false;
    if (graphs && injecting) {
        if (r->n_steps > kCtlMaxSteps) {
            graphs = false;
        } else {
            for (int s = 0; s < 2; ++s) {
                if (nmax[s] <= 0) continue;
                nmax[s] = (nmax[s] + 1023) & ~1023;
                // a graph captured for a slightly larger bound serves this stretch too: no recapture because the
                // stretch's maximum happened to fall on the other side of a multiple of 1024
                for (const auto& g : ctx->step_graph) {
                    const int have = s ? g.key.n_inject_electrons : g.key.n_inject_ions;
                    if (g.exec && have >= nmax[s] && have <= nmax[s] + nmax[s] / 4 + 4096) nmax[s] = have;
                }
            }
            const int big = nmax[0] > nmax[1] ? nmax[0] : nmax[1];
            int rc = ensure(ctx, ctx->run_ctl, (size_t)(16 + 2 * kCtlMaxSteps) * sizeof(int), "step control block");
            // sized for the paired launch (both species' random inputs and counters side by side)
            if (!rc) rc = ensure(ctx, ctx->rng_ws, (size_t)4 * ((big + 63) & ~63) * sizeof(float), "injection random inputs");
            if (!rc) rc = ensure(ctx, ctx->inject_ws, (32 + 2 * (32 + (size_t)big / 128)) * sizeof(int), "inject workspace");
            if (rc) return rc;
            ctl = (int*)ctx->run_ctl.ptr;
            const int* src[2] = {r->n_inject_ions, r->n_inject_electrons};
            for (int s = 0; s < 2; ++s) {
                if (src[s])
                    ESPIC_CUDA(ctx, cudaMemcpyAsync(ctl + 16 + s * kCtlMaxSteps, src[s], (size_t)r->n_steps * sizeof(int),
                                                    cudaMemcpyHostToDevice, ctx->stream));
                else
                    ESPIC_CUDA(ctx, cudaMemsetAsync(ctl + 16 + s * kCtlMaxSteps, 0, (size_t)r->n_steps * sizeof(int), ctx->stream));
            }
        }
    }
    for (int i = 0; i < r->n_steps; ++i) {
        const int k = r->first_step + i;
        const int n_i = r->n_inject_ions ? r->n_inject_ions[i] : 0, n_e = r->n_inject_electrons ? r->n_inject_electrons[i] : 0;
        espic_step_args a;
        memcpy(&a, first, sizeof a);
        a.step_index = k;
        a.n_inject_ions = n_i;
        a.n_inject_electrons = n_e;
        if (i > 0 && fused) a.ion_acc_pending = a.electron_acc_pending = 1;   // every step leaves its deposit pending
        const bool track = r->hole_positions && k >= r->track_from && k < r->track_until;
        a.hole_position_out = track ? r->hole_positions + k : nullptr;
        int rc;
        if (!graphs) {
            if ((rc = pic_step_impl(ctx, &a, nullptr))) return rc;
            continue;
        }
        // the replayed graph indexes the history through the device cursor and takes the injection's per-step values
        // from the control block; nothing in its arguments varies from step to step
        espic_step_args plain;
        memcpy(&plain, &a, sizeof a);
        a.step_index = 0;
        a.hole_position_out = track ? r->hole_positions : nullptr;
        a.n_inject_ions = nmax[0];
        a.n_inject_electrons = nmax[1];
        espic_ctx::StepGraph& g = ctx->step_graph[track ? 1 : 0];
        if (!g.exec || g.epoch != ctx->scratch_epoch || memcmp(&g.key, &a, sizeof a) != 0) {
            if (!warmed[track]) {
                // first an ordinary step: it sizes the scratch buffers and the per-kernel launch configuration
                if ((rc = pic_step_impl(ctx, &plain, nullptr))) return rc;
                warmed[track] = 1;
                continue;
            }
            if (capture_step(ctx, &a, track ? 1 : 0, injecting ? ctl : nullptr)) {   // not capturable here: carry on launch by launch
                graphs = false;
                if ((rc = pic_step_impl(ctx, &plain, nullptr))) return rc;
                continue;
            }
        }
        if (track && ctx->track_cursor_host != k) {
            k_set_int<<<1, 1, 0, ctx->stream>>>(ctx->status + 12, k);
            ctx->launches++;
        }
        if (injecting && (!ctl_set || ctx->ctl_host_k != k)) {
            if (!ctx->rng_seeded) rng_seed(ctx, 4357ul);
            k_set_ctl<<<1, 1, 0, ctx->stream>>>(ctl, k, i, ctx->rng_consumed);
            ctx->launches++;
            ctl_set = true;
        }
        ESPIC_CUDA(ctx, cudaGraphLaunch(g.exec, ctx->stream));
        ctx->launches += g.launches;
        if (track) ctx->track_cursor_host = k + 1;
        if (injecting) {
            ctx->ctl_host_k = k + 1;
            ctx->rng_consumed += (n_i > 0 ? 2ull : 0ull) + (n_e > 0 ? 2ull : 0ull);
        }
    }
    return ESPIC_OK;
}

int espic_inject_particles(espic_ctx* ctx, int n_inject, const float* grid, int n_points, float dt,
                           float background_density, const float* velocities, const float* uniform01, float a_b,
                           int k, int* particles_hist, float* particles, int particle_storage_length,
                           int* empty_slots, int* current_empty_slot, int* largest_index, float* density) {
    DeviceGuard guard(ctx);
    if (!ctx) return ESPIC_E_ARG;
    if (n_inject > 0 && (!grid || !velocities || !uniform01 || !particles_hist || !particles || !empty_slots ||
                         !current_empty_slot || !largest_index || !density))
        return fail(ctx, ESPIC_E_ARG, "espic_inject_particles: null pointer argument");
    return launch_inject(ctx, n_inject, grid, n_points, dt, background_density, velocities, uniform01, a_b, k,
                         particles_hist, particles, particle_storage_length, empty_slots, current_empty_slot,
                         largest_index, density);
}

int espic_seed(espic_ctx* ctx, unsigned long seed) {
    DeviceGuard guard(ctx);
    return ctx ? rng_seed(ctx, seed) : ESPIC_E_ARG;
}

int espic_draw_velocities(espic_ctx* ctx, int n, float v_th, float v_d, float* v_out) {
    DeviceGuard guard(ctx);
    if (!ctx || (n > 0 && !v_out)) return ESPIC_E_ARG;
    return rng_draw_velocities(ctx, n, v_th, v_d, v_out);
}

int espic_genrand(espic_ctx* ctx, int n, float* out) {
    DeviceGuard guard(ctx);
    if (!ctx || (n > 0 && !out)) return ESPIC_E_ARG;
    return rng_uniforms(ctx, n, out);
}

int espic_rng_get_state(espic_ctx* ctx, unsigned long long* state3) {
    if (!ctx || !state3) return ESPIC_E_ARG;
    state3[0] = ctx->rng_key;
    state3[1] = ctx->rng_consumed;
    state3[2] = (unsigned long long)ctx->rng_seeded;
    return ESPIC_OK;
}

int espic_rng_set_state(espic_ctx* ctx, const unsigned long long* state3) {
    if (!ctx || !state3) return ESPIC_E_ARG;
    ctx->rng_key = state3[0];
    ctx->rng_consumed = state3[1];
    ctx->rng_seeded = state3[2] != 0ull;
    return ESPIC_OK;
}

int espic_exchange_set_timeout(espic_ctx* ctx, double seconds) {
    if (!ctx || !(seconds > 0.)) return ESPIC_E_ARG;
    ctx->exchange_timeout_s = seconds;
    return ESPIC_OK;
}

int espic_field_filter(espic_ctx* ctx, int n_points, const float* grid, const float* data, float L, int n_fine,
                       const float* fine_mesh, float* res) {
    DeviceGuard guard(ctx);
    if (!ctx || !grid || !data || !fine_mesh || !res) return ESPIC_E_ARG;
    return launch_field_filter(ctx, n_points, grid, data, L, n_fine, fine_mesh, res);
}

int espic_hole_position(espic_ctx* ctx, int n_points, const float* grid, const float* potential, double dz,
                        float L, int n_fine, const float* fine_mesh, float* scratch_res, float* position_out) {
    DeviceGuard guard(ctx);
    if (!ctx || !grid || !potential || !fine_mesh || !scratch_res || !position_out) return ESPIC_E_ARG;
    return launch_hole_position(ctx, n_points, grid, potential, dz, L, n_fine, fine_mesh, scratch_res,
                                position_out);
}

int espic_phase_space_histogram(espic_ctx* ctx, const float* xs, const float* ys, int n_data, const int* tags,
                                int tag_value, float x_min, float step_x, float y_min, float step_y,
                                int n_bins_x, int n_bins_y, int* hist) {
    DeviceGuard guard(ctx);
    if (!ctx || !xs || !ys || !hist) return ESPIC_E_ARG;
    return launch_histogram(ctx, xs, ys, n_data, tags, tag_value, x_min, step_x, y_min, step_y, n_bins_x,
                            n_bins_y, hist);
}

// ---- (B) reference-named host-pointer drop-ins -------------------------------------------------
// One process-wide context on the current device, created on first use.  These functions
// return void like the reference's; any failure is fatal and loud (no CPU fallback exists).

static espic_ctx* g_default = nullptr;
static std::mutex g_mutex;

static espic_ctx* default_ctx() {
    if (!g_default) {
        int dev = 0;
        if (const char* s = getenv("ESPIC_DEVICE")) dev = atoi(s);
        int rc = espic_create(&g_default, dev);
        if (rc) {
            fprintf(stderr, "espic_b200: cannot create a context on CUDA device %d (code %d); "
                            "this library has no CPU path\n", dev, rc);
            abort();
        }
    }
    return g_default;
}

static void die(espic_ctx* ctx, const char* where, int rc) {
    fprintf(stderr, "espic_b200: %s failed (code %d): %s\n", where, rc, espic_last_error(ctx));
    abort();
}

// The reference's driver passes the same numpy arrays at every step.  Page-locking them once lets
// the copies below run at PCIe speed instead of through the driver's pageable bounce buffers.
// Registration failures (foreign allocators, read-only pages, ...) simply leave the range pageable.
struct PinnedRange {
    const void* ptr;
    size_t bytes;
};
static PinnedRange g_pinned[32];
static int g_n_pinned = 0;

// Page-locks [ptr, ptr + bytes) of a caller's array once so that the copies run at full PCIe speed
// (a pageable cudaMemcpyAsync is staged through a driver buffer).  `capacity` is the size of the
// caller's row: when the used part grows (injection raises largest_index) the registration is
// replaced by a larger one with headroom, never extended piecemeal -- a copy must not straddle
// registered and unregistered pages.
static void pin_host_range(const void* ptr, size_t bytes, size_t capacity) {
    size_t min_bytes = 8u << 20;  // small arrays are not worth a registration
    if (const char* e = getenv("ESPIC_PIN_MIN_BYTES")) min_bytes = (size_t)atoll(e);  // test knob
    if (!ptr || bytes < min_bytes) return;
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("ESPIC_PIN_HOST");
        enabled = e ? atoi(e) : 1;
    }
    if (!enabled) return;
    int slot = -1;
    for (int i = 0; i < g_n_pinned; ++i)
        if (g_pinned[i].ptr == ptr) slot = i;
    if (slot >= 0 && g_pinned[slot].bytes >= bytes) return;
    if (slot >= 0) {  // grown past the registered part
        cudaHostUnregister(const_cast<void*>(ptr));
        cudaGetLastError();
        g_pinned[slot] = g_pinned[--g_n_pinned];
    }
    if (g_n_pinned >= 32) return;
    size_t want = bytes + bytes / 4;
    if (want > capacity) want = capacity;
    if (want < bytes) want = bytes;
    if (cudaHostRegister(const_cast<void*>(ptr), want, cudaHostRegisterDefault) == cudaSuccess) {
        g_pinned[g_n_pinned++] = PinnedRange{ptr, want};
    } else {
        cudaGetLastError();  // leave it pageable
    }
}

static bool is_pinned(const void* ptr, size_t bytes) {
    for (int i = 0; i < g_n_pinned; ++i)
        if (g_pinned[i].ptr == ptr && g_pinned[i].bytes >= bytes) return true;
    return false;
}

static void unpin_all() {
    for (int i = 0; i < g_n_pinned; ++i) {
        cudaHostUnregister(const_cast<void*>(g_pinned[i].ptr));
        cudaGetLastError();
    }
    g_n_pinned = 0;
}

static void report_status(espic_ctx* ctx) {
    int st = 0;
    if (espic_status(ctx, &st)) return;
    // same wording as the reference's printf()s, once per call instead of once per particle
    if (st & ESPIC_ST_BAD_LEFT_INDEX) printf("bad left_index\n");
    if (st & ESPIC_ST_STACK_OVERFLOW) printf("bad current_empty_slot\n");
    if (st & ESPIC_ST_STACK_EMPTY) printf("no empty slots\n");
    if (st & ESPIC_ST_BAD_INJECT_INDEX) printf("bad left_index:\n");
    if (st & ESPIC_ST_BAD_OBJECT_MASK) printf("bad object_mask\n");
    if (st & ESPIC_ST_SAMPLER_STALL) printf("Warning: method draw_velocities_c doesn't converge\n");
}

#define H2D(dst, src, n, T) \
    if ((rc = check_cuda(ctx, cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(T), cudaMemcpyHostToDevice, ctx->stream), "H2D"))) die(ctx, __func__, rc)
#define D2H(dst, src, n, T) \
    if ((rc = check_cuda(ctx, cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream), "D2H"))) die(ctx, __func__, rc)

}  // extern "C"

// Host-array mover shared by move_particles_c and move_particles_c_minimal.
static void move_host_arrays(const char* who, float* grid, float* object_mask, float* potential, float dt,
                             float charge_to_mass, float background_density, int largest_index, float* particles,
                             float* density, int* empty_slots, int* current_empty_slot, float a_b,
                             int update_position, int periodic_particles, int largest_allowed_slot, int n_points,
                             int particle_storage_length, int minimal) {
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    const size_t S = (size_t)particle_storage_length, n = (size_t)n_points;
    const size_t n_stack = (size_t)largest_allowed_slot + 1;
    const size_t need = 4 * padded(n * 4) + padded(2 * S * 4) + padded(n_stack * 4) + 256;
    if ((rc = ensure(ctx, ctx->host_stage, need, "host staging"))) die(ctx, who, rc);
    Carver cv{(char*)ctx->host_stage.ptr};
    float* d_grid = cv.take<float>(n);
    float* d_mask = cv.take<float>(n);
    float* d_phi = cv.take<float>(n);
    float* d_den = cv.take<float>(n);
    float* d_part = cv.take<float>(2 * S);
    int* d_stack = cv.take<int>(n_stack);
    int* d_top = ctx->dev_int2;
    const size_t used = (size_t)(largest_index + 1);
    if (used) {
        pin_host_range(particles, used * sizeof(float), S * sizeof(float));
        pin_host_range(particles + S, used * sizeof(float), S * sizeof(float));
    }
    // the mover only APPENDS to the stack (ref infMagSim_c.c:214-217): nothing to upload, and only
    // the entries between the old and the new top come back
    const int top_in = *current_empty_slot;

    // Chunked pipeline: the used part of the store crosses PCIe in both directions every call
    // (16 B per slot), the mover itself needs ~3 ps per slot.  The slots are cut into chunks; chunk
    // i+1 is uploaded while chunk i is pushed and chunk i-1 is downloaded, on three streams (PCIe is
    // full duplex and the GPU has separate copy engines for the two directions).  Every chunk is a
    // complete mover call on its slot range (launch_move with slot_offset): removed slots reach the
    // stack chunk by chunk in ascending slot order, exactly the reference's order, and the integer
    // deposit accumulates over the chunks and is converted once, after the last one.
    size_t chunk = 8u << 20;  // slots per chunk: 32 MB per row and direction
    if (const char* e = getenv("ESPIC_HOST_CHUNK")) {  // tuning / test knob, read at every call
        const long long v = atoll(e);
        if (v >= 1024) chunk = (size_t)v & ~(size_t)1023;
    }
    // Chunk schedule: full chunks in the middle, halving chunks towards both ends -- the first upload and the last
    // download cannot overlap with anything, so they are kept short (1/8 of a chunk).
    std::vector<size_t> bounds;  // chunk c covers slots [bounds[c], bounds[c+1])
    bounds.push_back(0);
    if (used > chunk) {
        const size_t ramp[3] = {(chunk / 8) & ~(size_t)1023, (chunk / 4) & ~(size_t)1023, (chunk / 2) & ~(size_t)1023};
        const size_t ramp_total = ramp[0] + ramp[1] + ramp[2];
        size_t pos = 0;
        if (used > 2 * ramp_total + chunk && ramp[0] >= 1024) {
            for (int i = 0; i < 3; ++i) bounds.push_back(pos += ramp[i]);
            while (used - pos > ramp_total + chunk) bounds.push_back(pos += chunk);
            const size_t mid = (used - pos - ramp_total) & ~(size_t)1023;   // what is left in front of the down-ramp
            if (mid) bounds.push_back(pos += mid);
            bounds.push_back(pos += ramp[2]);
            bounds.push_back(pos += ramp[1]);
        } else {
            while (used - pos > chunk) bounds.push_back(pos += chunk);
        }
    }
    bounds.push_back(used);
    // Zero-copy mode (ESPIC_HOST_MODE=zerocopy; the default is the chunked pipeline): the mover reads and writes the
    // caller's page-locked rows directly over PCIe -- loads and stores of all resident warps keep both directions
    // of the link busy with no fill/drain bubble and no staging copy.  Needs both rows registered (they are, unless
    // registration failed or the store is small) and a device that can use host pointers of registered memory.
    bool zero_copy = false;
    {
        const char* mode = getenv("ESPIC_HOST_MODE");
        int can = 0;
        if (mode && !strcmp(mode, "zerocopy") && used >= 1024 && !minimal &&
            cudaDeviceGetAttribute(&can, cudaDevAttrCanUseHostPointerForRegisteredMem, ctx->device) == cudaSuccess && can)
            zero_copy = is_pinned(particles, used * sizeof(float)) && is_pinned(particles + S, used * sizeof(float));
    }
    if (zero_copy) {
        bounds.clear();
        bounds.push_back(0);
        bounds.push_back(used);
    }
    const size_t n_chunks = bounds.size() - 1;
    const bool pipelined = n_chunks > 1;
    cudaStream_t s_up = ctx->stream, s_run = ctx->stream, s_down = ctx->stream;
    if (pipelined) {
        for (int i = 0; i < 3; ++i)
            if (!ctx->pipe_stream[i] &&
                (rc = check_cuda(ctx, cudaStreamCreateWithFlags(&ctx->pipe_stream[i], cudaStreamNonBlocking), "stream")))
                die(ctx, who, rc);
        if ((int)(2 * n_chunks) > ctx->pipe_ev_cap) {
            const int cap = (int)(2 * n_chunks) + 16;
            cudaEvent_t* ev = (cudaEvent_t*)calloc((size_t)cap, sizeof(cudaEvent_t));
            if (!ev) die(ctx, who, ESPIC_E_NOMEM);
            for (int i = 0; i < ctx->pipe_ev_cap; ++i) ev[i] = ctx->pipe_ev[i];
            for (int i = ctx->pipe_ev_cap; i < cap; ++i)
                if ((rc = check_cuda(ctx, cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming), "event"))) die(ctx, who, rc);
            free(ctx->pipe_ev);
            ctx->pipe_ev = ev;
            ctx->pipe_ev_cap = cap;
        }
        // whatever the caller queued on the context's stream comes first
        if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, who, rc);
        s_up = ctx->pipe_stream[0];
        s_run = ctx->pipe_stream[1];
        s_down = ctx->pipe_stream[2];
    }
#define COPY(dst, src, count, T, kind, st) \
    if ((rc = check_cuda(ctx, cudaMemcpyAsync(dst, src, (size_t)(count) * sizeof(T), kind, st), "copy"))) die(ctx, who, rc)
    COPY(d_grid, grid, n, float, cudaMemcpyHostToDevice, s_run);
    if (minimal) {  // move_particles_c_minimal never looks at the mask (ref infMagSim_c.c:321-332, commented out)
        if ((rc = check_cuda(ctx, cudaMemsetAsync(d_mask, 0, n * sizeof(float), s_run), "memset"))) die(ctx, who, rc);
    } else {
        COPY(d_mask, object_mask, n, float, cudaMemcpyHostToDevice, s_run);
    }
    COPY(d_phi, potential, n, float, cudaMemcpyHostToDevice, s_run);
    COPY(d_top, current_empty_slot, 1, int, cudaMemcpyHostToDevice, s_run);
    cudaStream_t saved = ctx->stream;
    ctx->stream = s_run;
    for (size_t c = 0; c < n_chunks; ++c) {
        const size_t off = bounds[c];
        const size_t len = bounds[c + 1] - bounds[c];
        if (len && !zero_copy) {
            COPY(d_part + off, particles + off, len, float, cudaMemcpyHostToDevice, s_up);
            COPY(d_part + S + off, particles + S + off, len, float, cudaMemcpyHostToDevice, s_up);
        }
        if (pipelined) {
            cudaEventRecord(ctx->pipe_ev[2 * c], s_up);
            cudaStreamWaitEvent(s_run, ctx->pipe_ev[2 * c], 0);
        }
        rc = launch_move(ctx, d_grid, d_mask, d_phi, dt, charge_to_mass, background_density, (int)len - 1, nullptr,
                         zero_copy ? particles : d_part, d_den, d_stack, d_top, a_b, update_position, periodic_particles,
                         largest_allowed_slot, n_points, particle_storage_length, minimal, (int)off,
                         c + 1 == n_chunks ? 1 : 0);
        if (rc) {
            ctx->stream = saved;
            die(ctx, who, rc);
        }
        if (pipelined) {
            cudaEventRecord(ctx->pipe_ev[2 * c + 1], s_run);
            cudaStreamWaitEvent(s_down, ctx->pipe_ev[2 * c + 1], 0);
        }
        if (len && !zero_copy) {
            COPY(particles + off, d_part + off, len, float, cudaMemcpyDeviceToHost, s_down);
            COPY(particles + S + off, d_part + S + off, len, float, cudaMemcpyDeviceToHost, s_down);
        }
    }
    ctx->stream = saved;
    COPY(density, d_den, n, float, cudaMemcpyDeviceToHost, s_run);
    COPY(current_empty_slot, d_top, 1, int, cudaMemcpyDeviceToHost, s_run);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(s_run), "sync"))) die(ctx, who, rc);
    {
        const long long first = (long long)top_in + 1;
        long long last = *current_empty_slot;
        if (last > largest_allowed_slot) last = largest_allowed_slot;
        if (first >= 0 && last >= first)
            COPY(empty_slots + first, d_stack + first, last - first + 1, int, cudaMemcpyDeviceToHost, s_run);
    }
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(s_run), "sync"))) die(ctx, who, rc);
    if (pipelined && (rc = check_cuda(ctx, cudaStreamSynchronize(s_down), "sync"))) die(ctx, who, rc);
#undef COPY
    report_status(ctx);
}

extern "C" {

void move_particles_c(float* grid, float* object_mask, float* potential, float dt, float charge_to_mass,
                      float background_density, int largest_index, float* particles, float* density,
                      int* empty_slots, int* current_empty_slot, float a_b, int update_position,
                      int periodic_particles, int largest_allowed_slot, int n_points,
                      int particle_storage_length) {
    move_host_arrays("move_particles_c", grid, object_mask, potential, dt, charge_to_mass, background_density,
                     largest_index, particles, density, empty_slots, current_empty_slot, a_b, update_position,
                     periodic_particles, largest_allowed_slot, n_points, particle_storage_length, 0);
}

// ref infMagSim_c.c:231-341: the stripped mover -- never periodic (its periodic branches are commented
// out, so the flag is ignored), no frame acceleration, no object, and only slots that were inside the
// domain on entry can be removed (ref :333).
void move_particles_c_minimal(float* grid, float* object_mask, float* potential, float dt, float charge_to_mass,
                              float background_density, int largest_index, float* particles, float* density,
                              int* empty_slots, int* current_empty_slot, int update_position,
                              int periodic_particles, int largest_allowed_slot, int n_points,
                              int particle_storage_length) {
    (void)periodic_particles;
    move_host_arrays("move_particles_c_minimal", grid, object_mask, potential, dt, charge_to_mass,
                     background_density, largest_index, particles, density, empty_slots, current_empty_slot, 0.f,
                     update_position, 0, largest_allowed_slot, n_points, particle_storage_length, 1);
}

// ref infMagSim_c.c:343-354: Thomas algorithm on host double arrays; b and d are modified in place
// exactly as the reference modifies them (eliminated diagonal, back-substituted right-hand side).
void tridiagonal_solve_c(double* a, double* b, double* c, double* d, double* x, int n) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    const size_t m = (size_t)n;
    if ((rc = ensure(ctx, ctx->host_stage, 5 * padded(m * 8) + 256, "host staging"))) die(ctx, __func__, rc);
    Carver cv{(char*)ctx->host_stage.ptr};
    double* d_a = cv.take<double>(m);
    double* d_b = cv.take<double>(m);
    double* d_c = cv.take<double>(m);
    double* d_d = cv.take<double>(m);
    double* d_x = cv.take<double>(m);
    if (n > 1) {  // the reference reads a[0..n-2] and c[0..n-2]
        H2D(d_a, a, n - 1, double);
        H2D(d_c, c, n - 1, double);
    }
    H2D(d_b, b, n, double);
    H2D(d_d, d, n, double);
    if ((rc = launch_tridiagonal(ctx, d_a, d_b, d_c, d_d, d_x, n))) die(ctx, __func__, rc);
    D2H(b, d_b, n, double);
    D2H(d, d_d, n, double);
    D2H(x, d_x, n, double);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, __func__, rc);
}

// The host-array movers page-lock the caller's particle arrays on first use and keep them registered (the
// reference's driver passes the same arrays at every step).  A caller that frees or reallocates those arrays
// calls this first; so does anything that wants the pages back.
void espic_host_release(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    unpin_all();
}

void poisson_solve_c(float* grid, float* object_mask, float* charge, float debye_length, float* potential,
                     float object_potential, float object_transparency, int boltzmann_electrons,
                     int periodic_potential, int n_points) {
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    const size_t n = (size_t)n_points;
    if ((rc = ensure(ctx, ctx->host_stage, 4 * padded(n * 4) + 256, "host staging"))) die(ctx, __func__, rc);
    Carver cv{(char*)ctx->host_stage.ptr};
    float* d_grid = cv.take<float>(n);
    float* d_mask = cv.take<float>(n);
    float* d_rho = cv.take<float>(n);
    float* d_phi = cv.take<float>(n);
    H2D(d_grid, grid, n, float);
    H2D(d_mask, object_mask, n, float);
    H2D(d_rho, charge, n, float);
    H2D(d_phi, potential, n, float);
    rc = espic_poisson_solve(ctx, d_grid, d_mask, d_rho, debye_length, d_phi, object_potential,
                             object_transparency, boltzmann_electrons, periodic_potential, n_points);
    if (rc) die(ctx, __func__, rc);
    D2H(potential, d_phi, n, float);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, __func__, rc);
    report_status(ctx);
}

void sgenrand(unsigned long seed) {
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_seed(default_ctx(), seed);
}

void draw_velocities_c(int n, float v_th, float v_d, float* v_array) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    if ((rc = ensure(ctx, ctx->host_stage, padded((size_t)n * 4) + 256, "host staging"))) die(ctx, __func__, rc);
    float* d_v = (float*)ctx->host_stage.ptr;
    if ((rc = espic_draw_velocities(ctx, n, v_th, v_d, d_v))) die(ctx, __func__, rc);
    D2H(v_array, d_v, n, float);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, __func__, rc);
    report_status(ctx);
}

float genrand(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    float out = 0.f;
    float* d = (float*)(ctx->dev_int2 + 8);
    if ((rc = espic_genrand(ctx, 1, d))) die(ctx, __func__, rc);
    D2H(&out, d, 1, float);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, __func__, rc);
    return out;
}

void electric_field_filter_c(int n_points, float* grid, float* data_array, float L, int n_fine_mesh,
                             float* fine_mesh, float* Res) {
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    const size_t n = (size_t)n_points, m = (size_t)n_fine_mesh;
    if ((rc = ensure(ctx, ctx->host_stage, 2 * padded(n * 4) + 2 * padded(m * 4) + 256, "host staging")))
        die(ctx, __func__, rc);
    Carver cv{(char*)ctx->host_stage.ptr};
    float* d_grid = cv.take<float>(n);
    float* d_data = cv.take<float>(n);
    float* d_fine = cv.take<float>(m);
    float* d_res = cv.take<float>(m);
    H2D(d_grid, grid, n, float);
    H2D(d_data, data_array, n, float);
    H2D(d_fine, fine_mesh, m, float);
    if ((rc = espic_field_filter(ctx, n_points, d_grid, d_data, L, n_fine_mesh, d_fine, d_res))) die(ctx, __func__, rc);
    D2H(Res, d_res, m, float);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, __func__, rc);
}

void histogram2d_uniform_grid_c(float* X, float* Y, float x_min, float step_x, float y_min, float step_y,
                                int n_bins_x, int n_bins_y, int n_data, int* hist) {
    if (n_data <= 0) return;
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    const size_t n = (size_t)n_data, bins = (size_t)n_bins_x * n_bins_y;
    if ((rc = ensure(ctx, ctx->host_stage, 2 * padded(n * 4) + padded(bins * 4) + 256, "host staging")))
        die(ctx, __func__, rc);
    Carver cv{(char*)ctx->host_stage.ptr};
    float* d_x = cv.take<float>(n);
    float* d_y = cv.take<float>(n);
    int* d_h = cv.take<int>(bins);
    H2D(d_x, X, n, float);
    H2D(d_y, Y, n, float);
    H2D(d_h, hist, bins, int);
    rc = espic_phase_space_histogram(ctx, d_x, d_y, n_data, nullptr, 0, x_min, step_x, y_min, step_y, n_bins_x,
                                     n_bins_y, d_h);
    if (rc) die(ctx, __func__, rc);
    D2H(hist, d_h, bins, int);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, __func__, rc);
}

void inject_particles_c(int n_inject, float* grid, int n_points, float dt, float background_density,
                        float* velocities, float* uniform01, float a_b, int k, int* particles_hist,
                        float* particles, int particle_storage_length, int* empty_slots,
                        int* current_empty_slot, int* largest_index, float* density) {
    if (n_inject <= 0) return;
    std::lock_guard<std::mutex> lock(g_mutex);
    espic_ctx* ctx = default_ctx();
    DeviceGuard guard(ctx);
    int rc;
    const size_t S = (size_t)particle_storage_length, n = (size_t)n_points, m = (size_t)n_inject;
    const size_t need = 2 * padded(n * 4) + 2 * padded(m * 4) + padded(2 * S * 4) + 2 * padded(S * 4) + 256;
    if ((rc = ensure(ctx, ctx->host_stage, need, "host staging"))) die(ctx, __func__, rc);
    Carver cv{(char*)ctx->host_stage.ptr};
    float* d_grid = cv.take<float>(n);
    float* d_den = cv.take<float>(n);
    float* d_vel = cv.take<float>(m);
    float* d_uni = cv.take<float>(m);
    float* d_part = cv.take<float>(2 * S);
    int* d_stack = cv.take<int>(S);
    int* d_hist = cv.take<int>(S);
    int* d_top = ctx->dev_int2;
    int* d_last = ctx->dev_int2 + 1;
    H2D(d_grid, grid, n, float);
    H2D(d_den, density, n, float);
    H2D(d_vel, velocities, m, float);
    H2D(d_uni, uniform01, m, float);
    H2D(d_part, particles, 2 * S, float);
    H2D(d_stack, empty_slots, S, int);
    H2D(d_hist, particles_hist, S, int);
    H2D(d_top, current_empty_slot, 1, int);
    H2D(d_last, largest_index, 1, int);
    rc = espic_inject_particles(ctx, n_inject, d_grid, n_points, dt, background_density, d_vel, d_uni, a_b, k,
                                d_hist, d_part, particle_storage_length, d_stack, d_top, d_last, d_den);
    if (rc) die(ctx, __func__, rc);
    D2H(density, d_den, n, float);
    D2H(particles, d_part, 2 * S, float);
    D2H(empty_slots, d_stack, S, int);
    D2H(particles_hist, d_hist, S, int);
    D2H(current_empty_slot, d_top, 1, int);
    D2H(largest_index, d_last, 1, int);
    if ((rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sync"))) die(ctx, __func__, rc);
    report_status(ctx);
}

}  // extern "C"
