// Per-step diagnostics of the driver loop ("next" rows of SURVEY.md s.8f):
//   electric_field_filter_c (ref infMagSim_c.c:668-683) + np.gradient + argmax
//   (ref infMagSim_script.py:746-753)  -> hole position, one float per step;
//   histogram2d_uniform_grid_c (ref infMagSim_c.c:543-558) -> 100x100 phase-space counts.
#include "espic_internal.cuh"

namespace espic {

// ---- the hole-tracking filter -----------------------------------------------------------------------
// Res[i] = sum_j K((fine[i]-grid[j])/L) * signal[j],  K(s) = tanh(s)/cosh(s)   (ref infMagSim_c.c:676-680)
// is n_fine x n_points kernel evaluations -- 6.6e7 on the 65537-point grid, every step, each with two
// special functions.  K is smooth on the scale of the grid spacing (dz/L = 1.8e-4 there), so the grid is
// cut into blocks of G consecutive points with |grid[j] - centre|/L <= 6e-3 and K is expanded around the
// block centre to third order:
//     K(u - d) = K(u) - d K1(u) + d^2/2 K2(u) - d^3/6 K3(u) + O(d^4 K4/24)   (< 1e-9 relative)
//     Res[i]  = sum_blocks  K(u_ib) M0_b + K1(u_ib) M1_b + K2(u_ib) M2_b + K3(u_ib) M3_b,
// u_ib = (fine[i] - centre_b)/L, M_k,b = sum_{j in b} (-(grid[j]-centre_b)/L)^k / k! * signal[j]: the
// moments cost one pass over the grid, the evaluations drop by the factor G (64 at 65537 points, 4 at
// 4097, 1 -- i.e. the plain sum -- at 513) and everything runs in double precision: the result is the
// exactly rounded response to ~1e-9, where the float accumulation of the reference (one rounding per
// term, ref :679) is off by ~1e-5 of the response at 65537 points.  With t = tanh u, c = sech u the
// derivatives are  K = t c,  K1 = c (1 - 2 t^2),  K2 = -t c (5 - 6 t^2),  K3 = c (-5 + 28 t^2 - 24 t^4).
constexpr int kFilterMaxGroup = 64;
constexpr double kFilterMaxDelta = 6e-3;

struct FilterMoments {   // planes of n_blocks doubles
    double* centre;
    double* m[4];
};

// One warp per block of G grid points: centre and the four moments (fixed summation tree: deterministic).
// GRADIENT: the signal is numpy's float32 gradient of `data` (central differences, one-sided ends,
// ref infMagSim_script.py:746-749), formed here instead of in a pass of its own.
template <bool GRADIENT>
__global__ void __launch_bounds__(256) k_filter_moments(int n, int G, int n_blocks, const float* __restrict__ grid,
                                                        const float* __restrict__ data, float div_mid, float div_end,
                                                        double inv_L, FilterMoments mo, int* status) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < n_blocks; b += n_warps) {
        const int j0 = b * G, j1 = min(n, j0 + G);
        const double c = 0.5 * ((double)grid[j0] + (double)grid[j1 - 1]);
        double m0 = 0., m1 = 0., m2 = 0., m3 = 0., dmax = 0.;
        for (int j = j0 + lane; j < j1; j += 32) {
            float sig;
            if (GRADIENT) {
                if (j == 0) sig = __fdiv_rn(__fsub_rn(data[1], data[0]), div_end);
                else if (j == n - 1) sig = __fdiv_rn(__fsub_rn(data[n - 1], data[n - 2]), div_end);
                else sig = __fdiv_rn(__fsub_rn(data[j + 1], data[j - 1]), div_mid);
            } else {
                sig = data[j];
            }
            const double d = -((double)grid[j] - c) * inv_L, sd = (double)sig;
            m0 += sd;
            m1 += sd * d;
            m2 += sd * (0.5 * d * d);
            m3 += sd * (d * d * d * (1. / 6.));
            dmax = fmax(dmax, fabs(d));
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            m0 += __shfl_xor_sync(0xffffffffu, m0, s);
            m1 += __shfl_xor_sync(0xffffffffu, m1, s);
            m2 += __shfl_xor_sync(0xffffffffu, m2, s);
            m3 += __shfl_xor_sync(0xffffffffu, m3, s);
            dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, s));
        }
        if (lane == 0) {
            mo.centre[b] = c;
            mo.m[0][b] = m0;
            mo.m[1][b] = m1;
            mo.m[2][b] = m2;
            mo.m[3][b] = m3;
            // a grid far from uniform can stretch a block beyond the range the expansion was sized for
            if (dmax > 4. * kFilterMaxDelta) atomicOr(status, ESPIC_ST_FILTER_GRID);
        }
    }
}

// fine[argmax(res)] with numpy's first-maximum tie rule, by one CTA (any width that is a multiple of 32)
__device__ __forceinline__ void argmax_position(int n_fine, const float* res, const float* __restrict__ fine,
                                                float* __restrict__ out) {
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n_fine; i += blockDim.x) {
        const float v = __ldcg(res + i);   // written by other CTAs of the same launch
        if (v > best || (v == best && i < idx) || idx == 0x7fffffff) {
            best = v;
            idx = i;
        }
    }
    auto better = [](float va, int ia, float vb, int ib) { return (vb > va) || (vb == va && ib < ia); };
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
        if (better(best, idx, ov, oi)) { best = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = (threadIdx.x < (blockDim.x >> 5)) ? s_val[threadIdx.x] : -INFINITY;
        idx = (threadIdx.x < (blockDim.x >> 5)) ? s_idx[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
            if (better(best, idx, ov, oi)) { best = ov; idx = oi; }
        }
        if (threadIdx.x == 0) out[0] = (idx < n_fine) ? fine[idx] : fine[0];
    }
}

constexpr int kFilterFine = 7;       // fine-mesh points per CTA: the 1001-point mesh makes 143 CTAs, one wave
constexpr int kFilterThreads = 256;

// Res for kFilterFine mesh points per CTA from the block moments; with pos_out the last CTA to finish
// takes fine[argmax(Res)] (ref infMagSim_script.py:750-753).
__global__ void __launch_bounds__(kFilterThreads) k_filter_eval(int n_blocks, FilterMoments mo, double inv_L, int n_fine,
                                                                const float* __restrict__ fine, float* res,
                                                                float* __restrict__ pos_out, unsigned* ticket,
                                                                int* cursor) {
    __shared__ double s_red[kFilterFine][kFilterThreads / 32];
    __shared__ int s_last;
    for (int i0 = blockIdx.x * kFilterFine; i0 < n_fine; i0 += gridDim.x * kFilterFine) {
        double xf[kFilterFine], part[kFilterFine];
#pragma unroll
        for (int f = 0; f < kFilterFine; ++f) {
            xf[f] = (double)fine[min(i0 + f, n_fine - 1)];
            part[f] = 0.;
        }
        for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
            const double c = mo.centre[b], m0 = mo.m[0][b], m1 = mo.m[1][b], m2 = mo.m[2][b], m3 = mo.m[3][b];
#pragma unroll
            for (int f = 0; f < kFilterFine; ++f) {
                const double u = (xf[f] - c) * inv_L;
                const double e = exp(-fabs(u)), e2 = e * e, r = 1. / (1. + e2);
                const double t = copysign((1. - e2) * r, u), ch = 2. * e * r, t2 = t * t;   // tanh u, sech u
                const double k0 = t * ch, k1 = ch * (1. - 2. * t2), k2 = -k0 * (5. - 6. * t2),
                             k3 = ch * (-5. + t2 * (28. - 24. * t2));
                part[f] += k0 * m0 + k1 * m1 + k2 * m2 + k3 * m3;
            }
        }
#pragma unroll
        for (int f = 0; f < kFilterFine; ++f) {
            double p = part[f];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) p += __shfl_xor_sync(0xffffffffu, p, d);
            if ((threadIdx.x & 31) == 0) s_red[f][threadIdx.x >> 5] = p;
        }
        __syncthreads();
        if (threadIdx.x < kFilterFine && i0 + (int)threadIdx.x < n_fine) {
            double t = 0.;
            for (int w = 0; w < kFilterThreads / 32; ++w) t += s_red[threadIdx.x][w];
            res[i0 + threadIdx.x] = (float)t;
        }
        __syncthreads();
    }
    if (pos_out) {  // hole position = fine[argmax(Res)], by the last CTA to finish
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            // replayed step graphs (espic_pic_run) cannot carry a per-step pointer: they index pos_out by a
            // device cursor, advanced here
            const int at = cursor ? *cursor : 0;
            argmax_position(n_fine, res, fine, pos_out + at);
            if (threadIdx.x == 0) {
                *ticket = 0u;
                if (cursor) *cursor = at + 1;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_histogram(const float* __restrict__ xs, const float* __restrict__ ys,
                                                   int n, const int* __restrict__ tags, int tag_value,
                                                   float x_min, float step_x, float y_min, float step_y, int nx,
                                                   int ny, int* __restrict__ hist, int use_smem) {
    extern __shared__ int s_hist[];
    const int bins = nx * ny;
    if (use_smem) {
        for (int b = threadIdx.x; b < bins; b += blockDim.x) s_hist[b] = 0;
        __syncthreads();
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (tags && tags[i] != tag_value) continue;
        // ref :551-553: float quotient, floor in double
        const int bx = (int)floor((double)__fdiv_rn(__fsub_rn(__ldcs(xs + i), x_min), step_x));
        if (bx < 0 || bx >= nx) continue;
        const int by = (int)floor((double)__fdiv_rn(__fsub_rn(__ldcs(ys + i), y_min), step_y));
        if (by < 0 || by >= ny) continue;
        if (use_smem) atomicAdd(s_hist + bx * ny + by, 1);
        else atomicAdd(hist + bx * ny + by, 1);
    }
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < bins; b += blockDim.x) {
            const int c = s_hist[b];
            if (c) atomicAdd(hist + b, c);
        }
    }
}

// Block size of the expansion for this grid and filter width: |grid[j] - centre|/L <= kFilterMaxDelta for a
// uniform grid (the reference builds its grid with np.linspace); 1 = plain summation.
static int filter_group(const float* grid_ends, int n_points, float L) {
    const double dz = fabs((double)grid_ends[1] - (double)grid_ends[0]) / (double)(n_points > 1 ? n_points - 1 : 1);
    if (!(dz > 0.) || !(L > 0.f)) return 1;
    const double g = 2. * kFilterMaxDelta * (double)L / dz;
    return g >= (double)kFilterMaxGroup ? kFilterMaxGroup : (g < 1. ? 1 : (int)g);
}

template <bool GRADIENT>
static int run_filter(espic_ctx* ctx, int n_points, const float* grid, const float* data, float div_mid, float div_end,
                      float L, int n_fine, const float* fine, float* res, float* pos_out, int* cursor = nullptr) {
    // the end points of the grid decide the block size; they never change in a run, so they are fetched once
    // per (grid pointer, n_points) and cached in the context
    if (ctx->filter_grid != grid || ctx->filter_n != n_points) {
        float ends[2];
        ESPIC_CUDA(ctx, cudaMemcpyAsync(&ends[0], grid, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        ESPIC_CUDA(ctx, cudaMemcpyAsync(&ends[1], grid + n_points - 1, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        ESPIC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->filter_grid = grid;
        ctx->filter_n = n_points;
        ctx->filter_ends[0] = ends[0];
        ctx->filter_ends[1] = ends[1];
    }
    const int G = filter_group(ctx->filter_ends, n_points, L);
    const int n_blocks = (n_points + G - 1) / G;
    int rc = ensure(ctx, ctx->filter_ws, (size_t)5 * n_blocks * sizeof(double) + 64, "filter moments");
    if (rc) return rc;
    FilterMoments mo;
    mo.centre = (double*)ctx->filter_ws.ptr;
    for (int k = 0; k < 4; ++k) mo.m[k] = mo.centre + (size_t)(k + 1) * n_blocks;
    const double inv_L = 1. / (double)L;
    int mg = (n_blocks + 7) / 8;   // 8 warps per CTA, one block per warp
    if (mg > 2 * ctx->sm_count) mg = 2 * ctx->sm_count;
    k_filter_moments<GRADIENT><<<mg, 256, 0, ctx->stream>>>(n_points, G, n_blocks, grid, data, div_mid, div_end, inv_L, mo,
                                                          ctx->status);
    const int chunks = (n_fine + kFilterFine - 1) / kFilterFine;
    k_filter_eval<<<chunks < 2 * ctx->sm_count ? chunks : 2 * ctx->sm_count, kFilterThreads, 0, ctx->stream>>>(
        n_blocks, mo, inv_L, n_fine, fine, res, pos_out, (unsigned*)(ctx->status + 9), cursor);
    ctx->launches += 2;
    return check_cuda(ctx, cudaGetLastError(), "field filter launch");
}

int launch_field_filter(espic_ctx* ctx, int n_points, const float* grid, const float* data, float L,
                        int n_fine, const float* fine, float* res) {
    if (n_points < 1 || n_fine < 1) return 0;
    return run_filter<false>(ctx, n_points, grid, data, 1.f, 1.f, L, n_fine, fine, res, nullptr);
}

int launch_hole_position(espic_ctx* ctx, int n_points, const float* grid, const float* potential, double dz,
                         float L, int n_fine, const float* fine, float* scratch_res, float* pos_out, int* cursor) {
    if (n_points < 2 || n_fine < 1) return fail(ctx, ESPIC_E_ARG, "hole_position needs n_points>=2, n_fine>=1");
    return run_filter<true>(ctx, n_points, grid, potential, (float)(2. * dz), (float)dz, L, n_fine, fine, scratch_res,
                            pos_out, cursor);
}

int launch_histogram(espic_ctx* ctx, const float* xs, const float* ys, int n_data, const int* tags,
                     int tag_value, float x_min, float step_x, float y_min, float step_y, int nx, int ny,
                     int* hist) {
    if (n_data <= 0) return 0;
    if (nx < 1 || ny < 1) return fail(ctx, ESPIC_E_ARG, "histogram needs positive bin counts");
    const size_t smem = (size_t)nx * ny * sizeof(int);
    const int use_smem = smem <= 96 * 1024;
    if (use_smem && smem > 48 * 1024 && smem > ctx->hist_smem_set) {
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->hist_smem_set = smem;
    }
    int grid_dim = (n_data + 256 * 16 - 1) / (256 * 16);
    if (grid_dim > 4 * ctx->sm_count) grid_dim = 4 * ctx->sm_count;
    if (grid_dim < 1) grid_dim = 1;
    k_histogram<<<grid_dim, 256, use_smem ? smem : 0, ctx->stream>>>(xs, ys, n_data, tags, tag_value, x_min, step_x,
                                                                     y_min, step_y, nx, ny, hist, use_smem);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_histogram launch");
}

}  // namespace espic
