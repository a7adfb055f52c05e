// Per-step diagnostics of the driver loop ("next" rows of SURVEY.md s.8f):
//   electric_field_filter_c (ref infMagSim_c.c:668-683) + np.gradient + argmax
//   (ref infMagSim_script.py:746-753)  -> hole position, one float per step;
//   histogram2d_uniform_grid_c (ref infMagSim_c.c:543-558) -> 100x100 phase-space counts.
#include "espic_internal.cuh"

namespace espic {

// tanh(s)/cosh(s) with one exponential: with e = exp(-|s|),
// tanh = (1-e^2)/(1+e^2), 1/cosh = 2e/(1+e^2).  Single precision with the hardware exponential
// and reciprocal: ~3e-7 relative, two orders below the rounding noise of the reference's own
// float accumulation over n_points terms (ref infMagSim_c.c:676-680), so the maximum it feeds
// lands on the same fine-mesh point.
__device__ __forceinline__ float tanh_over_cosh(float s) {
    const float e = __expf(-fabsf(s));
    const float e2 = e * e;
    const float den = 1.f + e2;
    const float r = __fdividef(2.f * e * (1.f - e2), den * den);
    return copysignf(r, s);
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
constexpr int kFilterThreads = 1024;

// Res[i] = sum_j tanh((fine[i]-grid[j])/L) / cosh((fine[i]-grid[j])/L) * signal[j].
// GRADIENT=false: data is the signal.  GRADIENT=true: data is a potential and the signal is
// numpy's gradient of it in float32 (central differences, one-sided ends).
// n_fine x n_points kernel evaluations (6.6e7 on the 65537-point grid, every step) bound by the
// two special-function operations each costs: a CTA takes kFilterFine mesh points at once so that
// grid and signal are loaded (and the gradient formed) once for all of them.  A thread sums its
// n_points/1024 products in float (64 terms on the largest grid), the partial sums are reduced in
// double.
template <bool GRADIENT>
__global__ void __launch_bounds__(kFilterThreads) k_field_filter(int n, const float* __restrict__ grid,
                                                                 const float* __restrict__ data, float div_mid,
                                                                 float div_end, float L, int n_fine,
                                                                 const float* __restrict__ fine, float* res,
                                                                 float* __restrict__ pos_out, unsigned* ticket) {
    __shared__ double s_red[kFilterFine][kFilterThreads / 32];
    __shared__ int s_last;
    const float inv_L = 1.f / L;
    const bool exact_inv = (inv_L * L == 1.f) && (__float_as_uint(L) & 0x007fffffu) == 0u;  // L a power of two
    for (int i0 = blockIdx.x * kFilterFine; i0 < n_fine; i0 += gridDim.x * kFilterFine) {
        float xf[kFilterFine], part[kFilterFine];
#pragma unroll
        for (int f = 0; f < kFilterFine; ++f) {
            xf[f] = fine[min(i0 + f, n_fine - 1)];
            part[f] = 0.f;
        }
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            float sig;
            if (GRADIENT) {
                if (j == 0) sig = __fdiv_rn(__fsub_rn(data[1], data[0]), div_end);
                else if (j == n - 1) sig = __fdiv_rn(__fsub_rn(data[n - 1], data[n - 2]), div_end);
                else sig = __fdiv_rn(__fsub_rn(data[j + 1], data[j - 1]), div_mid);
            } else {
                sig = data[j];
            }
            const float g = grid[j];
#pragma unroll
            for (int f = 0; f < kFilterFine; ++f) {
                const float d = __fsub_rn(xf[f], g);
                const float s = exact_inv ? d * inv_L : __fdiv_rn(d, L);   // ref :679, a float quotient
                part[f] = fmaf(tanh_over_cosh(s), sig, part[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < kFilterFine; ++f) {
            double p = (double)part[f];
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
    if (pos_out) {  // hole position = fine[argmax(Res)] (ref infMagSim_script.py:746-753), by the last CTA to finish
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            argmax_position(n_fine, res, fine, pos_out);
            if (threadIdx.x == 0) *ticket = 0u;
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

int launch_field_filter(espic_ctx* ctx, int n_points, const float* grid, const float* data, float L,
                        int n_fine, const float* fine, float* res) {
    if (n_points < 1 || n_fine < 1) return 0;
    const int chunks = (n_fine + kFilterFine - 1) / kFilterFine;
    k_field_filter<false><<<chunks < 2 * ctx->sm_count ? chunks : 2 * ctx->sm_count, kFilterThreads, 0, ctx->stream>>>(
        n_points, grid, data, 1.f, 1.f, L, n_fine, fine, res, nullptr, nullptr);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_field_filter launch");
}

int launch_hole_position(espic_ctx* ctx, int n_points, const float* grid, const float* potential, double dz,
                         float L, int n_fine, const float* fine, float* scratch_res, float* pos_out) {
    if (n_points < 2 || n_fine < 1) return fail(ctx, ESPIC_E_ARG, "hole_position needs n_points>=2, n_fine>=1");
    const int chunks = (n_fine + kFilterFine - 1) / kFilterFine;
    k_field_filter<true><<<chunks < 2 * ctx->sm_count ? chunks : 2 * ctx->sm_count, kFilterThreads, 0, ctx->stream>>>(
        n_points, grid, potential, (float)(2. * dz), (float)dz, L, n_fine, fine, scratch_res, pos_out,
        (unsigned*)(ctx->status + 9));
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "hole position launch");
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
