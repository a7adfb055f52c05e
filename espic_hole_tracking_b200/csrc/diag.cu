// Per-step diagnostics of the driver loop ("next" rows of SURVEY.md s.8f):
//   electric_field_filter_c (ref infMagSim_c.c:668-683) + np.gradient + argmax
//   (ref infMagSim_script.py:746-753)  -> hole position, one float per step;
//   histogram2d_uniform_grid_c (ref infMagSim_c.c:543-558) -> 100x100 phase-space counts.
#include "espic_internal.cuh"

namespace espic {

// tanh(s)/cosh(s) with one exponential: with e = exp(-|s|),
// tanh = (1-e^2)/(1+e^2), 1/cosh = 2e/(1+e^2).
__device__ __forceinline__ double tanh_over_cosh(double s) {
    const double e = exp(-fabs(s));
    const double e2 = e * e;
    const double den = 1. + e2;
    const double r = 2. * e * (1. - e2) / (den * den);
    return s < 0. ? -r : r;
}

// GRADIENT=false: data is used as given.  GRADIENT=true: data is a potential and the signal
// is numpy's gradient of it in float32 (central differences, one-sided ends).
template <bool GRADIENT>
__global__ void __launch_bounds__(256) k_field_filter(int n, const float* __restrict__ grid,
                                                      const float* __restrict__ data, float div_mid,
                                                      float div_end, float L, int n_fine,
                                                      const float* __restrict__ fine, float* __restrict__ res) {
    __shared__ double s_red[8];
    for (int i = blockIdx.x; i < n_fine; i += gridDim.x) {
        const float xf = fine[i];
        double part = 0.;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            float sig;
            if (GRADIENT) {
                if (j == 0) sig = __fdiv_rn(__fsub_rn(data[1], data[0]), div_end);
                else if (j == n - 1) sig = __fdiv_rn(__fsub_rn(data[n - 1], data[n - 2]), div_end);
                else sig = __fdiv_rn(__fsub_rn(data[j + 1], data[j - 1]), div_mid);
            } else {
                sig = data[j];
            }
            const double s = (double)__fdiv_rn(__fsub_rn(xf, grid[j]), L);
            part += tanh_over_cosh(s) * (double)sig;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
            res[i] = (float)t;
        }
        __syncthreads();
    }
}

// fine[argmax(res)] with numpy's first-maximum tie rule
__global__ void __launch_bounds__(1024) k_argmax_position(int n_fine, const float* __restrict__ res,
                                                          const float* __restrict__ fine, float* __restrict__ out) {
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n_fine; i += blockDim.x) {
        const float v = res[i];
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
    k_field_filter<false><<<n_fine < 4 * ctx->sm_count ? n_fine : 4 * ctx->sm_count, 256, 0, ctx->stream>>>(
        n_points, grid, data, 1.f, 1.f, L, n_fine, fine, res);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_field_filter launch");
}

int launch_hole_position(espic_ctx* ctx, int n_points, const float* grid, const float* potential, double dz,
                         float L, int n_fine, const float* fine, float* scratch_res, float* pos_out) {
    if (n_points < 2 || n_fine < 1) return fail(ctx, ESPIC_E_ARG, "hole_position needs n_points>=2, n_fine>=1");
    k_field_filter<true><<<n_fine < 4 * ctx->sm_count ? n_fine : 4 * ctx->sm_count, 256, 0, ctx->stream>>>(
        n_points, grid, potential, (float)(2. * dz), (float)dz, L, n_fine, fine, scratch_res);
    k_argmax_position<<<1, 1024, 0, ctx->stream>>>(n_fine, scratch_res, fine, pos_out);
    ctx->launches += 2;
    return check_cuda(ctx, cudaGetLastError(), "hole position launch");
}

int launch_histogram(espic_ctx* ctx, const float* xs, const float* ys, int n_data, const int* tags,
                     int tag_value, float x_min, float step_x, float y_min, float step_y, int nx, int ny,
                     int* hist) {
    if (n_data <= 0) return 0;
    if (nx < 1 || ny < 1) return fail(ctx, ESPIC_E_ARG, "histogram needs positive bin counts");
    const size_t smem = (size_t)nx * ny * sizeof(int);
    const int use_smem = smem <= 96 * 1024;
    if (use_smem && smem > 48 * 1024)
        ESPIC_CUDA(ctx, cudaFuncSetAttribute(k_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid_dim = (n_data + 256 * 16 - 1) / (256 * 16);
    if (grid_dim > 4 * ctx->sm_count) grid_dim = 4 * ctx->sm_count;
    if (grid_dim < 1) grid_dim = 1;
    k_histogram<<<grid_dim, 256, use_smem ? smem : 0, ctx->stream>>>(xs, ys, n_data, tags, tag_value, x_min, step_x,
                                                                     y_min, step_y, nx, ny, hist, use_smem);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_histogram launch");
}

}  // namespace espic
