// K5: grid Poisson solve as one CTA.  Replaces poisson_solve_c + tridiagonal_solve_c
// (ref infMagSim_c.c:343-354, 356-541) and the density end-node fix / charge density of the
// driver loop (ref infMagSim_script.py:763-779, 807).
//
// The rows are built per grid node exactly as the reference builds them (float/double mix
// included).  The Thomas algorithm's three sequential recurrences are evaluated as block-wide
// scans in fp64:
//   D[j] = diag[j] - (sup[j-1]*sub[j-1])/D[j-1]        Moebius map  -> scan of 2x2 matrices
//   y[j] = rhs[j]  - (y[j-1]*sub[j-1])/D[j-1]          affine map   -> scan of (A,B) pairs
//   z[j] = y[j]    - (z[j+1]*sup[j])/D[j+1]            affine map, reversed
// Each of the 1024 threads owns a contiguous run of rows: it composes its run, the block scans
// the 1024 composites, and the thread then re-walks its run with the reference's own
// expression starting from the exact incoming value, so rounding differs from the sequential
// sweep only through that incoming value (~1e-16 relative; the potential is float).
//
// Periodic variant: bug-compatible with the reference: the second solve
// re-eliminates the already eliminated diagonal and v.w is taken as 0 (the reference reads
// one element past its arrays there; with glibc the product underflows to +0.0).
#include "espic_internal.cuh"

namespace espic {

constexpr int kSolveThreads = 1024;

struct Mat2 {
    double a, b, c, d;
};
struct Aff {
    double A, B;
};

__device__ __forceinline__ Mat2 mat_identity() { return Mat2{1., 0., 0., 1.}; }
__device__ __forceinline__ Aff aff_identity() { return Aff{1., 0.}; }

// later * earlier, rescaled (the map is projective) so long products neither overflow nor vanish
__device__ __forceinline__ Mat2 compose(const Mat2& l, const Mat2& e) {
    Mat2 r;
    r.a = l.a * e.a + l.b * e.c;
    r.b = l.a * e.b + l.b * e.d;
    r.c = l.c * e.a + l.d * e.c;
    r.d = l.c * e.b + l.d * e.d;
    double m = fmax(fmax(fabs(r.a), fabs(r.b)), fmax(fabs(r.c), fabs(r.d)));
    if (m > 0. && isfinite(m)) {
        int ex;
        frexp(m, &ex);
        r.a = ldexp(r.a, -ex);
        r.b = ldexp(r.b, -ex);
        r.c = ldexp(r.c, -ex);
        r.d = ldexp(r.d, -ex);
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
__device__ T block_exclusive_scan(T mine, T identity, T* s_warp) {
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

__device__ double block_sum(double v, double* s_red) {
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
    double r = s_red[0];
    __syncthreads();
    return r;
}

struct SolveArgs {
    const float* grid;
    const float* mask;
    const float* charge;
    float* potential;
    double* ws;  // 6*n doubles: low, diag, sup, rhs, D, D2
    int* status;
    float debye, obj_pot, transparency;
    int boltzmann, periodic, n;
};

// D_out[0] = d0; D_out[j] = dg[j] - (sup[j-1]*low[j])/D_out[j-1]
__device__ void eliminate_diagonal(const double* dg, const double* low, const double* sup, double* D_out,
                                   int n, Mat2* s_mat) {
    const int T = blockDim.x, t = threadIdx.x;
    const int m = (n - 1 + T - 1) / T;
    const int j0 = 1 + t * m, j1 = min(j0 + m, n);
    Mat2 comp = mat_identity();
    for (int j = j0; j < j1; ++j) comp = compose(Mat2{dg[j], -(sup[j - 1] * low[j]), 1., 0.}, comp);
    Mat2 pre = block_exclusive_scan(comp, mat_identity(), s_mat);
    const double d0 = dg[0];
    double d = (pre.a * d0 + pre.b) / (pre.c * d0 + pre.d);
    __syncthreads();  // every thread has read dg[0] before thread 0 may overwrite it (in-place use)
    if (t == 0) D_out[0] = d0;
    for (int j = j0; j < j1; ++j) {
        d = dg[j] - sup[j - 1] * low[j] / d;
        D_out[j] = d;
    }
    __syncthreads();
}

// in place: y[j] -= (y[j-1]*low[j])/D[j-1], j = 1..n-1
__device__ void forward_sweep(double* y, const double* low, const double* D, int n, Aff* s_aff) {
    const int T = blockDim.x, t = threadIdx.x;
    const int m = (n - 1 + T - 1) / T;
    const int j0 = 1 + t * m, j1 = min(j0 + m, n);
    Aff comp = aff_identity();
    for (int j = j0; j < j1; ++j) comp = compose(Aff{-(low[j] / D[j - 1]), y[j]}, comp);
    Aff pre = block_exclusive_scan(comp, aff_identity(), s_aff);
    double prev = pre.A * y[0] + pre.B;
    __syncthreads();
    for (int j = j0; j < j1; ++j) {
        prev = y[j] - prev * low[j] / D[j - 1];
        y[j] = prev;
    }
    __syncthreads();
}

// in place: z[j] -= (z[j+1]*sup[j])/D[j+1], j = n-2..0
__device__ void backward_sweep(double* z, const double* sup, const double* D, int n, Aff* s_aff) {
    const int T = blockDim.x, t = threadIdx.x;
    const int m = (n - 1 + T - 1) / T;
    const int i0 = 1 + t * m, i1 = min(i0 + m, n);  // step i handles row j = n-1-i
    Aff comp = aff_identity();
    for (int i = i0; i < i1; ++i) {
        const int j = n - 1 - i;
        comp = compose(Aff{-(sup[j] / D[j + 1]), z[j]}, comp);
    }
    Aff pre = block_exclusive_scan(comp, aff_identity(), s_aff);
    double next = pre.A * z[n - 1] + pre.B;
    __syncthreads();
    for (int i = i0; i < i1; ++i) {
        const int j = n - 1 - i;
        next = z[j] - next * sup[j] / D[j + 1];
        z[j] = next;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kSolveThreads) k_poisson(const SolveArgs a) {
    __shared__ Mat2 s_mat[32];
    __shared__ Aff s_aff[32];
    __shared__ double s_red[32];
    const int n = a.n, tid = threadIdx.x;
    double* low = a.ws;  // low[j] = sub-diagonal element of ROW j (the reference's lower_diagonal[j-1])
    double* dg = low + n;
    double* sup = dg + n;
    double* rhs = sup + n;
    double* D = rhs + n;
    double* D2 = D + n;

    // ---- rows (ref :385-503) ------------------------------------------------------------
    const float dz = (__ldg(a.grid + n - 1) - __ldg(a.grid)) / (float)(n - 1);
    const float lam = a.debye;
    const float k = 1.f;
    const float h2 = dz * dz / lam / lam;
    const float h1 = dz / lam;
    const double robin_diag = (double)(h2 * k) / 2. + (double)h1;
    const double robin_off = (double)(h2 * k) / 2. - (double)h1;
    const float rhs_scale = -dz * dz / lam / lam;
    const double near_one = 1. - (double)1.e-5f;
    const double tr = (double)a.transparency, op = 1. - (double)a.transparency;
    int bad_mask = 0;
    for (int j = tid; j < n; j += blockDim.x) {
        double p_low = 1., p_dg = -2., p_sup = 1.;
        double p_rhs = (j == n - 1) ? 0. : (double)a.charge[j];
        if (a.boltzmann) {
            const double ephi = exp((double)a.potential[j]);
            p_dg -= (double)(dz * dz) * ephi / (double)lam / (double)lam;
            p_rhs -= ephi * (1. - (double)a.potential[j]);
        }
        p_rhs *= (double)rhs_scale;
        if (j == 0) {
            p_dg = robin_diag;
            p_sup = robin_off;
            p_rhs = 0.;
        }
        if (j == n - 1) {
            if (a.periodic) {
                p_low = 0.;
                p_dg = 1.;
            } else {
                p_low = robin_off;
                p_dg = robin_diag;
            }
            p_rhs = 0.;
        }
        double o_low = p_low, o_dg = p_dg, o_sup = p_sup, o_rhs = p_rhs;
        if (j >= 1 && j <= n - 2) {
            const float m = a.mask[j], mp = a.mask[j - 1];
            if (m > 0.f) {
                if ((double)m < near_one) {
                    if (a.mask[j + 1] > 0.f) {
                        const float zl = (float)((1. - (double)m) * (double)dz);
                        o_low = (double)(zl / dz);
                        o_dg = -1. - (double)(zl / dz);
                        o_sup = 0.;
                        o_rhs *= (double)(zl / dz);
                        o_rhs -= (double)a.obj_pot;
                    } else if ((double)mp < near_one) {
                        const float zl = (float)((1. - (double)mp) * (double)dz);
                        const float zr = (float)((1. - (double)m) * (double)dz);
                        o_low = -(1. - (double)(zl / dz));
                        o_dg = (double)(-(zl + zr) / dz);
                        o_sup = -(1. - (double)(zr / dz));
                        o_rhs = -2. * (double)a.obj_pot;
                    } else if ((double)mp > near_one) {
                        const float zr = (float)((1. - (double)m) * (double)dz);
                        o_low = 0.;
                        o_dg = -2. * (double)zr / (double)dz;
                        o_sup = -2. * (1. - (double)(zr / dz));
                        o_rhs = -2. * (double)a.obj_pot;
                    } else {
                        bad_mask = 1;
                    }
                } else if ((double)m > near_one) {
                    if ((double)mp < near_one) {
                        const float zl = (float)((1. - (double)mp) * (double)dz);
                        o_low = -2. * (1. - (double)(zl / dz));
                        o_dg = -2. * (double)zl / (double)dz;
                        o_sup = 0.;
                        o_rhs = -2. * (double)a.obj_pot;
                    } else if ((double)mp > near_one) {
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
        low[j] = p_low * tr + o_low * op;
        dg[j] = p_dg * tr + o_dg * op;
        sup[j] = p_sup * tr + o_sup * op;
        rhs[j] = p_rhs * tr + o_rhs * op;
    }
    if (bad_mask) atomicOr(a.status, ESPIC_ST_BAD_OBJECT_MASK);
    __syncthreads();

    // ---- Thomas (ref :343-354) -------------------------------------------------------------
    eliminate_diagonal(dg, low, sup, D, n, s_mat);
    forward_sweep(rhs, low, D, n, s_aff);
    backward_sweep(rhs, sup, D, n, s_aff);
    for (int j = tid; j < n; j += blockDim.x) rhs[j] = rhs[j] / D[j];
    __syncthreads();

    double mean = 0.;
    if (a.periodic) {  // ref :504-514, :518-523
        eliminate_diagonal(D, low, sup, D2, n, s_mat);  // quirk (1): second elimination of the same diagonal
        // w = second "solve" of u = e_{n-1}: the forward sweep leaves u unchanged, so only the
        // backward sweep and the final division act.  low[] is free now: reuse it for w.
        double* w = low;
        for (int j = tid; j < n; j += blockDim.x) w[j] = (j == n - 1) ? 1. : 0.;
        __syncthreads();
        backward_sweep(w, sup, D2, n, s_aff);
        const double v_dot_result = -rhs[0];
        const double v_dot_w = 0.;  // quirk (2)
        __syncthreads();
        double part = 0.;
        for (int j = tid; j < n; j += blockDim.x) {
            const double r = rhs[j] - v_dot_result / (1. + v_dot_w) * (w[j] / D2[j]);
            rhs[j] = r;
            part += r;
        }
        mean = block_sum(part, s_red) / (double)n;
    }
    for (int j = tid; j < n; j += blockDim.x) a.potential[j] = (float)(rhs[j] - mean);
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

int launch_poisson(espic_ctx* ctx, const float* grid, const float* mask, const float* charge, float debye,
                   float* potential, float obj_pot, float transparency, int boltzmann, int periodic,
                   int n_points) {
    if (n_points < 3) return fail(ctx, ESPIC_E_ARG, "poisson_solve needs n_points >= 3");
    int rc = ensure(ctx, ctx->solve_ws, (size_t)6 * n_points * sizeof(double), "solve workspace");
    if (rc) return rc;
    SolveArgs a{};
    a.grid = grid;
    a.mask = mask;
    a.charge = charge;
    a.potential = potential;
    a.ws = (double*)ctx->solve_ws.ptr;
    a.status = ctx->status;
    a.debye = debye;
    a.obj_pot = obj_pot;
    a.transparency = transparency;
    a.boltzmann = boltzmann;
    a.periodic = periodic;
    a.n = n_points;
    k_poisson<<<1, kSolveThreads, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_poisson launch");
}

int launch_prepare_charge(espic_ctx* ctx, float* ion, float* ele, float* charge, int n, int periodic,
                          int fix_i, int fix_e) {
    if (n < 2) return fail(ctx, ESPIC_E_ARG, "n_points must be >= 2");
    k_prepare_charge<<<1, 1024, 0, ctx->stream>>>(ion, ele, charge, n, periodic, fix_i, fix_e);
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "k_prepare_charge launch");
}

}  // namespace espic
