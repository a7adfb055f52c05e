/*
 * espic_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the reference's per-timestep hot path (1-D, infinite-B
 * electrostatic PIC of ctzhou/ESPIC_hole_tracking).  It exists so that the CUDA product in
 * espic_hole_tracking_b200/csrc can be checked against an independent statement of the
 * same arithmetic.  Nothing in the product package may import, link or call this file:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker or the CPU arm.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so
 * this restatement is pinned against outputs of the reference itself: oracle/_ref/libref.so
 * is the unmodified infMagSim_c.c compiled with the reference's own flags (oracle/Makefile),
 * tests/test_oracle_vs_ref.py compares every function below bit-for-bit with it, and the
 * known-answer vectors recorded in SURVEY.md s.8c (MT19937 seeds 4357/384, the deposit
 * orientation cases) are committed under tests/golden/.
 *
 * Every function cites the reference lines it follows ("ref:" = /root/reference/...).
 * Arithmetic notes that matter for bit-exactness (the reference is built by gcc -O2 for
 * x86-64 without -march, i.e. FLT_EVAL_METHOD==0 and no FMA contraction; this file is
 * compiled with -ffp-contract=off for the same reason):
 *   - float expressions stay float unless a double literal (2., 0.99, 1.) appears in them;
 *   - the "still active" test  x < 0.99*flag  is evaluated in double;
 *   - 1.+fmodf(x,dz)/dz  adds in double and narrows to float on assignment.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* MT19937, 1997 "real" variant with the 69069 LCG seeding.  ref: infMagSim_c.c:36-103   */
/* ------------------------------------------------------------------------------------ */
enum { MT_LEN = 624, MT_GAP = 397 };
static uint32_t mt_state[MT_LEN];
static int mt_pos = MT_LEN + 1; /* MT_LEN+1 == "never seeded" (ref :51) */

void oracle_mt_seed(unsigned long seed) /* ref :54-65 */
{
    uint32_t s = (uint32_t)(seed & 0xffffffffUL);
    for (int i = 0; i < MT_LEN; ++i) {
        mt_state[i] = s;
        s = (uint32_t)(69069u * s);
    }
    mt_pos = MT_LEN;
}

static void mt_refill(void) /* ref :75-93 */
{
    if (mt_pos == MT_LEN + 1) oracle_mt_seed(4357UL);
    for (int k = 0; k < MT_LEN; ++k) {
        uint32_t hi = mt_state[k] & 0x80000000u;
        uint32_t lo = mt_state[(k + 1) % MT_LEN] & 0x7fffffffu;
        uint32_t y = hi | lo;
        uint32_t twist = (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mt_state[k] = mt_state[(k + MT_GAP) % MT_LEN] ^ twist;
    }
    mt_pos = 0;
}

uint32_t oracle_mt_next_u32(void) /* ref :95-99 (tempering) */
{
    if (mt_pos >= MT_LEN) mt_refill();
    uint32_t y = mt_state[mt_pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* ref :101  (float)y / (unsigned long)0xffffffff : both operands become float, so the
 * divisor is 4294967296.0f and the quotient can round up to exactly 1.0f. */
float oracle_mt_uniform(void)
{
    return (float)oracle_mt_next_u32() / 4294967296.0f;
}

/* ------------------------------------------------------------------------------------ */
/* Grid geometry shared by the mover, the injector and the solver.                        */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    float z_lo, z_hi, dz; /* ref infMagSim_c.c:126-128 */
    float parked;         /* inactive-slot marker 2*z_max, ref :129 */
    double live_limit;    /* 0.99*parked, compared in double, ref :146,172,211 */
    int n_nodes;
} geom_t;

static geom_t make_geom(const float *grid, int n_nodes)
{
    geom_t g;
    g.n_nodes = n_nodes;
    g.z_lo = grid[0];
    g.z_hi = grid[n_nodes - 1];
    g.dz = (g.z_hi - g.z_lo) / (float)(n_nodes - 1);
    g.parked = (float)(2. * (double)g.z_hi);
    g.live_limit = 0.99 * (double)g.parked;
    return g;
}

/* ref :148-153 / :174-179: fold a position back into [z_lo, z_hi) the reference's way */
static float fold_periodic(const geom_t *g, float x)
{
    float off = fmodf(x - g->z_lo, g->z_hi - g->z_lo);
    return (off >= 0.f) ? off + g->z_lo : off + g->z_hi;
}

/* ref :161-165 / :185-189 */
static int cell_of(const geom_t *g, float x, int periodic)
{
    int c = (int)((x - g->z_lo) / g->dz);
    if (periodic && c > g->n_nodes - 2) c = 0;
    if (c < 0 || c > g->n_nodes - 2)
        printf("bad left_index: %d, %f, %f, %f\n", c, g->z_lo, x, g->z_hi);
    return c;
}

/* ref :219-223.  NB the weight GROWS with distance from the left node (mirrored CIC) and is
 * taken from the absolute position, not from x - z_lo. */
static float left_node_share(const geom_t *g, float x)
{
    float frac = fmodf(x, g->dz) / g->dz;
    if (x > 0.f) return frac;
    return (float)(1. + (double)frac);
}

/* ------------------------------------------------------------------------------------ */
/* Fused gather / push / boundary / deposit.  ref: infMagSim_c.c:121-229                  */
/* Same argument order and meaning as move_particles_c.                                   */
/* ------------------------------------------------------------------------------------ */
void oracle_move_particles(const float *grid, const float *object_mask, const float *potential,
                           float dt, float charge_to_mass, float background_density,
                           int largest_index, float *particles, float *density,
                           int *empty_slots, int *current_empty_slot, float a_b,
                           int update_position, int periodic_particles,
                           int largest_allowed_slot, int n_points, int particle_storage_length)
{
    const geom_t g = make_geom(grid, n_points);
    const float edge = 1.e-6f; /* ref :130 */
    float *pos = particles;
    float *vel = particles + particle_storage_length;

    memset(density, 0, sizeof(float) * (size_t)n_points); /* ref :132-133 */

    int cell = 0; /* survives across slots exactly like the reference's function-scope variable */
    for (int slot = 0; slot <= largest_index; ++slot) {
        int alive;
        if (periodic_particles) { /* ref :145-154 */
            alive = (double)pos[slot] < g.live_limit;
            if (alive) pos[slot] = fold_periodic(&g, pos[slot]);
        } else { /* ref :156 */
            alive = (pos[slot] > g.z_lo + edge) && (pos[slot] < g.z_hi - edge);
        }
        const int alive_at_entry = alive;

        if (alive_at_entry) {
            cell = cell_of(&g, pos[slot], periodic_particles);
            /* piecewise-constant field of the cell, ref :166-168 */
            float e_cell = -(potential[cell + 1] - potential[cell]) / g.dz;
            float kick = charge_to_mass * e_cell - a_b;
            vel[slot] += kick * dt;
            if (update_position) { /* ref :169-198 */
                pos[slot] += vel[slot] * dt;
                if (periodic_particles) {
                    alive = (double)pos[slot] < g.live_limit;
                    if (alive) pos[slot] = fold_periodic(&g, pos[slot]);
                } else {
                    alive = (pos[slot] > g.z_lo + edge) && (pos[slot] < g.z_hi - edge);
                }
                if (alive) cell = cell_of(&g, pos[slot], periodic_particles);
            }
        }

        int swallowed = 0; /* absorbing object, ref :200-210 (inert while the mask is zero) */
        if (alive && object_mask[cell] > 0.f) {
            double covered = (object_mask[cell + 1] > 0.f) ? 1. - (double)object_mask[cell]
                                                           : (double)object_mask[cell];
            if ((double)pos[slot] > (double)grid[cell] + covered * (double)g.dz) swallowed = 1;
        }

        if (alive_at_entry || (double)pos[slot] < g.live_limit) { /* ref :211 */
            if (!alive || swallowed) { /* ref :212-217: park the slot and push it */
                pos[slot] = g.parked;
                *current_empty_slot += 1;
                if (*current_empty_slot > largest_allowed_slot)
                    printf("bad current_empty_slot: %d, %d\n", *current_empty_slot,
                           largest_allowed_slot);
                empty_slots[*current_empty_slot] = slot;
            } else { /* ref :219-225 */
                float share = left_node_share(&g, pos[slot]);
                density[cell] += share / background_density;
                density[cell + 1] += (1 - share) / background_density;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Thomas sweep exactly as the reference performs it (in place on diag and rhs).          */
/* ref: infMagSim_c.c:343-354.  sub[j] is the entry one row below diag[j].                */
/* ------------------------------------------------------------------------------------ */
void oracle_tridiagonal_solve(const double *sub, double *diag, const double *sup, double *rhs,
                              double *out, int n)
{
    for (int j = 1; j < n; ++j) {
        rhs[j] -= rhs[j - 1] * sub[j - 1] / diag[j - 1];
        diag[j] -= sup[j - 1] * sub[j - 1] / diag[j - 1];
    }
    for (int j = n - 1; j > 0; --j) rhs[j - 1] -= rhs[j] * sup[j - 1] / diag[j];
    for (int j = 0; j < n; ++j) out[j] = rhs[j] / diag[j];
}

typedef struct {
    double *sub, *diag, *sup, *rhs;
} rows_t;

static rows_t rows_alloc(int n)
{
    rows_t r;
    r.sub = (double *)calloc((size_t)n, sizeof(double));
    r.diag = (double *)calloc((size_t)n, sizeof(double));
    r.sup = (double *)calloc((size_t)n, sizeof(double));
    r.rhs = (double *)calloc((size_t)n, sizeof(double));
    return r;
}
static void rows_free(rows_t *r)
{
    free(r->sub);
    free(r->diag);
    free(r->sup);
    free(r->rhs);
}

/* ------------------------------------------------------------------------------------ */
/* Grid Poisson solve.  ref: infMagSim_c.c:356-541.  Same arguments as poisson_solve_c.    */
/*                                                                                        */
/* Two reference quirks are reproduced on purpose:                 */
/*  (1) in the periodic branch the second Thomas call re-eliminates the diagonal that the  */
/*      first call already modified (ref :506-507);                                        */
/*  (2) the statement accumulating v.w sits outside its loop and reads one element past    */
/*      both arrays (ref :510-512).  With glibc those two heap words are allocator         */
/*      bookkeeping whose product as doubles underflows to +0.0, so v.w == 0.0 here.       */
/* ------------------------------------------------------------------------------------ */
void oracle_poisson_solve(const float *grid, const float *object_mask, const float *charge,
                          float debye_length, float *potential, float object_potential,
                          float object_transparency, int boltzmann_electrons,
                          int periodic_potential, int n_points)
{
    const int n = n_points;
    const float tol = 1.e-5f;                               /* ref :360 */
    const float dz = (grid[n - 1] - grid[0]) / (float)(n - 1); /* ref :361-363 */
    const float k = 1.f;                                    /* ref :364 */
    const float h2 = dz * dz / debye_length / debye_length; /* float, as in ref :403,407 */
    const float h1 = dz / debye_length;
    /* ref :407-408 / :426-427: float product, then "/2." and "+" in double */
    const double robin_diag = (double)(h2 * k) / 2. + (double)h1;
    const double robin_off = (double)(h2 * k) / 2. - (double)h1;

    rows_t plain = rows_alloc(n), obj = rows_alloc(n);
    double *u = (double *)calloc((size_t)n, sizeof(double));
    double *w = (double *)calloc((size_t)n, sizeof(double));
    double *phi = (double *)calloc((size_t)n, sizeof(double));

    for (int j = 0; j < n; ++j) { /* ref :385-395 */
        plain.diag[j] = -2.;
        plain.sub[j] = 1.;
        plain.sup[j] = 1.;
        plain.rhs[j] = (j == n - 1) ? 0. : (double)charge[j];
    }
    if (boltzmann_electrons) { /* ref :396-401 */
        for (int j = 0; j < n; ++j) {
            double ephi = exp((double)potential[j]);
            plain.diag[j] -= (double)(dz * dz) * ephi / (double)debye_length / (double)debye_length;
            plain.rhs[j] -= ephi * (1. - (double)potential[j]);
        }
    }
    {
        const float scale = -dz * dz / debye_length / debye_length; /* ref :403 */
        for (int j = 0; j < n; ++j) plain.rhs[j] *= (double)scale;
    }
    plain.diag[0] = robin_diag; /* ref :407-409 */
    plain.sup[0] = robin_off;
    plain.rhs[0] = 0.;
    if (periodic_potential) { /* ref :414-421 */
        plain.sub[n - 2] = 0.;
        plain.diag[n - 1] = 1.;
        plain.rhs[n - 1] = 0.;
        u[n - 1] = 1.;
    } else { /* ref :426-428 */
        plain.sub[n - 2] = robin_off;
        plain.diag[n - 1] = robin_diag;
        plain.rhs[n - 1] = 0.;
    }
    memcpy(obj.sub, plain.sub, sizeof(double) * (size_t)n); /* ref :430-435 */
    memcpy(obj.diag, plain.diag, sizeof(double) * (size_t)n);
    memcpy(obj.sup, plain.sup, sizeof(double) * (size_t)n);
    memcpy(obj.rhs, plain.rhs, sizeof(double) * (size_t)n);

    /* rows touched by the absorbing object, ref :439-497 (zl, zr are floats there) */
    for (int j = 1; j < n - 1; ++j) {
        const float m = object_mask[j], m_prev = object_mask[j - 1];
        const double near_one = 1. - (double)tol;
        if (m > 0.f) {
            if ((double)m < near_one) {
                if (object_mask[j + 1] > 0.f) { /* ref :444-451 */
                    float zl = (float)((1. - (double)m) * (double)dz);
                    obj.sub[j - 1] = (double)(zl / dz);
                    obj.diag[j] = -1. - (double)(zl / dz);
                    obj.sup[j] = 0.;
                    obj.rhs[j] *= (double)(zl / dz);
                    obj.rhs[j] -= (double)object_potential;
                } else if ((double)m_prev < near_one) { /* ref :452-459 */
                    float zl = (float)((1. - (double)m_prev) * (double)dz);
                    float zr = (float)((1. - (double)m) * (double)dz);
                    obj.sub[j - 1] = -(1. - (double)(zl / dz));
                    obj.diag[j] = (double)(-(zl + zr) / dz);
                    obj.sup[j] = -(1. - (double)(zr / dz));
                    obj.rhs[j] = -2. * (double)object_potential;
                } else if ((double)m_prev > near_one) { /* ref :460-466 */
                    float zr = (float)((1. - (double)m) * (double)dz);
                    obj.sub[j - 1] = 0.;
                    obj.diag[j] = -2. * (double)zr / (double)dz;
                    obj.sup[j] = -2. * (1. - (double)(zr / dz));
                    obj.rhs[j] = -2. * (double)object_potential;
                } else {
                    printf("bad object_mask1\n");
                }
            } else if ((double)m > near_one) {
                if ((double)m_prev < near_one) { /* ref :471-477 */
                    float zl = (float)((1. - (double)m_prev) * (double)dz);
                    obj.sub[j - 1] = -2. * (1. - (double)(zl / dz));
                    obj.diag[j] = -2. * (double)zl / (double)dz;
                    obj.sup[j] = 0.;
                    obj.rhs[j] = -2. * (double)object_potential;
                } else if ((double)m_prev > near_one) { /* ref :478-480 */
                    obj.rhs[j] = 0.;
                } else {
                    printf("bad object_mask2\n");
                }
            } else {
                printf("bad object_mask3\n");
            }
        } else if (m_prev > 0.f) { /* ref :487-496 */
            float zr = (float)((1. - (double)m_prev) * (double)dz);
            obj.sub[j - 1] = 0.;
            obj.diag[j] = -1. - (double)(zr / dz);
            obj.sup[j] = (double)(zr / dz);
            obj.rhs[j] *= (double)(zr / dz);
            obj.rhs[j] -= (double)object_potential;
        }
    }
    { /* blend, ref :498-503 */
        const double t = (double)object_transparency, s = 1. - (double)object_transparency;
        for (int j = 0; j < n; ++j) {
            plain.sup[j] = plain.sup[j] * t + obj.sup[j] * s;
            plain.diag[j] = plain.diag[j] * t + obj.diag[j] * s;
            plain.sub[j] = plain.sub[j] * t + obj.sub[j] * s;
            plain.rhs[j] = plain.rhs[j] * t + obj.rhs[j] * s;
        }
    }

    oracle_tridiagonal_solve(plain.sub, plain.diag, plain.sup, plain.rhs, phi, n);
    double mean = 0.;
    if (periodic_potential) { /* ref :504-514, :518-523 */
        oracle_tridiagonal_solve(plain.sub, plain.diag, plain.sup, u, w, n); /* quirk (1) */
        double v_dot_phi = 0.;
        for (int j = 0; j < n; ++j) v_dot_phi += ((j == 0) ? -1. : 0.) * phi[j];
        const double v_dot_w = 0.; /* quirk (2) */
        for (int j = 0; j < n; ++j) phi[j] = phi[j] - v_dot_phi / (1. + v_dot_w) * w[j];
        for (int j = 0; j < n; ++j) mean += phi[j];
        mean /= n;
    }
    for (int j = 0; j < n; ++j) potential[j] = (float)(phi[j] - mean); /* ref :524-525 */

    rows_free(&plain);
    rows_free(&obj);
    free(u);
    free(w);
    free(phi);
}

/* ------------------------------------------------------------------------------------ */
/* Flux-weighted drifting-Maxwellian rejection sampler for wall injection.                */
/* ref: infMagSim_c.c:560-666.  Consumes the MT19937 stream above in the same order.      */
/* ------------------------------------------------------------------------------------ */
void oracle_draw_velocities(int n, float v_th, float v_d, float *v_out)
{
    const float half_window = 5.f; /* ref :576 */
    const float beta = (float)((double)v_d / (sqrt(2.) * (double)v_th)); /* ref :577 */
    const double b = (double)beta;
    const double gauss = 1. / sqrt(M_PI) * exp(-pow(b, 2.));
    /* ref :578: share of the wall flux entering through the right-hand wall over the left */
    const float ratio = (float)((gauss - b * (1. - erf(b))) / (gauss + b * (1. + erf(b))));
    const double vd = (double)v_d, vt = (double)v_th;
    const double disc = sqrt(pow(vd, 2.) + 4. * pow(vt, 2.));

    for (int i = 0; i < n; ++i) {
        /* ref :584: 1/(1+ratio) is int/float -> float; the subtraction is float */
        float side = oracle_mt_uniform() - 1.f / (1.f + ratio);
        float mode, lo, hi; /* floats in the reference, assigned from double expressions */
        if (side >= 0.f) { /* enters at z_min moving right, ref :608-610 */
            mode = (float)((-vd + disc) / 2.);
            hi = mode + half_window * v_th;
            lo = (float)fmax((double)(mode - half_window * v_th), 0.);
        } else { /* enters at z_max moving left, ref :648-650 */
            mode = (float)((-vd - disc) / 2.);
            hi = (float)fmin((double)(mode + half_window * v_th), 0.);
            lo = mode - half_window * v_th;
        }
        int tries = 1;
        for (;;) {
            if (tries == 1000) /* ref :589-607: the reference only prints and keeps looping */
                printf("Warning: method draw_velocities_c doesn't converge\n");
            float v = oracle_mt_uniform() * (hi - lo) + lo; /* ref :611-612 / :651-652 */
            float accept = oracle_mt_uniform();
            /* ref :614 / :654, evaluated in double -- except v_max+v_d and v+v_d, which are
             * float additions before pow() promotes them */
            double bound = log(fabs((double)v) / fabs((double)mode)) +
                           (pow((double)(mode + v_d), 2.) - pow((double)(v + v_d), 2.)) /
                               (2. * pow(vt, 2.));
            if (log((double)accept) <= bound) {
                v_out[i] = v;
                break;
            }
            ++tries;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Wall injection: C restatement of the Cython body, with the two random inputs           */
/* (velocities from the sampler above, uniforms from the host sampler) passed in.         */
/* ref: infMagSim_cython.py:230-285.                                                       */
/* ------------------------------------------------------------------------------------ */
void oracle_inject_particles(int n_inject, const float *grid, int n_points, float dt,
                             float background_density, const float *velocities,
                             const float *uniform01, float a_b, int k, int *particles_hist,
                             float *particles, int particle_storage_length, int *empty_slots,
                             int *current_empty_slot, int *largest_index, float *density)
{
    const geom_t g = make_geom(grid, n_points);
    const float edge = 2.e-6f; /* ref :238 */
    float *pos = particles;
    float *vel = particles + particle_storage_length;
    int top = *current_empty_slot, last = *largest_index;

    for (int l = 0; l < n_inject; ++l) {
        if (top < 0) printf("no empty slots\n"); /* ref :255-256 */
        const int slot = empty_slots[top];
        const float part_dt = dt * uniform01[l]; /* ref :246: float32 array times scalar */
        particles_hist[slot] = k;                /* ref :261 */
        vel[slot] = velocities[l] - a_b * part_dt; /* ref :262 */
        if (velocities[l] < 0.f) /* ref :263-268 */
            pos[slot] = g.z_hi - edge + part_dt * vel[slot];
        else
            pos[slot] = g.z_lo + edge + part_dt * vel[slot];
        if (slot > last) last = slot; /* ref :269-270 */
        int cell = (int)((pos[slot] - g.z_lo) / g.dz); /* ref :271 */
        if (cell < 0 || cell > n_points - 2) {
            /* ref :272-274: only a message; the slot is NOT consumed and the next injected
             * particle overwrites it */
            printf("bad left_index: %d %f %f %f %f %f\n", cell, g.z_lo, pos[slot], g.z_hi,
                   vel[slot], part_dt);
        } else { /* ref :275-283: math.fmod and the division run in double, then narrow */
            double frac = fmod((double)pos[slot], (double)g.dz) / (double)g.dz;
            float share = (pos[slot] > 0.f) ? (float)frac : (float)(1. + frac);
            density[cell] += share / background_density;
            density[cell + 1] += (1 - share) / background_density;
            empty_slots[top] = -1;
            top -= 1;
        }
    }
    *largest_index = last;
    *current_empty_slot = top;
}

/* ------------------------------------------------------------------------------------ */
/* "Next" rows of SURVEY.md s.8f, restated for the diagnostics kernels.                   */
/* ------------------------------------------------------------------------------------ */

/* ref: infMagSim_c.c:668-683 -- tanh/cosh matched filter used by the hole tracker.
 * tanh and cosh take double arguments; the running sum is a float. */
void oracle_field_filter(int n_points, const float *grid, const float *signal, float width,
                         int n_fine, const float *fine_mesh, float *response)
{
    for (int i = 0; i < n_fine; ++i) {
        float acc = 0.f;
        for (int j = 0; j < n_points; ++j) {
            double s = (double)((fine_mesh[i] - grid[j]) / width);
            acc = (float)((double)acc + tanh(s) / cosh(s) * (double)signal[j]);
        }
        response[i] = acc;
    }
}

/* ref: infMagSim_c.c:543-558 -- phase-space histogram on a uniform mesh */
void oracle_histogram2d(const float *xs, const float *ys, float x0, float step_x, float y0,
                        float step_y, int nx, int ny, int n_data, int *hist)
{
    for (int i = 0; i < n_data; ++i) {
        int bx = (int)floor((double)((xs[i] - x0) / step_x));
        int by = (int)floor((double)((ys[i] - y0) / step_y));
        if (bx >= 0 && bx < nx && by >= 0 && by < ny) hist[bx * ny + by] += 1;
    }
}
