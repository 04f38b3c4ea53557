/* oracle/nhm_step.c — a second, independent CPU restatement (plain C99 + OpenMP) of the reference algorithm for the
 * triply periodic benchmark physics (BASELINE configs C2 / C3): NonhydrostaticModel on a regular (Periodic, Periodic, Periodic)
 * RectilinearGrid, Centered(order=2) or WENO(order=5) flux-form advection, optional T,S tracers with SeawaterBuoyancy
 * (LinearEquationOfState) + hydrostatic pressure split, ScalarDiffusivity, RungeKutta3, FFTBasedPoissonSolver.
 *
 * TEST INFRASTRUCTURE ONLY (like everything under oracle/): it cross-checks the NumPy oracle (tests/test_oracle_c.py) and is the
 * multi-core CPU baseline of bench.py.  It is never linked into or loaded by the product library.
 * Parity status: pinned against the NumPy oracle, which is pinned by the reference's known-answer tests; parity against the Julia
 * binary itself is unpinned (see oracle/__init__.py).
 *
 * Like the reference (one KernelAbstractions thread per cell, every thread evaluates both faces of its cell in each direction:
 * src/Advection/momentum_advection_operators.jl:46-83, src/Operators/difference_operators.jl:20-27) each face flux is
 * computed twice.  Paths below are relative to /root/reference.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define H 3

/* Arithmetic type.  The default build (double) is the restatement proper.  -DNHC_LONG_DOUBLE evaluates the SAME formulas, with the SAME
 * Float64-rounded coefficients and inputs, in x87 extended precision (64-bit mantissa): tests use it to separate the round-off of the
 * reference's own Float64 formulas (its expanded WENO smoothness indicators cancel catastrophically for T ≈ 20, S ≈ 35) from kernel
 * errors — see tests/test_gpu_parity.py::test_cuda_matches_c_twin_at_128_cubed. */
#ifdef NHC_LONG_DOUBLE
typedef long double real;
#define FMA fmal
#define COS cosl
#define SIN sinl
#define POW powl
#define PI_ 3.141592653589793238462643383279502884L
#else
typedef double real;
#define FMA fma
#define COS cos
#define SIN sin
#define POW pow
#define PI_ M_PI
#endif
#define FABS(x) ((x) < 0 ? -(x) : (x))


typedef struct {
    int Nx, Ny, Nz, weno, ntr;
    int beta_difference_form;                 /* 0 = the reference's expanded smoothness indicators; 1 = see weno5() */
    long sy, sz, n;
    real d[3], rd[3], A[3], V, rV, L[3];
    real nu, kappa, grav, alpha, beta;
    real *f[5], *Gn[5], *Gm[5], *pHY, *p;   /* u v w T S */
    real *lam[3];
    real *re, *im;                          /* spectral work arrays, Nx*Ny*Nz */
    /* reconstruction coefficients (src/Advection/reconstruction_coefficients.jl:49-64) */
    real c4[4], w5p[3][3];
    real gam[3], zet[3];
} M;

static long IDX(const M* m, int i, int j, int k) { return (i + H) + (long)(j + H) * m->sy + (long)(k + H) * m->sz; }
static long ST(const M* m, int d) { return d == 0 ? 1 : (d == 1 ? m->sy : m->sz); }

/* ---- fill_halo_regions! for Periodic sides: src/BoundaryConditions/fill_halo_regions_periodic.jl:5-32 ---------------- */
static void fill_halo(const M* m, real* f) {
    const int Nx = m->Nx, Ny = m->Ny, Nz = m->Nz;
#pragma omp parallel for collapse(2)
    for (int k = -H; k < Nz + H; ++k)
        for (int j = -H; j < Ny + H; ++j) {
            int kk = ((k % Nz) + Nz) % Nz, jj = ((j % Ny) + Ny) % Ny;
            for (int i = -H; i < Nx + H; ++i) {
                if (i >= 0 && i < Nx && j >= 0 && j < Ny && k >= 0 && k < Nz) continue;
                int ii = ((i % Nx) + Nx) % Nx;
                f[IDX(m, i, j, k)] = f[IDX(m, ii, jj, kk)];
            }
        }
}

/* ---- reconstructions ------------------------------------------------------------------------------------------------- */
/* newton_div(Float32, a, b)  src/Utils/newton_div.jl:8-20 */
static real newton_div(real a, real b) {
    real inv = (real)(1.0f / (float)b);
    real x = a * inv;
    return FMA(FMA(x, -b, a), inv, x);
}
static real beta3(real a, real b, real c, real C1, real C2, real C3, real C4, real C5, real C6) {
    return a * (C1 * a + C2 * b + C3 * c) + b * (C4 * b + C5 * c) + c * c * C6;   /* weno_interpolants.jl:204-216,261 */
}
/* WENO{3} biased reconstruction at the face whose right cell is p[0] (ψ[i]); stride s  (weno_interpolants.jl:290-337,409-437,500) */
static real weno5(const M* m, const real* p, long s, int left) {
    real q0, q1, q2, q3, q4;
    if (left) { q0 = p[-3 * s]; q1 = p[-2 * s]; q2 = p[-s]; q3 = p[0]; q4 = p[s]; }
    else      { q0 = p[2 * s];  q1 = p[s];      q2 = p[0];  q3 = p[-s]; q4 = p[-2 * s]; }
    real b0, b1, b2;
    if (!m->beta_difference_form) {
        b0 = beta3(q2, q3, q4, 10, -31, 11, 25, -19, 4);
        b1 = beta3(q1, q2, q3, 4, -13, 5, 13, -13, 4);
        b2 = beta3(q0, q1, q2, 4, -19, 11, 25, -31, 10);
    } else {
        /* the same three quadratic forms written as 13/4 (second difference)^2 + 3/4 (one-sided first difference)^2 — an exact algebraic
         * identity (expand: 10 a^2 - 31 ab + 11 ac + 25 b^2 - 19 bc + 4 c^2 etc.) that does not cancel for fields with a large mean.
         * NOT the reference's order of operations: used by tests to attribute T / S differences to the reference's own round-off. */
        const real e1 = q1 - q0, e2 = q2 - q1, e3 = q3 - q2, e4 = q4 - q3;
        const real d0 = e4 - e3, d1 = e3 - e2, d2 = e2 - e1, g0 = e4 - 3 * e3, g1 = e2 + e3, g2 = 3 * e2 - e1;
        b0 = (real)3.25 * (d0 * d0) + (real)0.75 * (g0 * g0);
        b1 = (real)3.25 * (d1 * d1) + (real)0.75 * (g1 * g1);
        b2 = (real)3.25 * (d2 * d2) + (real)0.75 * (g2 * g2);
    }
    const real eps = (real)1e-8f;
    real tau = FABS(b0 - b2);
    real r0 = newton_div(tau, b0 + eps), r1 = newton_div(tau, b1 + eps), r2 = newton_div(tau, b2 + eps);
    real a0 = 0.3 * (1 + r0 * r0), a1 = 0.6 * (1 + r1 * r1), a2 = 0.1 * (1 + r2 * r2);
    real rs = 1 / (a0 + a1 + a2);
    real p0 = m->w5p[0][0] * q2 + m->w5p[0][1] * q3 + m->w5p[0][2] * q4;
    real p1 = m->w5p[1][0] * q1 + m->w5p[1][1] * q2 + m->w5p[1][2] * q3;
    real p2 = m->w5p[2][0] * q0 + m->w5p[2][1] * q1 + m->w5p[2][2] * q2;
    return (a0 * rs) * p0 + (a1 * rs) * p1 + (a2 * rs) * p2;
}
/* Centered(order=4) of a·q at the face whose right cell is p[0]  (centered_reconstruction.jl:47-55) */
static real sym4(const M* m, const real* p, long s, real a) {
    real r = m->c4[0] * (a * p[-2 * s]);
    r = r + m->c4[1] * (a * p[-s]);
    r = r + m->c4[2] * (a * p[0]);
    r = r + m->c4[3] * (a * p[s]);
    return r;
}

/* advective flux of momentum component c through the face normal to d at flux index o (centre-type when d == c)
 * src/Advection/upwind_biased_advective_fluxes.jl:23-93, centered_advective_fluxes.jl:15-27 */
static real mom_flux(const M* m, int c, int d, long o) {
    const real* psi = m->f[c] + o;
    const real* adv = m->f[d] + o;
    const long sd = ST(m, d), sc = ST(m, c);
    const real A = m->A[d];
    if (d == c) {
        if (!m->weno) { real ut = 0.5 * adv[0] + 0.5 * adv[sd], pt = 0.5 * psi[0] + 0.5 * psi[sd]; return A * ut * pt; }
        real ut = sym4(m, adv + sd, sd, A);
        return ut * weno5(m, psi + sd, sd, ut > 0);
    }
    if (!m->weno) { real ut = 0.5 * adv[-sc] + 0.5 * adv[0], pt = 0.5 * psi[-sd] + 0.5 * psi[0]; return A * ut * pt; }
    real ut = sym4(m, adv, sc, A);
    return ut * weno5(m, psi, sd, ut > 0);
}
/* viscous flux A_d τ_cd, τ = -2 ν Σ  (src/TurbulenceClosures/abstract_scalar_diffusivity_closure.jl:189-204,
 * velocity_tracer_gradients.jl:25-42, closure_kernel_operators.jl:22-41) */
static real visc_flux(const M* m, int c, int d, long o) {
    real sig;
    if (d == c) sig = (m->f[d][o + ST(m, d)] - m->f[d][o]) * m->rd[d];
    else {
        int lo = d < c ? d : c, hi = d < c ? c : d;
        real dl = (m->f[lo][o] - m->f[lo][o - ST(m, hi)]) * m->rd[hi];
        real dh = (m->f[hi][o] - m->f[hi][o - ST(m, lo)]) * m->rd[lo];
        sig = 0.5 * (dl + dh);
    }
    return m->A[d] * (-2 * (m->nu * sig));
}
/* tracer fluxes  (upwind_biased_advective_fluxes.jl:99-121, centered_advective_fluxes.jl:31-33; :240-242) */
static real tr_flux(const M* m, int t, int d, long o) {
    const real* c = m->f[t] + o;
    const long sd = ST(m, d);
    const real u = m->f[d][o], A = m->A[d];
    real F;
    if (!m->weno) F = (A * u) * (0.5 * c[-sd] + 0.5 * c[0]);
    else F = A * u * weno5(m, c, sd, u > 0);
    real grad = (c[0] - c[-sd]) * m->rd[d];
    return F + m->A[d] * (-(m->kappa * grad));
}

/* ---- update_hydrostatic_pressure!  update_hydrostatic_pressure.jl:12-49 ------------------------------------------------ */
static real buoy(const M* m, long o) { return m->grav * (m->alpha * m->f[3][o] - m->beta * m->f[4][o]); }
static void hydrostatic_pressure(const M* m) {
#pragma omp parallel for
    for (int j = -1; j <= m->Ny; ++j)
        for (int i = -1; i <= m->Nx; ++i) {
            long o = IDX(m, i, j, m->Nz);
            real bup = buoy(m, o), p = 0;
            for (int k = m->Nz - 1; k >= 0; --k) {
                o -= m->sz;
                real bk = buoy(m, o), bf = 0.5 * (bk + bup);
                p = (k == m->Nz - 1) ? -bf * m->d[2] : p - bf * m->d[2];
                m->pHY[o] = p;
                bup = bk;
            }
        }
}

/* ---- compute_tendencies!  compute_nonhydrostatic_tendencies.jl:18-163 ---------------------------------------------------- */
static void tendencies(const M* m) {
    const int nf = 3 + m->ntr;
    if (m->ntr) hydrostatic_pressure(m);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < m->Nz; ++k)
        for (int j = 0; j < m->Ny; ++j)
            for (int i = 0; i < m->Nx; ++i) {
                const long o = IDX(m, i, j, k);
                for (int c = 0; c < 3; ++c) {
                    real div = 0;
                    for (int d = 0; d < 3; ++d) {
                        const long sd = ST(m, d);
                        /* flux indices: centre-type (d == c): o and o - sd ; face-type: o + sd and o */
                        const long hi = d == c ? o : o + sd, lo = d == c ? o - sd : o;
                        div += (mom_flux(m, c, d, hi) + visc_flux(m, c, d, hi)) - (mom_flux(m, c, d, lo) + visc_flux(m, c, d, lo));
                    }
                    real G = -(m->rV * div);
                    if (m->ntr && c < 2) G = G - (m->pHY[o] - m->pHY[o - ST(m, c)]) * m->rd[c];
                    m->Gn[c][o] = G;
                }
                for (int t = 3; t < nf; ++t) {
                    real div = 0;
                    for (int d = 0; d < 3; ++d) div += tr_flux(m, t, d, o + ST(m, d)) - tr_flux(m, t, d, o);
                    m->Gn[t][o] = -(m->rV * div);
                }
            }
}

/* ---- FFT (radix-2 when possible, naive DFT otherwise): stands in for FFTW's plan_fft! (src/Solvers/plan_transforms.jl:16-34) - */
static void fft_line(real* re, real* im, int n, int inverse, real* wr, real* wi, const real* tc, const real* ts) {
    /* tc[q] = cos(2πq/n), ts[q] = sin(2πq/n): the twiddle table of this line length (built once per transform direction) */
    const real sg = inverse ? 1 : -1;
    if ((n & (n - 1)) == 0) {
        for (int i = 1, j = 0; i < n; ++i) {
            int bit = n >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) { real t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
        }
        for (int len = 2; len <= n; len <<= 1) {
            const int step = n / len;
            for (int i = 0; i < n; i += len)
                for (int k = 0; k < len / 2; ++k) {
                    real c = tc[k * step], s = sg * ts[k * step];
                    real ur = re[i + k], ui = im[i + k];
                    real vr = re[i + k + len / 2] * c - im[i + k + len / 2] * s, vi = re[i + k + len / 2] * s + im[i + k + len / 2] * c;
                    re[i + k] = ur + vr; im[i + k] = ui + vi;
                    re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
                }
        }
    } else {
        for (int q = 0; q < n; ++q) {
            real sr = 0, si = 0;
            for (int t = 0; t < n; ++t) {
                const int a = (int)(((long)q * t) % n);
                const real c = tc[a], s = sg * ts[a];
                sr += re[t] * c - im[t] * s;
                si += re[t] * s + im[t] * c;
            }
            wr[q] = sr; wi[q] = si;
        }
        memcpy(re, wr, sizeof(real) * n); memcpy(im, wi, sizeof(real) * n);
    }
}
static void fft3(const M* m, int inverse) {
    const int N[3] = {m->Nx, m->Ny, m->Nz};
    const long st[3] = {1, m->Nx, (long)m->Nx * m->Ny};
    for (int d = 0; d < 3; ++d) {
        const int n = N[d], a = (d + 1) % 3, b = (d + 2) % 3;
        real* tc = (real*)malloc(sizeof(real) * 2 * n);
        real* ts = tc + n;
        for (int q = 0; q < n; ++q) { tc[q] = COS(2 * PI_ * q / n); ts[q] = SIN(2 * PI_ * q / n); }
#pragma omp parallel
        {
            real* lr = (real*)malloc(sizeof(real) * 4 * n);
            real *li = lr + n, *wr = lr + 2 * n, *wi = lr + 3 * n;
#pragma omp for collapse(2)
            for (int ib = 0; ib < N[b]; ++ib)
                for (int ia = 0; ia < N[a]; ++ia) {
                    const long base = ia * st[a] + ib * st[b];
                    for (int t = 0; t < n; ++t) { lr[t] = m->re[base + t * st[d]]; li[t] = m->im[base + t * st[d]]; }
                    fft_line(lr, li, n, inverse, wr, wi, tc, ts);
                    const real sc = inverse ? 1.0 / n : 1.0;
                    for (int t = 0; t < n; ++t) { m->re[base + t * st[d]] = lr[t] * sc; m->im[base + t * st[d]] = li[t] * sc; }
                }
            free(lr);
        }
        free(tc);
    }
}

/* ---- compute_pressure_correction! + make_pressure_correction!  pressure_correction.jl:8-53, solve_for_pressure.jl:12-18,
 *      fft_based_poisson_solver.jl:95-125 ------------------------------------------------------------------------------------ */
static void pressure_correct(M* m, real dt) {
    for (int c = 0; c < 3; ++c) fill_halo(m, m->f[c]);
    const int Nx = m->Nx, Ny = m->Ny, Nz = m->Nz;
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                const long o = IDX(m, i, j, k), s = i + (long)Nx * (j + (long)Ny * k);
                real dx = m->A[0] * m->f[0][o + 1] - m->A[0] * m->f[0][o];
                real dy = m->A[1] * m->f[1][o + m->sy] - m->A[1] * m->f[1][o];
                real dz = m->A[2] * m->f[2][o + m->sz] - m->A[2] * m->f[2][o];
                m->re[s] = m->rV * (dx + dy + dz);
                m->im[s] = 0;
            }
    fft3(m, 0);
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                const long s = i + (long)Nx * (j + (long)Ny * k);
                const real l = m->lam[0][i] + m->lam[1][j] + m->lam[2][k];
                if (i == 0 && j == 0 && k == 0) { m->re[s] = 0; m->im[s] = 0; }
                else { m->re[s] = -m->re[s] / l; m->im[s] = -m->im[s] / l; }
            }
    fft3(m, 1);
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) m->p[IDX(m, i, j, k)] = m->re[i + (long)Nx * (j + (long)Ny * k)];
    fill_halo(m, m->p);
    const real dtp = dt > 2.220446049250313e-16 ? dt : 2.220446049250313e-16;
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                const long o = IDX(m, i, j, k);
                m->f[0][o] -= (m->p[o] - m->p[o - 1]) * m->rd[0];
                m->f[1][o] -= (m->p[o] - m->p[o - m->sy]) * m->rd[1];
                m->f[2][o] -= (m->p[o] - m->p[o - m->sz]) * m->rd[2];
            }
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) { const long o = IDX(m, i, j, k); m->p[o] = m->p[o] / dtp; }
}

static void update_state(M* m) {
    for (int f = 0; f < 3 + m->ntr; ++f) fill_halo(m, m->f[f]);
    tendencies(m);
}

/* ---- C interface ------------------------------------------------------------------------------------------------------------ */
void* nhc_create(int Nx, int Ny, int Nz, double Lx, double Ly, double Lz, int weno, int ntr, double nu, double kappa,
                 double grav, double alpha, double beta) {
    M* m = (M*)calloc(1, sizeof(M));
    m->Nx = Nx; m->Ny = Ny; m->Nz = Nz; m->weno = weno; m->ntr = ntr;
    m->sy = Nx + 2 * H; m->sz = m->sy * (Ny + 2 * H); m->n = m->sz * (Nz + 2 * H);
    m->L[0] = Lx; m->L[1] = Ly; m->L[2] = Lz;
    m->d[0] = (double)(Lx / Nx); m->d[1] = (double)(Ly / Ny); m->d[2] = (double)(Lz / Nz);   /* the grid spacings are Float64 numbers in both builds */
    for (int d = 0; d < 3; ++d) m->rd[d] = 1 / m->d[d];
    m->A[0] = m->d[1] * m->d[2]; m->A[1] = m->d[0] * m->d[2]; m->A[2] = m->d[0] * m->d[1];
    m->V = m->A[2] * m->d[2]; m->rV = 1 / m->V;
    m->nu = nu; m->kappa = kappa; m->grav = grav; m->alpha = alpha; m->beta = beta;
    for (int f = 0; f < 5; ++f) { m->f[f] = (real*)calloc(m->n, sizeof(real)); m->Gn[f] = (real*)calloc(m->n, sizeof(real)); m->Gm[f] = (real*)calloc(m->n, sizeof(real)); }
    m->pHY = (real*)calloc(m->n, sizeof(real)); m->p = (real*)calloc(m->n, sizeof(real));
    m->re = (real*)calloc((size_t)Nx * Ny * Nz, sizeof(real)); m->im = (real*)calloc((size_t)Nx * Ny * Nz, sizeof(real));
    const int N[3] = {Nx, Ny, Nz};
    for (int d = 0; d < 3; ++d) {     /* poisson_eigenvalues.jl:8-31 */
        m->lam[d] = (real*)calloc(N[d], sizeof(real));
        for (int i = 0; i < N[d]; ++i) m->lam[d][i] = POW(2 * SIN(i * PI_ / N[d]) / (m->L[d] / N[d]), 2);
    }
    /* coefficients: rationals rounded to Float64, the last of each set is 1 - sum(others) */
    const double k0 = -1.0 / 12, k1 = 7.0 / 12, k2 = 7.0 / 12, k3 = 1 - ((k0 + k1) + k2);
    m->c4[0] = k3; m->c4[1] = k2; m->c4[2] = k1; m->c4[3] = k0;
    const double p[3][2] = {{1.0 / 3, 5.0 / 6}, {-1.0 / 6, 5.0 / 6}, {1.0 / 3, -7.0 / 6}};
    for (int r = 0; r < 3; ++r) { m->w5p[r][0] = p[r][0]; m->w5p[r][1] = p[r][1]; m->w5p[r][2] = 1 - (p[r][0] + p[r][1]); }
    m->gam[0] = 8.0 / 15; m->gam[1] = 5.0 / 12; m->gam[2] = 3.0 / 4;     /* runge_kutta_3.jl:69-78 */
    m->zet[0] = 0; m->zet[1] = -17.0 / 60; m->zet[2] = -5.0 / 12;
    return m;
}
void nhc_set_beta_difference_form(void* h, int on) { ((M*)h)->beta_difference_form = on; }
void nhc_destroy(void* h) {
    M* m = (M*)h;
    for (int f = 0; f < 5; ++f) { free(m->f[f]); free(m->Gn[f]); free(m->Gm[f]); }
    free(m->pHY); free(m->p); free(m->re); free(m->im);
    for (int d = 0; d < 3; ++d) free(m->lam[d]);
    free(m);
}
static real* field_of(M* m, int f) { return f == 5 ? m->p : m->f[f]; }
void nhc_set(void* h, int f, const double* a) {
    M* m = (M*)h;
    for (int k = 0; k < m->Nz; ++k) for (int j = 0; j < m->Ny; ++j) for (int i = 0; i < m->Nx; ++i)
        field_of(m, f)[IDX(m, i, j, k)] = (real)a[i + (long)m->Nx * (j + (long)m->Ny * k)];
}
void nhc_get(void* h, int f, double* a) {
    M* m = (M*)h;
    for (int k = 0; k < m->Nz; ++k) for (int j = 0; j < m->Ny; ++j) for (int i = 0; i < m->Nx; ++i)
        a[i + (long)m->Nx * (j + (long)m->Ny * k)] = (double)field_of(m, f)[IDX(m, i, j, k)];
}
/* tail of set!(model; …): projection with Δt = 1, then update_state!  (set_nonhydrostatic_model.jl:44-57) */
void nhc_finalize(void* h) {
    M* m = (M*)h;
    pressure_correct(m, 1.0);
    update_state(m);
}
/* time_step!(model::AbstractModel{<:RungeKutta3TimeStepper}, Δt)  runge_kutta_3.jl:93-170 */
void nhc_step(void* h, double dt_in) {
    M* m = (M*)h;
    const real dt = (real)dt_in;
    const int nf = 3 + m->ntr;
    for (int s = 0; s < 3; ++s) {
        const real g = m->gam[s], z = m->zet[s];
        for (int f = 0; f < nf; ++f) {
#pragma omp parallel for collapse(2)
            for (int k = 0; k < m->Nz; ++k)
                for (int j = 0; j < m->Ny; ++j)
                    for (int i = 0; i < m->Nx; ++i) {
                        const long o = IDX(m, i, j, k);
                        if (s == 0) m->f[f][o] = m->f[f][o] + dt * g * m->Gn[f][o];
                        else m->f[f][o] = m->f[f][o] + dt * (g * m->Gn[f][o] + z * m->Gm[f][o]);
                    }
        }
        pressure_correct(m, dt * (g + z));
        if (s < 2) for (int f = 0; f < nf; ++f) { real* t = m->Gm[f]; m->Gm[f] = m->Gn[f]; m->Gn[f] = t; }
        update_state(m);
    }
}
