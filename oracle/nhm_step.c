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
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define H 3

typedef struct {
    int Nx, Ny, Nz, weno, ntr;
    long sy, sz, n;
    double d[3], rd[3], A[3], V, rV, L[3];
    double nu, kappa, grav, alpha, beta;
    double *f[5], *Gn[5], *Gm[5], *pHY, *p;   /* u v w T S */
    double *lam[3];
    double *re, *im;                          /* spectral work arrays, Nx*Ny*Nz */
    /* reconstruction coefficients (src/Advection/reconstruction_coefficients.jl:49-64) */
    double c4[4], w5p[3][3];
    double gam[3], zet[3];
} M;

static long IDX(const M* m, int i, int j, int k) { return (i + H) + (long)(j + H) * m->sy + (long)(k + H) * m->sz; }
static long ST(const M* m, int d) { return d == 0 ? 1 : (d == 1 ? m->sy : m->sz); }

/* ---- fill_halo_regions! for Periodic sides: src/BoundaryConditions/fill_halo_regions_periodic.jl:5-32 ---------------- */
static void fill_halo(const M* m, double* f) {
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
static double newton_div(double a, double b) {
    double inv = (double)(1.0f / (float)b);
    double x = a * inv;
    return fma(fma(x, -b, a), inv, x);
}
static double beta3(double a, double b, double c, double C1, double C2, double C3, double C4, double C5, double C6) {
    return a * (C1 * a + C2 * b + C3 * c) + b * (C4 * b + C5 * c) + c * c * C6;   /* weno_interpolants.jl:204-216,261 */
}
/* WENO{3} biased reconstruction at the face whose right cell is p[0] (ψ[i]); stride s  (weno_interpolants.jl:290-337,409-437,500) */
static double weno5(const M* m, const double* p, long s, int left) {
    double q0, q1, q2, q3, q4;
    if (left) { q0 = p[-3 * s]; q1 = p[-2 * s]; q2 = p[-s]; q3 = p[0]; q4 = p[s]; }
    else      { q0 = p[2 * s];  q1 = p[s];      q2 = p[0];  q3 = p[-s]; q4 = p[-2 * s]; }
    double b0 = beta3(q2, q3, q4, 10, -31, 11, 25, -19, 4);
    double b1 = beta3(q1, q2, q3, 4, -13, 5, 13, -13, 4);
    double b2 = beta3(q0, q1, q2, 4, -19, 11, 25, -31, 10);
    const double eps = (double)1e-8f;
    double tau = fabs(b0 - b2);
    double r0 = newton_div(tau, b0 + eps), r1 = newton_div(tau, b1 + eps), r2 = newton_div(tau, b2 + eps);
    double a0 = 0.3 * (1 + r0 * r0), a1 = 0.6 * (1 + r1 * r1), a2 = 0.1 * (1 + r2 * r2);
    double rs = 1 / (a0 + a1 + a2);
    double p0 = m->w5p[0][0] * q2 + m->w5p[0][1] * q3 + m->w5p[0][2] * q4;
    double p1 = m->w5p[1][0] * q1 + m->w5p[1][1] * q2 + m->w5p[1][2] * q3;
    double p2 = m->w5p[2][0] * q0 + m->w5p[2][1] * q1 + m->w5p[2][2] * q2;
    return (a0 * rs) * p0 + (a1 * rs) * p1 + (a2 * rs) * p2;
}
/* Centered(order=4) of a·q at the face whose right cell is p[0]  (centered_reconstruction.jl:47-55) */
static double sym4(const M* m, const double* p, long s, double a) {
    double r = m->c4[0] * (a * p[-2 * s]);
    r = r + m->c4[1] * (a * p[-s]);
    r = r + m->c4[2] * (a * p[0]);
    r = r + m->c4[3] * (a * p[s]);
    return r;
}

/* advective flux of momentum component c through the face normal to d at flux index o (centre-type when d == c)
 * src/Advection/upwind_biased_advective_fluxes.jl:23-93, centered_advective_fluxes.jl:15-27 */
static double mom_flux(const M* m, int c, int d, long o) {
    const double* psi = m->f[c] + o;
    const double* adv = m->f[d] + o;
    const long sd = ST(m, d), sc = ST(m, c);
    const double A = m->A[d];
    if (d == c) {
        if (!m->weno) { double ut = 0.5 * adv[0] + 0.5 * adv[sd], pt = 0.5 * psi[0] + 0.5 * psi[sd]; return A * ut * pt; }
        double ut = sym4(m, adv + sd, sd, A);
        return ut * weno5(m, psi + sd, sd, ut > 0);
    }
    if (!m->weno) { double ut = 0.5 * adv[-sc] + 0.5 * adv[0], pt = 0.5 * psi[-sd] + 0.5 * psi[0]; return A * ut * pt; }
    double ut = sym4(m, adv, sc, A);
    return ut * weno5(m, psi, sd, ut > 0);
}
/* viscous flux A_d τ_cd, τ = -2 ν Σ  (src/TurbulenceClosures/abstract_scalar_diffusivity_closure.jl:189-204,
 * velocity_tracer_gradients.jl:25-42, closure_kernel_operators.jl:22-41) */
static double visc_flux(const M* m, int c, int d, long o) {
    double sig;
    if (d == c) sig = (m->f[d][o + ST(m, d)] - m->f[d][o]) * m->rd[d];
    else {
        int lo = d < c ? d : c, hi = d < c ? c : d;
        double dl = (m->f[lo][o] - m->f[lo][o - ST(m, hi)]) * m->rd[hi];
        double dh = (m->f[hi][o] - m->f[hi][o - ST(m, lo)]) * m->rd[lo];
        sig = 0.5 * (dl + dh);
    }
    return m->A[d] * (-2 * (m->nu * sig));
}
/* tracer fluxes  (upwind_biased_advective_fluxes.jl:99-121, centered_advective_fluxes.jl:31-33; :240-242) */
static double tr_flux(const M* m, int t, int d, long o) {
    const double* c = m->f[t] + o;
    const long sd = ST(m, d);
    const double u = m->f[d][o], A = m->A[d];
    double F;
    if (!m->weno) F = (A * u) * (0.5 * c[-sd] + 0.5 * c[0]);
    else F = A * u * weno5(m, c, sd, u > 0);
    double grad = (c[0] - c[-sd]) * m->rd[d];
    return F + m->A[d] * (-(m->kappa * grad));
}

/* ---- update_hydrostatic_pressure!  update_hydrostatic_pressure.jl:12-49 ------------------------------------------------ */
static double buoy(const M* m, long o) { return m->grav * (m->alpha * m->f[3][o] - m->beta * m->f[4][o]); }
static void hydrostatic_pressure(const M* m) {
#pragma omp parallel for
    for (int j = -1; j <= m->Ny; ++j)
        for (int i = -1; i <= m->Nx; ++i) {
            long o = IDX(m, i, j, m->Nz);
            double bup = buoy(m, o), p = 0;
            for (int k = m->Nz - 1; k >= 0; --k) {
                o -= m->sz;
                double bk = buoy(m, o), bf = 0.5 * (bk + bup);
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
                    double div = 0;
                    for (int d = 0; d < 3; ++d) {
                        const long sd = ST(m, d);
                        /* flux indices: centre-type (d == c): o and o - sd ; face-type: o + sd and o */
                        const long hi = d == c ? o : o + sd, lo = d == c ? o - sd : o;
                        div += (mom_flux(m, c, d, hi) + visc_flux(m, c, d, hi)) - (mom_flux(m, c, d, lo) + visc_flux(m, c, d, lo));
                    }
                    double G = -(m->rV * div);
                    if (m->ntr && c < 2) G = G - (m->pHY[o] - m->pHY[o - ST(m, c)]) * m->rd[c];
                    m->Gn[c][o] = G;
                }
                for (int t = 3; t < nf; ++t) {
                    double div = 0;
                    for (int d = 0; d < 3; ++d) div += tr_flux(m, t, d, o + ST(m, d)) - tr_flux(m, t, d, o);
                    m->Gn[t][o] = -(m->rV * div);
                }
            }
}

/* ---- FFT (radix-2 when possible, naive DFT otherwise): stands in for FFTW's plan_fft! (src/Solvers/plan_transforms.jl:16-34) - */
static void fft_line(double* re, double* im, int n, int inverse, double* wr, double* wi) {
    if ((n & (n - 1)) == 0) {
        for (int i = 1, j = 0; i < n; ++i) {
            int bit = n >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
        }
        for (int len = 2; len <= n; len <<= 1) {
            double ang = 2 * M_PI / len * (inverse ? 1 : -1);
            for (int i = 0; i < n; i += len)
                for (int k = 0; k < len / 2; ++k) {
                    double c = cos(ang * k), s = sin(ang * k);
                    double ur = re[i + k], ui = im[i + k];
                    double vr = re[i + k + len / 2] * c - im[i + k + len / 2] * s, vi = re[i + k + len / 2] * s + im[i + k + len / 2] * c;
                    re[i + k] = ur + vr; im[i + k] = ui + vi;
                    re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
                }
        }
    } else {
        for (int q = 0; q < n; ++q) {
            double sr = 0, si = 0;
            for (int t = 0; t < n; ++t) {
                double ang = 2 * M_PI * (double)(((long)q * t) % n) / n * (inverse ? 1 : -1);
                sr += re[t] * cos(ang) - im[t] * sin(ang);
                si += re[t] * sin(ang) + im[t] * cos(ang);
            }
            wr[q] = sr; wi[q] = si;
        }
        memcpy(re, wr, sizeof(double) * n); memcpy(im, wi, sizeof(double) * n);
    }
}
static void fft3(const M* m, int inverse) {
    const int N[3] = {m->Nx, m->Ny, m->Nz};
    const long st[3] = {1, m->Nx, (long)m->Nx * m->Ny};
    for (int d = 0; d < 3; ++d) {
        const int n = N[d], a = (d + 1) % 3, b = (d + 2) % 3;
#pragma omp parallel
        {
            double* lr = (double*)malloc(sizeof(double) * 4 * n);
            double *li = lr + n, *wr = lr + 2 * n, *wi = lr + 3 * n;
#pragma omp for collapse(2)
            for (int ib = 0; ib < N[b]; ++ib)
                for (int ia = 0; ia < N[a]; ++ia) {
                    const long base = ia * st[a] + ib * st[b];
                    for (int t = 0; t < n; ++t) { lr[t] = m->re[base + t * st[d]]; li[t] = m->im[base + t * st[d]]; }
                    fft_line(lr, li, n, inverse, wr, wi);
                    const double sc = inverse ? 1.0 / n : 1.0;
                    for (int t = 0; t < n; ++t) { m->re[base + t * st[d]] = lr[t] * sc; m->im[base + t * st[d]] = li[t] * sc; }
                }
            free(lr);
        }
    }
}

/* ---- compute_pressure_correction! + make_pressure_correction!  pressure_correction.jl:8-53, solve_for_pressure.jl:12-18,
 *      fft_based_poisson_solver.jl:95-125 ------------------------------------------------------------------------------------ */
static void pressure_correct(M* m, double dt) {
    for (int c = 0; c < 3; ++c) fill_halo(m, m->f[c]);
    const int Nx = m->Nx, Ny = m->Ny, Nz = m->Nz;
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                const long o = IDX(m, i, j, k), s = i + (long)Nx * (j + (long)Ny * k);
                double dx = m->A[0] * m->f[0][o + 1] - m->A[0] * m->f[0][o];
                double dy = m->A[1] * m->f[1][o + m->sy] - m->A[1] * m->f[1][o];
                double dz = m->A[2] * m->f[2][o + m->sz] - m->A[2] * m->f[2][o];
                m->re[s] = m->rV * (dx + dy + dz);
                m->im[s] = 0;
            }
    fft3(m, 0);
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                const long s = i + (long)Nx * (j + (long)Ny * k);
                const double l = m->lam[0][i] + m->lam[1][j] + m->lam[2][k];
                if (i == 0 && j == 0 && k == 0) { m->re[s] = 0; m->im[s] = 0; }
                else { m->re[s] = -m->re[s] / l; m->im[s] = -m->im[s] / l; }
            }
    fft3(m, 1);
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) m->p[IDX(m, i, j, k)] = m->re[i + (long)Nx * (j + (long)Ny * k)];
    fill_halo(m, m->p);
    const double dtp = dt > 2.220446049250313e-16 ? dt : 2.220446049250313e-16;
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
    m->d[0] = Lx / Nx; m->d[1] = Ly / Ny; m->d[2] = Lz / Nz;
    for (int d = 0; d < 3; ++d) m->rd[d] = 1 / m->d[d];
    m->A[0] = m->d[1] * m->d[2]; m->A[1] = m->d[0] * m->d[2]; m->A[2] = m->d[0] * m->d[1];
    m->V = m->A[2] * m->d[2]; m->rV = 1 / m->V;
    m->nu = nu; m->kappa = kappa; m->grav = grav; m->alpha = alpha; m->beta = beta;
    for (int f = 0; f < 5; ++f) { m->f[f] = (double*)calloc(m->n, 8); m->Gn[f] = (double*)calloc(m->n, 8); m->Gm[f] = (double*)calloc(m->n, 8); }
    m->pHY = (double*)calloc(m->n, 8); m->p = (double*)calloc(m->n, 8);
    m->re = (double*)calloc((size_t)Nx * Ny * Nz, 8); m->im = (double*)calloc((size_t)Nx * Ny * Nz, 8);
    const int N[3] = {Nx, Ny, Nz};
    for (int d = 0; d < 3; ++d) {     /* poisson_eigenvalues.jl:8-31 */
        m->lam[d] = (double*)calloc(N[d], 8);
        for (int i = 0; i < N[d]; ++i) m->lam[d][i] = pow(2 * sin(i * M_PI / N[d]) / (m->L[d] / N[d]), 2);
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
void nhc_destroy(void* h) {
    M* m = (M*)h;
    for (int f = 0; f < 5; ++f) { free(m->f[f]); free(m->Gn[f]); free(m->Gm[f]); }
    free(m->pHY); free(m->p); free(m->re); free(m->im);
    for (int d = 0; d < 3; ++d) free(m->lam[d]);
    free(m);
}
static double* field_of(M* m, int f) { return f == 5 ? m->p : m->f[f]; }
void nhc_set(void* h, int f, const double* a) {
    M* m = (M*)h;
    for (int k = 0; k < m->Nz; ++k) for (int j = 0; j < m->Ny; ++j) for (int i = 0; i < m->Nx; ++i)
        field_of(m, f)[IDX(m, i, j, k)] = a[i + (long)m->Nx * (j + (long)m->Ny * k)];
}
void nhc_get(void* h, int f, double* a) {
    M* m = (M*)h;
    for (int k = 0; k < m->Nz; ++k) for (int j = 0; j < m->Ny; ++j) for (int i = 0; i < m->Nx; ++i)
        a[i + (long)m->Nx * (j + (long)m->Ny * k)] = field_of(m, f)[IDX(m, i, j, k)];
}
/* tail of set!(model; …): projection with Δt = 1, then update_state!  (set_nonhydrostatic_model.jl:44-57) */
void nhc_finalize(void* h) {
    M* m = (M*)h;
    pressure_correct(m, 1.0);
    update_state(m);
}
/* time_step!(model::AbstractModel{<:RungeKutta3TimeStepper}, Δt)  runge_kutta_3.jl:93-170 */
void nhc_step(void* h, double dt) {
    M* m = (M*)h;
    const int nf = 3 + m->ntr;
    for (int s = 0; s < 3; ++s) {
        const double g = m->gam[s], z = m->zet[s];
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
        if (s < 2) for (int f = 0; f < nf; ++f) { double* t = m->Gm[f]; m->Gm[f] = m->Gn[f]; m->Gn[f] = t; }
        update_state(m);
    }
}
