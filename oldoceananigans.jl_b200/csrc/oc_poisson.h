// oc_poisson.h — kernels around the FFTs of the pressure solve.
//
// Replaces, for FFTBasedPoissonSolver (src/Solvers/fft_based_poisson_solver.jl:95-137):
//   _compute_source_term!            src/Models/NonhydrostaticModels/solve_for_pressure.jl:12-18
//   permute_*/unpermute_* + copyto! + twiddle broadcasts (Makhoul DCT)   src/Solvers/index_permutations.jl:38-90,
//                                                                       discrete_transforms.jl:108-176
//   ϕc = -b/(λx+λy+λz) ; ϕc[1,1,1] = 0                                 fft_based_poisson_solver.jl:110,115
//   copy_real_component! + _make_pressure_correction! + pNHS ./= Δt     :122-137, pressure_correction.jl:31-53
//
// Scheme (validated against scipy's DCT-II/III in tests/test_transform_math.py):
//   rhs kernel   : div(u*,v*,w*) written at the Makhoul-permuted position in every Bounded dimension
//                  (v[n] = x[2n], v[N-1-n] = x[2n+1]); real-to-complex layout when x is not Bounded.
//   forward FFT  : ONE multi-dimensional FFT over all non-trivial dimensions (cuFFT).
//   mid kernel   : per reflection orbit {k, N-k} of each Bounded dimension, in registers:
//                  X[k] = ω_k V[k] + conj(ω_k) V[N-k]   (ω_k = exp(-iπk/2N); DCT-II, FFTW REDFT10 scaling)
//                  Φ = -X / (λx+λy+λz), Φ[0,0,0] = 0, times 1/(Nx·Ny·Nz)
//                  W[k] = ½ conj(ω_k) (Φ[k] - i Φ[N-k]),  Φ[N] := 0   (DCT-III · 1/2N folded in)
//   inverse FFT  : ONE multi-dimensional inverse FFT.
//   projection   : reads ϕ = Δt·p at un-permuted positions, U -= ∇ϕ, pNHS = ϕ/Δt — no separate
//                  copy-real / halo-fill / scale passes.
#pragma once
#include "oc_common.h"

namespace oc {

template <class FT>
struct Cplx {
    FT x, y;
};

struct Cd {
    double x, y;
};
OC_HD Cd cmul(Cd a, Cd b) { return Cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
OC_HD Cd cconj(Cd a) { return Cd{a.x, -a.y}; }
OC_HD Cd cadd(Cd a, Cd b) { return Cd{a.x + b.x, a.y + b.y}; }

// Layout of the transform buffer.
struct SpectralLayout {
    int N[3];          // physical sizes
    int bounded[3];
    int r2c;           // 1: real rows padded to 2*(Nx/2+1) reals, complex rows Nx/2+1; 0: full complex, Nx per row
    int nxc;           // complex row length
    int nxr;           // real row length in reals (r2c) — for c2c the real value sits at 2*i
    OC_HD long long real_index(int i, int j, int k) const {
        return r2c ? ((long long)i + (long long)nxr * (j + (long long)N[1] * k))
                   : 2 * ((long long)i + (long long)N[0] * (j + (long long)N[1] * k));
    }
    OC_HD long long cplx_index(int i, int j, int k) const { return (long long)i + (long long)nxc * (j + (long long)N[1] * k); }
};

// Makhoul permutation: x[z] goes to position z/2 (z even) or N-1-(z-1)/2 (z odd)   index_permutations.jl:18-36
OC_HD int makhoul(int z, int N, int bounded) {
    if (!bounded) return z;
    return (z & 1) ? N - 1 - (z >> 1) : (z >> 1);
}

template <class FT>
struct PoissonRhsKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    SpectralLayout L;
    const FT* u;
    const FT* v;
    const FT* w;
    FT* buf;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        int o = g.idx(i, j, k);
        // divᶜᶜᶜ = V⁻¹ (δxᶜ(Ax u) + δyᶜ(Ay v) + δzᶜ(Az w))      divergence_operators.jl:16-19
        const FT Ax = g.area_at(0, false, k), Ay = g.area_at(1, false, k);
        FT dx = g.flat[0] ? FT(0) : Ax * u[o + 1] - Ax * u[o];
        FT dy = g.flat[1] ? FT(0) : Ay * v[o + g.sy] - Ay * v[o];
        FT dz = g.flat[2] ? FT(0) : g.A[2] * w[o + g.sz] - g.A[2] * w[o];
        FT div = g.rV_at(false, k) * (dx + dy + dz);
        // stretched grid: the tridiagonal system is the Poisson equation times Δzᶜᶜᶜ
        // (_fourier_tridiagonal_source_term!, solve_for_pressure.jl:36-42); z is not transformed (L.bounded[2] = 0)
        if (g.stretched()) div = g.dzc[k] * div;
        long long r = L.real_index(makhoul(i, L.N[0], L.bounded[0]), makhoul(j, L.N[1], L.bounded[1]),
                                   makhoul(k, L.N[2], L.bounded[2]));
        buf[r] = div;
        if (!L.r2c) buf[r + 1] = FT(0);
    }
};

// Host rhs (dense Nx×Ny×Nz) -> buffer, for the stand-alone solve!(ϕ, solver, b) entry point.
template <class FT>
struct PoissonLoadKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    const FT* rhs;
    FT* buf;
    const FT* dzc;     // stretched grid: set_source_term! multiplies by Δzᶜᶜᶜ (fourier_tridiagonal_poisson_solver.jl:239-246); else nullptr
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= L.N[0]) return;
        long long r = L.real_index(makhoul(i, L.N[0], L.bounded[0]), makhoul(j, L.N[1], L.bounded[1]),
                                   makhoul(k, L.N[2], L.bounded[2]));
        FT val = rhs[(long long)i + (long long)L.N[0] * (j + (long long)L.N[1] * k)];
        buf[r] = dzc ? val * dzc[k] : val;
        if (!L.r2c) buf[r + 1] = FT(0);
    }
};

template <class FT>
struct PoissonMidKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    Cplx<FT>* spec;
    const double* lam[3];      // eigenvalues, Float64   poisson_eigenvalues.jl:8-31
    const Cd* tw[3];           // ω_k = exp(-iπk/2N) for Bounded dims (nullptr otherwise)
    int nrep[3];               // number of orbit representatives per dim
    double norm;               // 1/(Nx Ny Nz)

    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int rx = b.x * nt + tid, ry = b.y, rz = b.z;
        if (rx >= nrep[0]) return;
        // orbit members per dim: index a and its reflection (N - a) % N (only for Bounded dims)
        int idx[3][2];
        int rep[3] = {rx, ry, rz};
        int np[3];
        for (int d = 0; d < 3; ++d) {
            idx[d][0] = rep[d];
            if (L.bounded[d]) { idx[d][1] = (L.N[d] - rep[d]) % L.N[d]; np[d] = 2; }
            else { idx[d][1] = rep[d]; np[d] = 1; }
        }
        Cd v[2][2][2];
        for (int c = 0; c < 2; ++c)
            for (int bb = 0; bb < 2; ++bb)
                for (int aa = 0; aa < 2; ++aa) {
                    if (aa < np[0] && bb < np[1] && c < np[2]) {
                        Cplx<FT> e = spec[L.cplx_index(idx[0][aa], idx[1][bb], idx[2][c])];
                        v[c][bb][aa] = Cd{(double)e.x, (double)e.y};
                    } else {
                        v[c][bb][aa] = Cd{0.0, 0.0};
                    }
                }
        // forward post-twiddle along every Bounded dim: X[k] = ω_k V[k] + conj(ω_k) V[N-k]
        for (int d = 0; d < 3; ++d) {
            if (!L.bounded[d]) continue;
            Cd w0 = tw[d][idx[d][0]], w1 = tw[d][idx[d][1]];
            for (int p = 0; p < 2; ++p)
                for (int q = 0; q < 2; ++q) {
                    Cd* e0 = d == 0 ? &v[p][q][0] : (d == 1 ? &v[p][0][q] : &v[0][p][q]);
                    Cd* e1 = d == 0 ? &v[p][q][1] : (d == 1 ? &v[p][1][q] : &v[1][p][q]);
                    Cd a0 = *e0, a1 = *e1;
                    *e0 = cadd(cmul(w0, a0), cmul(cconj(w0), a1));
                    *e1 = cadd(cmul(w1, a1), cmul(cconj(w1), a0));
                }
        }
        // eigenvalue divide
        for (int c = 0; c < np[2]; ++c)
            for (int bb = 0; bb < np[1]; ++bb)
                for (int aa = 0; aa < np[0]; ++aa) {
                    int kx = idx[0][aa], ky = idx[1][bb], kz = idx[2][c];
                    double l = lam[0][kx] + lam[1][ky] + lam[2][kz];
                    Cd e = v[c][bb][aa];
                    if (kx == 0 && ky == 0 && kz == 0) e = Cd{0.0, 0.0};
                    else { double s = -norm / l; e.x *= s; e.y *= s; }
                    v[c][bb][aa] = e;
                }
        // inverse pre-twiddle: W[k] = ½ conj(ω_k) (Φ[k] - i Φ[N-k]), Φ[N] := 0
        for (int d = 0; d < 3; ++d) {
            if (!L.bounded[d]) continue;
            Cd w0 = tw[d][idx[d][0]], w1 = tw[d][idx[d][1]];
            bool z0 = idx[d][0] == 0, z1 = idx[d][1] == 0;
            for (int p = 0; p < 2; ++p)
                for (int q = 0; q < 2; ++q) {
                    Cd* e0 = d == 0 ? &v[p][q][0] : (d == 1 ? &v[p][0][q] : &v[0][p][q]);
                    Cd* e1 = d == 0 ? &v[p][q][1] : (d == 1 ? &v[p][1][q] : &v[1][p][q]);
                    Cd a0 = *e0, a1 = *e1;
                    Cd r0 = z0 ? Cd{0.0, 0.0} : a1;     // Φ[N - k0]
                    Cd r1 = z1 ? Cd{0.0, 0.0} : a0;     // Φ[N - k1]
                    // (a - i r) = (a.x + r.y, a.y - r.x)
                    Cd t0 = Cd{a0.x + r0.y, a0.y - r0.x}, t1 = Cd{a1.x + r1.y, a1.y - r1.x};
                    Cd h0 = cmul(cconj(w0), t0), h1 = cmul(cconj(w1), t1);
                    *e0 = Cd{0.5 * h0.x, 0.5 * h0.y};
                    *e1 = Cd{0.5 * h1.x, 0.5 * h1.y};
                }
        }
        for (int c = 0; c < np[2]; ++c)
            for (int bb = 0; bb < np[1]; ++bb)
                for (int aa = 0; aa < np[0]; ++aa) {
                    Cd e = v[c][bb][aa];
                    spec[L.cplx_index(idx[0][aa], idx[1][bb], idx[2][c])] = Cplx<FT>{(FT)e.x, (FT)e.y};
                }
    }
};

// No Bounded dimension (plain DFTs): Φ = -X / (λx+λy+λz) / (Nx·Ny·Nz), Φ[0,0,0] = 0 — a pure streaming pass
// (fft_based_poisson_solver.jl:110,115).  A block owns `chunk` consecutive spectral elements of one k-plane.
template <class FT>
struct PoissonDivideKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    Cplx<FT>* spec;
    const double* lam[3];
    double norm;
    int chunk;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int k = b.y;
        const int plane = L.nxc * L.N[1];
        const int n1 = (b.x + 1) * chunk < plane ? (b.x + 1) * chunk : plane;
        const double lz = lam[2][k];
        Cplx<FT>* row = spec + (long long)plane * k;
        for (int n = b.x * chunk + tid; n < n1; n += nt) {
            const int j = n / L.nxc, i = n - j * L.nxc;
            const double l = lam[0][i] + lam[1][j] + lz;
            Cplx<FT> e = row[n];
            double s = -norm / l;
            if (i == 0 && j == 0 && k == 0) s = 0.0;
            row[n] = Cplx<FT>{(FT)((double)e.x * s), (FT)((double)e.y * s)};
        }
    }
};

// The common LES topology (Periodic, Periodic, Bounded): only z carries a DCT.  One thread per (x, y) and reflection orbit
// {kz, Nz-kz}; same operations, in the same order, as PoissonMidKernel restricted to d = 2 — but a pure streaming pass
// (coalesced 16-byte complex loads along x, no local arrays).
template <class FT>
struct PoissonMidZKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    Cplx<FT>* spec;
    const double* lam[3];
    const Cd* twz;
    double norm;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int plane = L.nxc * L.N[1];
        const int n = b.x * nt + tid;
        if (n >= plane) return;
        const int j = n / L.nxc, i = n - j * L.nxc;
        const int k0 = b.y, k1 = (L.N[2] - k0) % L.N[2];
        Cplx<FT>* p0 = spec + (long long)plane * k0 + n;
        Cplx<FT>* p1 = spec + (long long)plane * k1 + n;
        const Cplx<FT> s0 = *p0, s1 = *p1;
        Cd a0{(double)s0.x, (double)s0.y}, a1{(double)s1.x, (double)s1.y};
        const Cd w0 = twz[k0], w1 = twz[k1];
        // forward post-twiddle: X[k] = ω_k V[k] + conj(ω_k) V[N-k]
        Cd e0 = cadd(cmul(w0, a0), cmul(cconj(w0), a1));
        Cd e1 = cadd(cmul(w1, a1), cmul(cconj(w1), a0));
        // eigenvalue divide
        const double lxy = lam[0][i] + lam[1][j];
        if (i == 0 && j == 0 && k0 == 0) e0 = Cd{0.0, 0.0};
        else { const double sc = -norm / (lxy + lam[2][k0]); e0.x *= sc; e0.y *= sc; }
        if (i == 0 && j == 0 && k1 == 0) e1 = Cd{0.0, 0.0};
        else { const double sc = -norm / (lxy + lam[2][k1]); e1.x *= sc; e1.y *= sc; }
        // inverse pre-twiddle: W[k] = ½ conj(ω_k) (Φ[k] - i Φ[N-k]), Φ[N] := 0
        const Cd r0 = k0 == 0 ? Cd{0.0, 0.0} : e1;
        const Cd r1 = k1 == 0 ? Cd{0.0, 0.0} : e0;
        const Cd t0{e0.x + r0.y, e0.y - r0.x}, t1{e1.x + r1.y, e1.y - r1.x};
        const Cd h0 = cmul(cconj(w0), t0), h1 = cmul(cconj(w1), t1);
        *p0 = Cplx<FT>{(FT)(0.5 * h0.x), (FT)(0.5 * h0.y)};
        *p1 = Cplx<FT>{(FT)(0.5 * h1.x), (FT)(0.5 * h1.y)};
    }
};

// ---------------------------------------------------------------------------------------------------------
// FourierTridiagonalPoissonSolver for a vertically stretched grid (src/Solvers/fourier_tridiagonal_poisson_solver.jl:74-246,
// batched_tridiagonal_solver.jl:220-243): transforms in x and y only; per horizontal wavenumber the system
//     ϕ[k-1]/Δzᶠ[k] + D[k] ϕ[k] + ϕ[k+1]/Δzᶠ[k+1] = Δzᶜ[k] b̂[k],   D[k] = -(1/Δzᶠ[k+1] + 1/Δzᶠ[k]) - Δzᶜ[k](λx+λy)
// (homogeneous Neumann: the missing neighbour's term dropped at k = 1, Nz) is solved with the Thomas algorithm.  The matrix
// depends only on the grid, so the elimination factors are tabulated once: R[k] = 1/β_k and T[k] = c_{k-1}/β_{k-1} with
// β_1 = D[1], β_k = D[k] - a_{k-1} T[k] (a = c = 1/Δzᶠ[k+1]).  A pivot that is not "definitely diagonally dominant"
// (|β| ≤ 10 eps, :232) gets R = 0, and so does the last pivot of the singular (λx = λy = 0) column: its free constant is fixed by
// removing the volume mean of ϕ (:222-226), which lives entirely in that column.
// ---------------------------------------------------------------------------------------------------------
template <class FT>
struct TridiagSetupKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    const double* lam[2];
    const FT* dzc;
    const FT* rdzf;
    FT* R;
    FT* T;
    double eps10;
    int zero_col;       // the local column (0, 0) is the (kx, ky) = (0, 0) column (always on one GPU; on slabs: the first rank only)
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int i = b.x * nt + tid, j = b.y;
        if (i >= L.nxc) return;
        const int Nz = L.N[2];
        const double l = lam[0][i] + lam[1][j];
        const long long plane = (long long)L.nxc * L.N[1];
        long long c = L.cplx_index(i, j, 0);
        double beta = 0.0;
        for (int k = 0; k < Nz; ++k, c += plane) {
            // compute_main_diagonal! (HomogeneousZFormulation) :172-185 — FT reciprocals, Float64 eigenvalue term, stored as FT
            FT off;
            if (k == 0) off = -rdzf[1];
            else if (k == Nz - 1) off = -rdzf[Nz - 1];
            else off = -(rdzf[k + 1] + rdzf[k]);
            const double D = (double)(FT)((double)off - (double)dzc[k] * l);
            double t = 0.0;
            if (k > 0) {
                const double a = (double)rdzf[k];            // lower_diagonal[k-1] = upper_diagonal[k-1] = 1/Δzᶠ[k]  :198-202
                t = a / beta;
                beta = D - a * t;
            } else {
                beta = D;
            }
            double r = (beta > eps10 || beta < -eps10) ? 1.0 / beta : 0.0;
            if (zero_col && i == 0 && j == 0 && k == Nz - 1) r = 0.0;
            R[c] = (FT)r;
            T[c] = (FT)t;
        }
    }
};

template <class FT>
struct TridiagSolveKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 64;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    Cplx<FT>* spec;
    const FT* R;
    const FT* T;
    const FT* rdzf;
    const Cd* tw[2];           // ω_k = exp(-iπk/2N) for Bounded x / y (nullptr otherwise)
    int nrep[2];
    double norm;               // 1/(Nx·Ny): cuFFT's inverse is un-normalised; the DCT-III's 1/2N is the ½ of the pre-twiddle times 1/N

    // forward post-twiddle of the 2×2 orbit along every Bounded horizontal dimension (see PoissonMidKernel)
    OC_HD void twiddle_fwd(Cd v[2][2], const int idx[2][2]) const {
        for (int d = 0; d < 2; ++d) {
            if (!L.bounded[d]) continue;
            const Cd w0 = tw[d][idx[d][0]], w1 = tw[d][idx[d][1]];
            for (int q = 0; q < 2; ++q) {
                Cd* e0 = d == 0 ? &v[q][0] : &v[0][q];
                Cd* e1 = d == 0 ? &v[q][1] : &v[1][q];
                const Cd a0 = *e0, a1 = *e1;
                *e0 = cadd(cmul(w0, a0), cmul(cconj(w0), a1));
                *e1 = cadd(cmul(w1, a1), cmul(cconj(w1), a0));
            }
        }
    }
    OC_HD void twiddle_inv(Cd v[2][2], const int idx[2][2]) const {
        for (int d = 0; d < 2; ++d) {
            if (!L.bounded[d]) continue;
            const Cd w0 = tw[d][idx[d][0]], w1 = tw[d][idx[d][1]];
            const bool z0 = idx[d][0] == 0, z1 = idx[d][1] == 0;
            for (int q = 0; q < 2; ++q) {
                Cd* e0 = d == 0 ? &v[q][0] : &v[0][q];
                Cd* e1 = d == 0 ? &v[q][1] : &v[1][q];
                const Cd a0 = *e0, a1 = *e1;
                const Cd r0 = z0 ? Cd{0.0, 0.0} : a1, r1 = z1 ? Cd{0.0, 0.0} : a0;
                const Cd t0{a0.x + r0.y, a0.y - r0.x}, t1{a1.x + r1.y, a1.y - r1.x};
                const Cd h0 = cmul(cconj(w0), t0), h1 = cmul(cconj(w1), t1);
                *e0 = Cd{0.5 * h0.x, 0.5 * h0.y};
                *e1 = Cd{0.5 * h1.x, 0.5 * h1.y};
            }
        }
    }

    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int rx = b.x * nt + tid, ry = b.y;
        if (rx >= nrep[0]) return;
        const int Nz = L.N[2];
        const int rep[2] = {rx, ry};
        int idx[2][2], np[2];
        for (int d = 0; d < 2; ++d) {
            idx[d][0] = rep[d];
            if (L.bounded[d]) { idx[d][1] = (L.N[d] - rep[d]) % L.N[d]; np[d] = 2; }
            else { idx[d][1] = rep[d]; np[d] = 1; }
        }
        const long long plane = (long long)L.nxc * L.N[1];
        long long col[2][2];
        for (int bb = 0; bb < 2; ++bb)
            for (int aa = 0; aa < 2; ++aa) col[bb][aa] = L.cplx_index(idx[0][aa], idx[1][bb], 0);
        const bool zero_mode = idx[0][0] == 0 && idx[1][0] == 0;      // then every orbit member is the (0, 0) column
        Cd prev[2][2];
        for (int bb = 0; bb < 2; ++bb) for (int aa = 0; aa < 2; ++aa) prev[bb][aa] = Cd{0.0, 0.0};
        // ---- forward elimination (twiddled right-hand side in, partially eliminated ϕ out, in place)
        for (int k = 0; k < Nz; ++k) {
            Cd v[2][2];
            for (int bb = 0; bb < 2; ++bb)
                for (int aa = 0; aa < 2; ++aa) {
                    if (aa < np[0] && bb < np[1]) { const Cplx<FT> e = spec[col[bb][aa] + plane * k]; v[bb][aa] = Cd{(double)e.x, (double)e.y}; }
                    else v[bb][aa] = Cd{0.0, 0.0};
                }
            twiddle_fwd(v, idx);
            const double a = k > 0 ? (double)rdzf[k] : 0.0;
            for (int bb = 0; bb < np[1]; ++bb)
                for (int aa = 0; aa < np[0]; ++aa) {
                    const double r = (double)R[col[bb][aa] + plane * k];
                    Cd e{(v[bb][aa].x - a * prev[bb][aa].x) * r, (v[bb][aa].y - a * prev[bb][aa].y) * r};
                    prev[bb][aa] = e;
                    spec[col[bb][aa] + plane * k] = Cplx<FT>{(FT)e.x, (FT)e.y};
                }
        }
        // ---- back substitution; the finished level is scaled, pre-twiddled for the inverse transforms and stored.
        // The singular column first stores the raw solution and its sum over k, then removes the mean in a second sweep.
        Cd sum{0.0, 0.0};
        for (int k = Nz - 1; k >= 0; --k) {
            Cd v[2][2];
            for (int bb = 0; bb < 2; ++bb) for (int aa = 0; aa < 2; ++aa) v[bb][aa] = Cd{0.0, 0.0};
            for (int bb = 0; bb < np[1]; ++bb)
                for (int aa = 0; aa < np[0]; ++aa) {
                    Cd e;
                    if (k == Nz - 1) e = prev[bb][aa];
                    else {
                        const Cplx<FT> s = spec[col[bb][aa] + plane * k];
                        const double t = (double)T[col[bb][aa] + plane * (k + 1)];
                        e = Cd{(double)s.x - t * prev[bb][aa].x, (double)s.y - t * prev[bb][aa].y};
                    }
                    prev[bb][aa] = e;
                    v[bb][aa] = e;
                }
            if (zero_mode) {
                sum = cadd(sum, v[0][0]);
                for (int bb = 0; bb < np[1]; ++bb)
                    for (int aa = 0; aa < np[0]; ++aa) spec[col[bb][aa] + plane * k] = Cplx<FT>{(FT)v[bb][aa].x, (FT)v[bb][aa].y};
                continue;
            }
            for (int bb = 0; bb < 2; ++bb) for (int aa = 0; aa < 2; ++aa) { v[bb][aa].x *= norm; v[bb][aa].y *= norm; }
            twiddle_inv(v, idx);
            for (int bb = 0; bb < np[1]; ++bb)
                for (int aa = 0; aa < np[0]; ++aa) spec[col[bb][aa] + plane * k] = Cplx<FT>{(FT)v[bb][aa].x, (FT)v[bb][aa].y};
        }
        if (!zero_mode) return;
        const Cd mean{sum.x / Nz, sum.y / Nz};
        for (int k = 0; k < Nz; ++k) {
            const Cplx<FT> s = spec[col[0][0] + plane * k];
            Cd v[2][2];
            for (int bb = 0; bb < 2; ++bb)
                for (int aa = 0; aa < 2; ++aa) v[bb][aa] = Cd{((double)s.x - mean.x) * norm, ((double)s.y - mean.y) * norm};
            twiddle_inv(v, idx);
            spec[col[0][0] + plane * k] = Cplx<FT>{(FT)v[0][0].x, (FT)v[0][0].y};
        }
    }
};

// The common LES topology (Periodic, Periodic, Bounded-stretched): no twiddles, and — the coefficients being real — the real and
// imaginary parts of every column are two independent real systems.  One thread per (column, part): consecutive threads touch
// consecutive reals (fully coalesced 8-byte accesses), twice the parallelism of a thread per complex column, and the loads of
// the next PF levels are issued ahead of the serial recurrence.  Same arithmetic, in the same order, as TridiagSolveKernel.
template <class FT>
struct TridiagSolvePPKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int PF = 8;
    SpectralLayout L;
    FT* spec;                  // the spectral buffer viewed as reals: (re, im) interleaved
    const FT* R;
    const FT* T;
    const FT* rdzf;
    double norm;
    int zero_col;       // see TridiagSetupKernel
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const long long plane = (long long)L.nxc * L.N[1];
        const long long n2 = 2 * plane;
        const long long r = (long long)b.x * nt + tid;
        if (r >= n2) return;
        const int Nz = L.N[2];
        FT* p = spec + r;
        const FT* Rp = R + (r >> 1);
        const FT* Tp = T + (r >> 1);
        const bool zero_mode = zero_col && (r >> 1) == 0;
        FT vb[PF], cb[PF];
        // ---- forward elimination
        for (int q = 0; q < PF; ++q)
            if (q < Nz) { vb[q] = p[n2 * q]; cb[q] = Rp[plane * q]; }
        double prev = 0.0;
        for (int k0 = 0; k0 < Nz; k0 += PF) {
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int k = k0 + q;
                if (k < Nz) {
                    const double v = (double)vb[q], rr = (double)cb[q];
                    if (k + PF < Nz) { vb[q] = p[n2 * (k + PF)]; cb[q] = Rp[plane * (k + PF)]; }
                    const double a = k > 0 ? (double)rdzf[k] : 0.0;
                    const double e = (v - a * prev) * rr;
                    prev = e;
                    p[n2 * k] = (FT)e;
                }
            }
        }
        // ---- back substitution (descending levels; level Nz-1 is already final)
        double sum = prev;
        p[n2 * (Nz - 1)] = (FT)(zero_mode ? prev : prev * norm);
        for (int q = 0; q < PF; ++q) {
            const int k = Nz - 2 - q;
            if (k >= 0) { vb[q] = p[n2 * k]; cb[q] = Tp[plane * (k + 1)]; }
        }
        for (int k0 = Nz - 2; k0 >= 0; k0 -= PF) {
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                const int k = k0 - q;
                if (k >= 0) {
                    const double sv = (double)vb[q], t = (double)cb[q];
                    if (k - PF >= 0) { vb[q] = p[n2 * (k - PF)]; cb[q] = Tp[plane * (k - PF + 1)]; }
                    const double e = sv - t * prev;
                    prev = e;
                    sum += e;
                    p[n2 * k] = (FT)(zero_mode ? e : e * norm);
                }
            }
        }
        if (!zero_mode) return;
        const double mean = sum / Nz;        // ϕ .-= mean(ϕ): the volume mean lives in the (0, 0) column
        for (int k = 0; k < Nz; ++k) p[n2 * k] = (FT)(((double)p[n2 * k] - mean) * norm);
    }
};

// ϕ at logical cell (i,j,k) from the transform buffer; i = -1 / N handled by the caller
template <class FT>
OC_HD FT phi_at(const SpectralLayout& L, const FT* buf, int i, int j, int k) {
    return buf[L.real_index(makhoul(i, L.N[0], L.bounded[0]), makhoul(j, L.N[1], L.bounded[1]),
                            makhoul(k, L.N[2], L.bounded[2]))];
}

// Fused projection: U -= ∇ϕ over 1:N (gradient across a wall is 0 by the no-flux halo of ϕ), pNHS = ϕ/Δt
template <class FT>
struct ProjectionKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    SpectralLayout L;
    const FT* buf;
    FT* u;
    FT* v;
    FT* w;
    FT* pNHS;
    double dt_plus;       // max(eps(FT), Δt)
    const FT* prev_row;   // slab decomposition: ϕ of the y-neighbour's last row, dense (Nx, Nz); nullptr on one GPU
    const FT* prev_col;   // pencils: ϕ of the x-neighbour's last column, dense (Ny, Nz); else nullptr
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        int o = g.idx(i, j, k);
        FT p0 = phi_at<FT>(L, buf, i, j, k);
        if (!g.flat[0]) {
            FT pm = (i > 0) ? phi_at<FT>(L, buf, i - 1, j, k)
                            : (prev_col ? prev_col[j + (long long)g.N[1] * k] : (g.wlo[0] ? p0 : phi_at<FT>(L, buf, g.N[0] - 1, j, k)));
            u[o] = u[o] - (p0 - pm) * g.rd[0];
        }
        if (!g.flat[1]) {
            FT pm = (j > 0) ? phi_at<FT>(L, buf, i, j - 1, k)
                            : (prev_row ? prev_row[i + (long long)g.N[0] * k] : (g.wlo[1] ? p0 : phi_at<FT>(L, buf, i, g.N[1] - 1, k)));
            v[o] = v[o] - (p0 - pm) * g.rd[1];
        }
        if (!g.flat[2]) {
            FT pm = (k > 0) ? phi_at<FT>(L, buf, i, j, k - 1) : (g.wlo[2] ? p0 : phi_at<FT>(L, buf, i, j, g.N[2] - 1));
            w[o] = w[o] - (p0 - pm) * g.rdz_at(true, k);
        }
        pNHS[o] = (FT)((double)p0 / dt_plus);
    }
};

// slab decomposition: the last local row of ϕ, dense (Nx, Nz), for the y-neighbour's pressure gradient at its first row
template <class FT>
struct PhiRowKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    const FT* buf;
    FT* row;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, k = b.y;
        if (i >= L.N[0]) return;
        row[i + (long long)L.N[0] * k] = phi_at<FT>(L, buf, i, L.N[1] - 1, k);
    }
};

// pencils: the last local column of ϕ, dense (Ny, Nz), for the x-neighbour's pressure gradient at its first column
template <class FT>
struct PhiColKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    SpectralLayout L;
    const FT* buf;
    FT* col;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int j = b.x * nt + tid, k = b.y;
        if (j >= L.N[1]) return;
        col[j + (long long)L.N[1] * k] = phi_at<FT>(L, buf, L.N[0] - 1, j, k);
    }
};

// Staged API: buffer -> field interior (copy_real_component!), and host output for solve!
template <class FT>
struct PoissonUnpackKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    SpectralLayout L;
    const FT* buf;
    FT* field;      // halo'd field or nullptr
    FT* dense;      // dense Nx×Ny×Nz or nullptr
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= L.N[0]) return;
        FT p0 = phi_at<FT>(L, buf, i, j, k);
        if (field) field[g.idx(i, j, k)] = p0;
        if (dense) dense[(long long)i + (long long)L.N[0] * (j + (long long)L.N[1] * k)] = p0;
    }
};

// Staged API: _make_pressure_correction! from the halo-filled pNHS field, then pNHS ./= Δt (interior only)
template <class FT>
struct GradSubKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* u;
    FT* v;
    FT* w;
    FT* p;
    double dt_plus;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        int o = g.idx(i, j, k);
        FT p0 = p[o];
        if (!g.flat[0]) u[o] = u[o] - (p0 - p[o - 1]) * g.rd[0];
        if (!g.flat[1]) v[o] = v[o] - (p0 - p[o - g.sy]) * g.rd[1];
        if (!g.flat[2]) w[o] = w[o] - (p0 - p[o - g.sz]) * g.rdz_at(true, k);
    }
};
template <class FT>
struct ScaleKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* p;
    double dt_plus;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        int o = g.idx(i, j, k);
        p[o] = (FT)((double)p[o] / dt_plus);
    }
};

}  // namespace oc
