// oc_fft.h — the multi-dimensional FFT used by the pressure solve.
//
// Product build: cuFFT (libcufft), one plan over all non-unit dimensions, in place, R2C/C2R when the
// x direction is not Bounded, C2C otherwise.  Replaces plan_transforms / cuFFT plan_fft!/plan_ifft!
// (src/Solvers/plan_transforms.jl:16-136, ext/OceananigansCUDAExt.jl:112-123).
// OC_HOSTSIM build (tests only): a naive O(N²) DFT with the same layout, so the surrounding kernels can be
// exercised on a CPU-only container.
#pragma once
#include "oc_poisson.h"

#ifndef OC_HOSTSIM
#include <cufft.h>
#endif

namespace oc {

template <class FT>
class Fft3 {
public:
    SpectralLayout L{};
    size_t buffer_bytes = 0;
    size_t work_bytes = 0;

    // zbatch: z is not transformed (FourierTridiagonalPoissonSolver on a stretched grid: plan_transforms(grid, storage,
    // planner_flag, tridiagonal_dim), fourier_tridiagonal_poisson_solver.jl:107) — batched (y, x) transforms, one per level
    // force_c2c: a full complex buffer even though x is not Bounded (pencil decompositions: x is not local, every stage is complex)
    std::string init(const int N[3], const int bounded[3], Stream stream, bool plan = true, bool zbatch = false, bool force_c2c = false) {
        for (int d = 0; d < 3; ++d) { L.N[d] = N[d]; L.bounded[d] = bounded[d]; }
        zbatch_ = zbatch;
        if (zbatch) L.bounded[2] = 0;                 // no Makhoul permutation, no twiddles along z
        L.r2c = (!bounded[0] && N[0] > 1 && !force_c2c) ? 1 : 0;
        L.nxc = L.r2c ? N[0] / 2 + 1 : N[0];
        L.nxr = 2 * L.nxc;
        buffer_bytes = sizeof(FT) * 2 * (size_t)L.nxc * N[1] * N[2];
        rank_ = 0;
        for (int d = zbatch ? 1 : 2; d >= 0; --d)
            if (N[d] > 1) dims_[rank_++] = N[d];      // slowest first
        if (!plan) { rank_ = 0; return ""; }
#ifndef OC_HOSTSIM
        if (rank_ == 0) return "";
        const bool dbl = sizeof(FT) == 8;
        cufftType fwd = L.r2c ? (dbl ? CUFFT_D2Z : CUFFT_R2C) : (dbl ? CUFFT_Z2Z : CUFFT_C2C);
        cufftType inv = L.r2c ? (dbl ? CUFFT_Z2D : CUFFT_C2R) : fwd;
        size_t wf = 0, wi = 0;
        if (cufftCreate(&fwd_) != CUFFT_SUCCESS || cufftCreate(&inv_) != CUFFT_SUCCESS) return "cufftCreate failed";
        cufftSetAutoAllocation(fwd_, 0);
        cufftSetAutoAllocation(inv_, 0);
        if (zbatch) {
            // explicit in-place layouts: rows of nxr reals / nxc complex numbers, Ny rows per level, levels contiguous
            int remb[2], cemb[2];
            if (rank_ == 2) { remb[0] = N[1]; remb[1] = L.r2c ? L.nxr : N[0]; cemb[0] = N[1]; cemb[1] = L.nxc; }
            else if (N[0] > 1) { remb[0] = L.r2c ? L.nxr : N[0]; cemb[0] = L.nxc; }
            else { remb[0] = N[1]; cemb[0] = N[1]; }
            const int rdist = (L.r2c ? L.nxr : N[0]) * N[1], cdist = L.nxc * N[1];
            if (cufftMakePlanMany(fwd_, rank_, dims_, remb, 1, rdist, cemb, 1, cdist, fwd, N[2], &wf) != CUFFT_SUCCESS)
                return "cufftMakePlanMany(forward, batched over z) failed";
            if (cufftMakePlanMany(inv_, rank_, dims_, cemb, 1, cdist, remb, 1, rdist, inv, N[2], &wi) != CUFFT_SUCCESS)
                return "cufftMakePlanMany(inverse, batched over z) failed";
        } else {
        if (cufftMakePlanMany(fwd_, rank_, dims_, nullptr, 1, 0, nullptr, 1, 0, fwd, 1, &wf) != CUFFT_SUCCESS)
            return "cufftMakePlanMany(forward) failed";
        if (L.r2c || true) {
            if (cufftMakePlanMany(inv_, rank_, dims_, nullptr, 1, 0, nullptr, 1, 0, inv, 1, &wi) != CUFFT_SUCCESS)
                return "cufftMakePlanMany(inverse) failed";
        }
        }
        work_bytes = wf > wi ? wf : wi;
        if (work_bytes) {
            if (cudaMalloc(&work_, work_bytes) != cudaSuccess) return "cudaMalloc(cuFFT work area) failed";
            cufftSetWorkArea(fwd_, work_);
            cufftSetWorkArea(inv_, work_);
        }
        cufftSetStream(fwd_, stream);
        cufftSetStream(inv_, stream);
        planned_ = true;
#endif
        return "";
    }

    ~Fft3() {
#ifndef OC_HOSTSIM
        if (planned_) { cufftDestroy(fwd_); cufftDestroy(inv_); }
        if (work_) cudaFree(work_);
#endif
    }

    // in place on `buf` (device pointer; host pointer under OC_HOSTSIM)
    std::string forward(void* buf) { return exec(buf, true); }
    std::string inverse(void* buf) { return exec(buf, false); }

private:
    int rank_ = 0;
    int dims_[3] = {1, 1, 1};
    bool zbatch_ = false;
#ifndef OC_HOSTSIM
    cufftHandle fwd_ = 0, inv_ = 0;
    void* work_ = nullptr;
    bool planned_ = false;

    std::string exec(void* buf, bool fwd) {
        if (rank_ == 0) return "";
        cufftResult r;
        if (sizeof(FT) == 8) {
            if (L.r2c) r = fwd ? cufftExecD2Z(fwd_, (cufftDoubleReal*)buf, (cufftDoubleComplex*)buf)
                               : cufftExecZ2D(inv_, (cufftDoubleComplex*)buf, (cufftDoubleReal*)buf);
            else r = cufftExecZ2Z(fwd ? fwd_ : inv_, (cufftDoubleComplex*)buf, (cufftDoubleComplex*)buf, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
        } else {
            if (L.r2c) r = fwd ? cufftExecR2C(fwd_, (cufftReal*)buf, (cufftComplex*)buf)
                               : cufftExecC2R(inv_, (cufftComplex*)buf, (cufftReal*)buf);
            else r = cufftExecC2C(fwd ? fwd_ : inv_, (cufftComplex*)buf, (cufftComplex*)buf, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
        }
        if (r != CUFFT_SUCCESS) return "cuFFT exec failed with code " + std::to_string((int)r);
        return "";
    }
#else
    // naive DFT, test-only
    std::string exec(void* bufv, bool fwd) {
        const int Nx = L.N[0], Ny = L.N[1], Nz = L.N[2];
        std::vector<Cd> full((size_t)Nx * Ny * Nz);
        FT* buf = (FT*)bufv;
        auto at = [&](int i, int j, int k) -> Cd& { return full[(size_t)i + (size_t)Nx * (j + (size_t)Ny * k)]; };
        if (fwd) {
            for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
                long long r = L.real_index(i, j, k);
                at(i, j, k) = Cd{(double)buf[r], L.r2c ? 0.0 : (double)buf[r + 1]};
            }
        } else {
            for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
                if (!L.r2c || i < L.nxc) {
                    long long c = L.cplx_index(i, j, k);
                    at(i, j, k) = Cd{(double)buf[2 * c], (double)buf[2 * c + 1]};
                } else {   // Hermitian partner
                    long long c = L.cplx_index(Nx - i, (Ny - j) % Ny, zbatch_ ? k : (Nz - k) % Nz);
                    at(i, j, k) = Cd{(double)buf[2 * c], -(double)buf[2 * c + 1]};
                }
            }
        }
        const double sgn = fwd ? -1.0 : 1.0;
        const int n[3] = {Nx, Ny, Nz};
        for (int d = 0; d < 3; ++d) {
            if (n[d] == 1 || (d == 2 && zbatch_)) continue;
            std::vector<Cd> line(n[d]), out(n[d]);
            int o1 = (d + 1) % 3, o2 = (d + 2) % 3;
            for (int a = 0; a < n[o1]; ++a) for (int bq = 0; bq < n[o2]; ++bq) {
                int ijk[3];
                ijk[o1] = a; ijk[o2] = bq;
                for (int m = 0; m < n[d]; ++m) { ijk[d] = m; line[m] = at(ijk[0], ijk[1], ijk[2]); }
                for (int q = 0; q < n[d]; ++q) {
                    Cd s{0, 0};
                    for (int m = 0; m < n[d]; ++m) {
                        double ang = sgn * 2.0 * M_PI * (double)((long long)q * m % n[d]) / n[d];
                        s = cadd(s, cmul(line[m], Cd{std::cos(ang), std::sin(ang)}));
                    }
                    out[q] = s;
                }
                for (int m = 0; m < n[d]; ++m) { ijk[d] = m; at(ijk[0], ijk[1], ijk[2]) = out[m]; }
            }
        }
        if (fwd) {
            for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < L.nxc; ++i) {
                long long c = L.cplx_index(i, j, k);
                buf[2 * c] = (FT)at(i, j, k).x; buf[2 * c + 1] = (FT)at(i, j, k).y;
            }
        } else {
            for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
                long long r = L.real_index(i, j, k);
                buf[r] = (FT)at(i, j, k).x;
                if (!L.r2c) buf[r + 1] = (FT)at(i, j, k).y;
            }
        }
        return "";
    }
#endif
};

}  // namespace oc
