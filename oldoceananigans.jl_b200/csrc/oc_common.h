// oc_common.h — geometry, coefficient tables and small helpers shared by all kernels.
//
// Internal field layout: every 3-D field of a model shares ONE geometry.  A field
// pointer `p` addresses interior element (0,0,0) (reference index (1,1,1)); element (i,j,k) is
// p[i + j*sy + k*sz].  Rows are padded so that p and every row start are 128-byte aligned; halos
// (H planes each side, +1 plane for the extra Face point of Bounded dimensions) surround the interior.
// Flat dimensions are stored as Periodic with N = 1 and H = 3 (all planes equal), which makes every
// difference zero and every interpolation the identity, exactly as src/Operators/*.jl define for Flat.
#pragma once
#include "oc_exec.h"

namespace oc {

template <class FT>
struct Geom {
    int N[3];        // interior size
    int H[3];        // internal halo size
    int bounded[3];  // 1 = Bounded topology
    // a wall on the low / high side of THIS domain: both equal `bounded` on one GPU; on a slab of a distributed Bounded dimension only
    // the first rank has the low wall and the last rank the high one (RightConnected / LeftConnected / FullyConnected local grids,
    // distributed_grids.jl:75-126) — every wall-aware kernel tests these, never `bounded`
    int wlo[3], whi[3];
    int flat[3];     // 1 = Flat (stored as periodic N=1)
    int sy, sz;      // strides in elements (sx = 1)
    FT d[3];         // Δx, Δy, Δz
    FT rd[3];        // 1/Δ           reciprocal_metric_operators.jl:7
    FT A[3];         // Ax=Δy·Δz, Ay=Δx·Δz, Az=Δx·Δy   spacings_and_areas_and_volumes.jl:309-333
    FT V, rV;        // V = Az·Δz, 1/V                  :376, reciprocal_metric_operators.jl:13
    // Vertically stretched grid (z Bounded, variably spaced; x and y stay regular): level tables, indexable with the
    // 0-based level k = -(H[2]+1) … N[2]+H[2]+1 (reference index k+1).  nullptr on regular grids, where d[2], rd[2], A[], V, rV
    // are the (constant) metrics.  dzc = Δzᵃᵃᶜ, dzf = Δzᵃᵃᶠ (grid_generation.jl:33-94), rdzc / rdzf = their reciprocals
    // (reciprocal_metric_operators.jl:7), rVc / rVf = 1/(Az·Δz) at Center / Face levels (:13).
    const FT* dzc;
    const FT* dzf;
    const FT* rdzc;
    const FT* rdzf;
    const FT* rVc;
    const FT* rVf;
    OC_HD int st(int dim) const { return dim == 0 ? 1 : (dim == 1 ? sy : sz); }
    OC_HD int idx(int i, int j, int k) const { return i + j * sy + k * sz; }
    OC_HD bool stretched() const { return dzc != nullptr; }
    // metrics at level k for a point whose z-location is Face (zf = true) or Center
    OC_HD FT dz_at(bool zf, int k) const { return dzc ? (zf ? dzf[k] : dzc[k]) : d[2]; }
    OC_HD FT rdz_at(bool zf, int k) const { return dzc ? (zf ? rdzf[k] : rdzc[k]) : rd[2]; }
    OC_HD FT rV_at(bool zf, int k) const { return dzc ? (zf ? rVf[k] : rVc[k]) : rV; }
    // Ax = Δy·Δz, Ay = Δx·Δz, Az = Δx·Δy   (spacings_and_areas_and_volumes.jl:309-333)
    OC_HD FT area_at(int dim, bool zf, int k) const {
        if (dim == 2 || !dzc) return A[dim];
        return d[dim == 0 ? 1 : 0] * (zf ? dzf[k] : dzc[k]);
    }
    OC_HD FT vol_at(bool zf, int k) const { return dzc ? A[2] * (zf ? dzf[k] : dzc[k]) : V; }
};

// Reconstruction coefficients, computed on the host exactly like the reference
// (reconstruction_coefficients.jl:49-64: rationals rounded to FT, the last one is 1 - sum(others)).
template <class FT>
struct AdvCoef {
    FT c4[4];        // Centered(order=4) in stencil order ψ[i-2..i+1]
    FT w5p[3][3];    // WENO{3} coeff_p(r)   weno_interpolants.jl:117-118
    FT w5c[3];       // C★ = 3/10, 3/5, 1/10  :81-83
    FT w3p[2][2];    // WENO{2} coeff_p(r)
    FT w3c[2];       // C★ = 2/3, 1/3         :78-79
    FT eps;          // ϵ = 1f-8 widened      :71
    // UpwindBiased(order = 5, 3): uniform_reconstruction_coefficients(FT, Val(:left / :right), buffer) re-ordered to the UPWIND-ordered
    // stencil q0 … (Left: ψ[i-B] …, Right: ψ[i+B-1] …; calc_reconstruction_stencil, reconstruction_coefficients.jl:87-89,122-152)
    FT u5l[5], u5r[5];
    FT u3l[3], u3r[3];
    // WENO(order = 7 | 9) and their Centered(6) / Centered(8) advecting-velocity schemes: one device table (HiOrderTab layout below),
    // nullptr unless the model uses them — the kernel parameter block of every other scheme stays what it was
    const FT* hi;
};

// layout of AdvCoef::hi (FT values): smoothness coefficients, coeff_p, C★ of WENO{4} and WENO{5}
// (weno_interpolants.jl:81-90,175-185, :117-118), Centered(6) and Centered(8) in stencil order
struct HiOrderTab {
    enum { S7 = 0, P7 = S7 + 4 * 10, C7 = P7 + 4 * 4, S9 = C7 + 4, P9 = S9 + 5 * 15, C9 = P9 + 5 * 5, CEN6 = C9 + 5, CEN8 = CEN6 + 6, SIZE = CEN8 + 8 };
};

// advection scheme codes (oc_advection in include/oceananigans_b200.h)
enum { ADV_CENTERED2 = 0, ADV_WENO5 = 1, ADV_CENTERED4 = 2, ADV_UPWIND3 = 3, ADV_UPWIND5 = 4, ADV_WENO3 = 5, ADV_UPWIND1 = 6, ADV_NONE = 7,
       ADV_MIXED = 8 /* FluxFormAdvection(x, y, z): the scheme of each flux direction is a run-time code (TendencyArgs::adv_dir) */,
       ADV_WENO7 = 9, ADV_WENO9 = 10 };
OC_HD constexpr bool adv_is_centered(int adv) { return adv == ADV_CENTERED2 || adv == ADV_CENTERED4; }

template <class FT>
OC_HD FT oc_abs(FT x) { return x < FT(0) ? -x : x; }

template <class FT>
OC_HD FT oc_max(FT a, FT b) { return a > b ? a : b; }

// IEEE square root (sqrt.rn on the device; the kernels are compiled without fast-math)
OC_HD double oc_sqrt_(double x) {
#ifdef OC_HOSTSIM
    return std::sqrt(x);
#else
    return sqrt(x);
#endif
}
OC_HD float oc_sqrt_(float x) {
#ifdef OC_HOSTSIM
    return std::sqrt(x);
#else
    return sqrtf(x);
#endif
}
template <class FT>
OC_HD FT oc_sqrt(FT x) { return oc_sqrt_(x); }

OC_HD double oc_fma(double a, double b, double c) {
#ifdef OC_HOSTSIM
    return std::fma(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// newton_div(Float32, a, b)   src/Utils/newton_div.jl:8-23
OC_HD double newton_div(double a, double b) {
    float bl = (float)b;
#ifdef OC_HOSTSIM
    float inv = 1.0f / bl;
#else
#ifdef __CUDA_ARCH__
    float inv = __frcp_rn(bl);
#else
    float inv = 1.0f / bl;
#endif
#endif
    double invd = (double)inv;
    double x = a * invd;
    return oc_fma(oc_fma(x, -b, a), invd, x);
}
OC_HD float newton_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return a * __frcp_rn(b);
#else
    return a * (1.0f / b);
#endif
}

}  // namespace oc
