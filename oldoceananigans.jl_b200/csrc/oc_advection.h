// oc_advection.h — point reconstructions for flux-form advection: Centered(order=2/4), WENO(order=5)
// with its WENO(order=3) -> UpwindBiased(order=1) boundary chain.
//
// Replaces (paths under /root/reference/src/Advection): centered_reconstruction.jl:47-55,
// upwind_biased_reconstruction.jl:57-86, weno_interpolants.jl:71-83,117-137,169-174,204-216,261-266,
// 290-338,409-437,500, reconstruction_coefficients.jl:122-152, topologically_conditional_interpolation.jl
// :46-52,99-120.  All functions take a pointer `p` to ψ at the FACE index (ψ[i] in the reference's
// notation; ψ[i+n] = p[n*s]) and work for global or shared memory alike.
#pragma once
#include "oc_common.h"

namespace oc {

// ---- symmetric (advecting-velocity) interpolation: Σ c_n · (a·q[i+n]) --------------------------------
// Centered(order=2):  0.5 q[i-1] + 0.5 q[i]      (centered_reconstruction.jl:47-49, coefficients (1/2, 1/2))
template <class FT>
OC_HD FT sym2(const FT* p, int s, FT a) {
    return FT(0.5) * (a * p[-s]) + FT(0.5) * (a * p[0]);
}
// Centered(order=4):  c0 q[i-2] + c1 q[i-1] + c2 q[i] + c3 q[i+1], summed left to right (@muladd)
template <class FT>
OC_HD FT sym4(const AdvCoef<FT>& C, const FT* p, int s, FT a) {
    FT r = C.c4[0] * (a * p[-2 * s]);
    r = r + C.c4[1] * (a * p[-s]);
    r = r + C.c4[2] * (a * p[0]);
    r = r + C.c4[3] * (a * p[s]);
    return r;
}

// The same along z on a vertically stretched grid: the area belongs to the advecting velocity's own point, a_n = h·Δzᶜ[k+n]
// (Ax_qᶠᶜᶜ / Ay_qᶜᶠᶜ inside ℑzᵃᵃᶠ, upwind_biased_advective_fluxes.jl:79-91); dz points at Δzᶜ of level k.
template <class FT>
OC_HD FT sym2z(const FT* p, int s, FT h, const FT* dz) {
    return FT(0.5) * ((h * dz[-1]) * p[-s]) + FT(0.5) * ((h * dz[0]) * p[0]);
}
template <class FT>
OC_HD FT sym4z(const AdvCoef<FT>& C, const FT* p, int s, FT h, const FT* dz) {
    FT r = C.c4[0] * ((h * dz[-2]) * p[-2 * s]);
    r = r + C.c4[1] * ((h * dz[-1]) * p[-s]);
    r = r + C.c4[2] * ((h * dz[0]) * p[0]);
    r = r + C.c4[3] * ((h * dz[1]) * p[s]);
    return r;
}

// ---- WENO building blocks -----------------------------------------------------------------------------
// β for WENO{3}: ψ1(C1ψ1 + C2ψ2 + C3ψ3) + ψ2(C4ψ2 + C5ψ3) + ψ3ψ3C6     weno_interpolants.jl:204-216,261
template <class FT>
OC_HD FT beta3(FT a, FT b, FT c, FT C1, FT C2, FT C3, FT C4, FT C5, FT C6) {
    return a * (C1 * a + C2 * b + C3 * c) + b * (C4 * b + C5 * c) + c * c * C6;
}

// WENO(order=5) biased reconstruction at the face.  q0..q4 is the 5-point upwind-ordered stencil:
//   LeftBias : (ψ[i-3], ψ[i-2], ψ[i-1], ψ[i],   ψ[i+1])
//   RightBias: (ψ[i+2], ψ[i+1], ψ[i],   ψ[i-1], ψ[i-2])
// so that S0 = (q2,q3,q4), S1 = (q1,q2,q3), S2 = (q0,q1,q2) for both biases (weno_interpolants.jl:435-437).
template <class FT>
OC_HD FT weno5_value(const AdvCoef<FT>& C, FT q0, FT q1, FT q2, FT q3, FT q4) {
    // smoothness_coefficients :172-174
    FT b0 = beta3<FT>(q2, q3, q4, FT(10), FT(-31), FT(11), FT(25), FT(-19), FT(4));
    FT b1 = beta3<FT>(q1, q2, q3, FT(4), FT(-13), FT(5), FT(13), FT(-13), FT(4));
    FT b2 = beta3<FT>(q0, q1, q2, FT(4), FT(-19), FT(11), FT(25), FT(-31), FT(10));
    FT tau = oc_abs<FT>(b0 - b2);                                   // :309
    FT r0 = newton_div(tau, b0 + C.eps);                            // :293
    FT r1 = newton_div(tau, b1 + C.eps);
    FT r2 = newton_div(tau, b2 + C.eps);
    FT a0 = C.w5c[0] * (FT(1) + r0 * r0);
    FT a1 = C.w5c[1] * (FT(1) + r1 * r1);
    FT a2 = C.w5c[2] * (FT(1) + r2 * r2);
    FT rs = FT(1) / (a0 + a1 + a2);                                 // :336
    FT p0 = C.w5p[0][0] * q2 + C.w5p[0][1] * q3 + C.w5p[0][2] * q4; // biased_p :136-137
    FT p1 = C.w5p[1][0] * q1 + C.w5p[1][1] * q2 + C.w5p[1][2] * q3;
    FT p2 = C.w5p[2][0] * q0 + C.w5p[2][1] * q1 + C.w5p[2][2] * q2;
    return (a0 * rs) * p0 + (a1 * rs) * p1 + (a2 * rs) * p2;        // :500
}

// WENO(order=3): q0..q2 upwind ordered: Left (ψ[i-2], ψ[i-1], ψ[i]); Right (ψ[i+1], ψ[i], ψ[i-1]);
// S0 = (q1,q2), S1 = (q0,q1)   (:432-433); β = ψ1(ψ1 - 2ψ2) + ψ2ψ2   (:169-170)
template <class FT>
OC_HD FT weno3_value(const AdvCoef<FT>& C, FT q0, FT q1, FT q2) {
    FT b0 = q1 * (FT(1) * q1 + FT(-2) * q2) + q2 * q2 * FT(1);
    FT b1 = q0 * (FT(1) * q0 + FT(-2) * q1) + q1 * q1 * FT(1);
    FT tau = oc_abs<FT>(b0 - b1);                                   // :308
    FT r0 = newton_div(tau, b0 + C.eps);
    FT r1 = newton_div(tau, b1 + C.eps);
    FT a0 = C.w3c[0] * (FT(1) + r0 * r0);
    FT a1 = C.w3c[1] * (FT(1) + r1 * r1);
    FT rs = FT(1) / (a0 + a1);
    FT p0 = C.w3p[0][0] * q1 + C.w3p[0][1] * q2;
    FT p1 = C.w3p[1][0] * q0 + C.w3p[1][1] * q1;
    return (a0 * rs) * p0 + (a1 * rs) * p1;
}

// Order selector near walls of a Bounded dimension, in terms of the 0-based FACE index f of the stencil
// (topologically_conditional_interpolation.jl:46-52 with required_halo_size = 3 and 2):
//   face-type  : high order iff 3 <= f <= N-3 ; mid order iff 2 <= f <= N-2
//   centre-type: (evaluated at face f = c+1)  high iff 3 <= f <= N-2 ; mid iff 2 <= f <= N-1
struct OrderWindow {
    int lo_hi = 0, hi_hi = 0, lo_mid = 0, hi_mid = 0;
};
// wlo / whi: a wall on the low / high side (a connected side of a slab never lowers the order: RightConnected / LeftConnected,
// topologically_conditional_interpolation.jl:55-69)
OC_HD OrderWindow order_window(bool wlo, bool whi, bool centre_type, int N) {
    OrderWindow w;
    w.lo_hi = w.lo_mid = -(1 << 30);
    w.hi_hi = w.hi_mid = (1 << 30);
    if (wlo) { w.lo_hi = 3; w.lo_mid = 2; }
    if (whi) {
        if (centre_type) { w.hi_hi = N - 2; w.hi_mid = N - 1; }
        else { w.hi_hi = N - 3; w.hi_mid = N - 2; }
    }
    return w;
}

// _biased_interpolate for the WENO(order=5) scheme at face index f (pointer p at ψ[f], stride s).
template <class FT>
OC_HD FT weno5_biased(const AdvCoef<FT>& C, const FT* p, int s, bool left, int f, const OrderWindow& w) {
    if (f >= w.lo_hi && f <= w.hi_hi) {
        int o0 = left ? -3 * s : 2 * s, ds = left ? s : -s;
        return weno5_value<FT>(C, p[o0], p[o0 + ds], p[o0 + 2 * ds], p[o0 + 3 * ds], p[o0 + 4 * ds]);
    } else if (f >= w.lo_mid && f <= w.hi_mid) {
        int o0 = left ? -2 * s : s, ds = left ? s : -s;
        return weno3_value<FT>(C, p[o0], p[o0 + ds], p[o0 + 2 * ds]);
    }
    return left ? p[-s] : p[0];                                      // UpwindBiased(order=1)
}

// _symmetric_interpolate for the WENO(order=5) scheme (Centered(4) -> Centered(2) near walls) of a·q
template <class FT>
OC_HD FT weno5_symmetric(const AdvCoef<FT>& C, const FT* p, int s, FT a, int f, const OrderWindow& w) {
    if (f >= w.lo_hi && f <= w.hi_hi) return sym4<FT>(C, p, s, a);
    return sym2<FT>(p, s, a);
}

template <class FT>
OC_HD FT weno5_symmetric_z(const AdvCoef<FT>& C, const FT* p, int s, FT h, const FT* dz, int f, const OrderWindow& w) {
    if (f >= w.lo_hi && f <= w.hi_hi) return sym4z<FT>(C, p, s, h, dz);
    return sym2z<FT>(p, s, h, dz);
}


// ---- WENO(order = 7 | 9): buffer B = 4 | 5 ------------------------------------------------------
// The reference's generic machinery (weno_interpolants.jl): sub-stencil r = 0 … B-1 of the upwind-ordered stencil q0 … q(2B-2) is
// S_r = (q[B-1-r] … q[2B-2-r]) (:409-437); β_r by the metaprogrammed quadratic form (:204-266) with the tabulated coefficients
// (:175-185); τ = |β₀ + 3β₁ − 3β₂ − β₃| (B = 4), |β₀ + 2β₁ − 6β₂ + 2β₃ + β₄| (B = 5) (:303-307); α_r = C★_r (1 + (τ / (β_r + ϵ))²) with
// newton_div (:290-297); Σ (α_r / Σα) p_r with p_r = Σ coeff_p(r) · S_r (:136-137, :332-338, :500).
template <int B, class FT>
OC_HD FT weno_hi_value(const AdvCoef<FT>& C, const FT* q) {
    constexpr int NS = B * (B + 1) / 2;
    const FT* S = C.hi + (B == 4 ? HiOrderTab::S7 : HiOrderTab::S9);
    const FT* P = C.hi + (B == 4 ? HiOrderTab::P7 : HiOrderTab::P9);
    const FT* CS = C.hi + (B == 4 ? HiOrderTab::C7 : HiOrderTab::C9);
    FT beta[B], pr[B];
#ifndef OC_HOSTSIM
#pragma unroll
#endif
    for (int r = 0; r < B; ++r) {
        const FT* psi = q + (B - 1 - r);
        const FT* c = S + r * NS;
        FT b = FT(0);
        int ci = 0;
        for (int s0 = 0; s0 < B - 1; ++s0) {
            FT inner = c[ci] * psi[s0];
            for (int i = s0 + 1; i < B; ++i) inner = inner + c[ci + i - s0] * psi[i];
            ci += B - s0;
            const FT term = psi[s0] * inner;
            b = s0 == 0 ? term : b + term;
        }
        beta[r] = b + psi[B - 1] * psi[B - 1] * c[ci];
        FT p = P[r * B] * psi[0];
        for (int j = 1; j < B; ++j) p = p + P[r * B + j] * psi[j];
        pr[r] = p;
    }
    FT tau;
    if (B == 4) tau = oc_abs<FT>(beta[0] + FT(3) * beta[1] - FT(3) * beta[2] - beta[3]);
    else tau = oc_abs<FT>(beta[0] + FT(2) * beta[1] - FT(6) * beta[2] + FT(2) * beta[3] + beta[B - 1]);
    FT alpha[B];
    FT sum = FT(0);
    for (int r = 0; r < B; ++r) {
        const FT t = newton_div(tau, beta[r] + C.eps);
        alpha[r] = CS[r] * (FT(1) + t * t);
        sum = r == 0 ? alpha[0] : sum + alpha[r];
    }
    const FT rs = FT(1) / sum;
    FT out = (alpha[0] * rs) * pr[0];
    for (int r = 1; r < B; ++r) out = out + (alpha[r] * rs) * pr[r];
    return out;
}

// order windows of buffer b in terms of the 0-based face index f: face-type b <= f <= N - b, centre-type (evaluated at face c + 1)
// b <= f <= N + 1 - b (topologically_conditional_interpolation.jl:46-52 with required_halo_size = b)
// — from the OrderWindow of buffer 3 (hi_hi = N - 3 face-type, N - 2 centre-type; lo_hi < 0: no wall on the low side; hi_hi = 2^30: none
// on the high side — each side on its own: the outer slabs of a distributed Bounded dimension have one wall only)
OC_HD bool in_order_window(int b, int f, const OrderWindow& w) {
    return (w.lo_hi < 0 || f >= b) && f <= w.hi_hi + 3 - b;
}

// _biased_interpolate of WENO(2B-1), B = 4 | 5: the scheme inside its window, else its buffer_scheme (WENO(2B-3) …) — the chain ends in
// weno5_biased (WENO(5) -> WENO(3) -> UpwindBiased(1))
template <int B, class FT>
OC_HD FT weno_hi_biased(const AdvCoef<FT>& C, const FT* p, int s, bool left, int f, const OrderWindow& w) {
    if (in_order_window(B, f, w)) {
        FT q[2 * B - 1];
        const int o0 = left ? -B * s : (B - 1) * s, ds = left ? s : -s;
        for (int n = 0; n < 2 * B - 1; ++n) q[n] = p[o0 + n * ds];
        return weno_hi_value<B, FT>(C, q);
    }
    if constexpr (B == 5) return weno_hi_biased<4, FT>(C, p, s, left, f, w);
    else return weno5_biased<FT>(C, p, s, left, f, w);
}

// Centered(6) / Centered(8) of a·q in stencil order ψ[i-B'] … ψ[i+B'-1], summed left to right; `dz` != nullptr: the stretched-z form with
// the area inside the interpolation, a_n = a·Δzᶜ[k+n] (sym4z)
template <int BC, class FT>
OC_HD FT sym_hi(const AdvCoef<FT>& C, const FT* p, int s, FT a, const FT* dz) {
    const FT* c = C.hi + (BC == 3 ? HiOrderTab::CEN6 : HiOrderTab::CEN8);
    FT r = c[0] * ((dz ? a * dz[-BC] : a) * p[-BC * s]);
    for (int n = 1; n < 2 * BC; ++n) r = r + c[n] * ((dz ? a * dz[n - BC] : a) * p[(n - BC) * s]);
    return r;
}
// _symmetric_interpolate of WENO(2B-1): advecting_velocity_scheme = Centered(2B-2) inside the window of B, else the buffer scheme's
template <int B, class FT>
OC_HD FT weno_hi_symmetric(const AdvCoef<FT>& C, const FT* p, int s, FT a, const FT* dz, int f, const OrderWindow& w) {
    if (in_order_window(B, f, w)) return sym_hi<B - 1, FT>(C, p, s, a, dz);
    if constexpr (B == 5) return weno_hi_symmetric<4, FT>(C, p, s, a, dz, f, w);
    else return dz ? weno5_symmetric_z<FT>(C, p, s, a, dz, f, w) : weno5_symmetric<FT>(C, p, s, a, f, w);
}

// ---- the other schemes of the reference's family up to order 5 -----------------------------------
// _biased_interpolate with the boundary chain of topologically_conditional_interpolation.jl:46-52,99-120:
//   UpwindBiased(5) -> UpwindBiased(3) -> UpwindBiased(1);  UpwindBiased(3) -> UpwindBiased(1);  WENO(3) -> UpwindBiased(1).
// The hi / mid windows of OrderWindow are the required_halo_size = 3 / 2 conditions.
template <int ADV, class FT>
OC_HD FT biased_any(const AdvCoef<FT>& C, const FT* p, int s, bool left, int f, const OrderWindow& w) {
    if (ADV == ADV_WENO5) return weno5_biased<FT>(C, p, s, left, f, w);
    if (ADV == ADV_WENO7) return weno_hi_biased<4, FT>(C, p, s, left, f, w);
    if (ADV == ADV_WENO9) return weno_hi_biased<5, FT>(C, p, s, left, f, w);
    if (ADV == ADV_UPWIND5 && f >= w.lo_hi && f <= w.hi_hi) {
        // upwind-ordered stencil: Left (ψ[i-3] … ψ[i+1]), Right (ψ[i+2] … ψ[i-2])
        const int o0 = left ? -3 * s : 2 * s, ds = left ? s : -s;
        const FT* c = left ? C.u5l : C.u5r;
        FT r = c[0] * p[o0];
        r = r + c[1] * p[o0 + ds];
        r = r + c[2] * p[o0 + 2 * ds];
        r = r + c[3] * p[o0 + 3 * ds];
        r = r + c[4] * p[o0 + 4 * ds];
        return r;
    }
    if ((ADV == ADV_UPWIND5 || ADV == ADV_UPWIND3) && f >= w.lo_mid && f <= w.hi_mid) {
        const int o0 = left ? -2 * s : s, ds = left ? s : -s;
        const FT* c = left ? C.u3l : C.u3r;
        FT r = c[0] * p[o0];
        r = r + c[1] * p[o0 + ds];
        r = r + c[2] * p[o0 + 2 * ds];
        return r;
    }
    if (ADV == ADV_WENO3 && f >= w.lo_mid && f <= w.hi_mid) {
        const int o0 = left ? -2 * s : s, ds = left ? s : -s;
        return weno3_value<FT>(C, p[o0], p[o0 + ds], p[o0 + 2 * ds]);
    }
    return left ? p[-s] : p[0];                                      // UpwindBiased(order=1)
}

// _symmetric_interpolate of a·q: Centered(4) inside the scheme's window, Centered(2) otherwise.
//   WENO(5), UpwindBiased(5): advecting_velocity_scheme = Centered(4) in the hi window; Centered(4) itself: in the mid window
//   (its required_halo_size is 2); every other scheme: Centered(2).
template <int ADV, class FT>
OC_HD FT symmetric_any(const AdvCoef<FT>& C, const FT* p, int s, FT a, int f, const OrderWindow& w) {
    if (ADV == ADV_WENO7) return weno_hi_symmetric<4, FT>(C, p, s, a, nullptr, f, w);
    if (ADV == ADV_WENO9) return weno_hi_symmetric<5, FT>(C, p, s, a, nullptr, f, w);
    if ((ADV == ADV_WENO5 || ADV == ADV_UPWIND5) && f >= w.lo_hi && f <= w.hi_hi) return sym4<FT>(C, p, s, a);
    if (ADV == ADV_CENTERED4 && f >= w.lo_mid && f <= w.hi_mid) return sym4<FT>(C, p, s, a);
    return sym2<FT>(p, s, a);
}
template <int ADV, class FT>
OC_HD FT symmetric_any_z(const AdvCoef<FT>& C, const FT* p, int s, FT h, const FT* dz, int f, const OrderWindow& w) {
    if (ADV == ADV_WENO7) return weno_hi_symmetric<4, FT>(C, p, s, h, dz, f, w);
    if (ADV == ADV_WENO9) return weno_hi_symmetric<5, FT>(C, p, s, h, dz, f, w);
    if ((ADV == ADV_WENO5 || ADV == ADV_UPWIND5) && f >= w.lo_hi && f <= w.hi_hi) return sym4z<FT>(C, p, s, h, dz);
    return sym2z<FT>(p, s, h, dz);
}

}  // namespace oc
