// oc_march.h — z-marching, TMA-staged fused tendency (+ RK3/AB2 substep) kernel for one prognostic field.
//
// Replaces the same reference functions as oc_tendency.h (compute_Gu!/Gv!/Gw!/Gc!, compute_flux_bc_tendencies!,
// rk3_substep_field!, ab2_step_field!, _cache_field_tendencies!; see the citations there) for grids without
// Flat dimensions.  oc_tendency.h stays the general kernel (Flat dimensions, tiny grids).
//
// Design.  A CTA owns a TX×TY column of cells and marches through a chunk of z-levels.
//  * Every stencil operand comes from shared memory.  Per field the CTA keeps a RING of x–y planes (with the
//    halo the stencils need) that one elected thread fills with TMA (cp.async.bulk.tensor.3d → mbarrier
//    complete_tx), PF = 3 planes ahead of the level being computed: the advected field needs levels
//    k-2 … k+3 (WENO-5 in z), advecting velocities 1 to 4 levels.  HBM sees each plane once; halo overlap between
//    neighbouring CTAs is served by L2.
//  * Per level: phase 0 evaluates every face flux of the level once — (TX+1)·TY x-faces, TX·(TY+1) y-faces and the
//    TX·TY upper z-faces (the lower z-face flux is the previous level's upper one and is carried in shared memory) —
//    __syncthreads — phase 1 forms the flux divergence, adds Coriolis / hydrostatic pressure gradient / flux-BC
//    terms, writes Gⁿ and the substepped field (coalesced stores) and issues the TMA loads three levels ahead.
//    One __syncthreads per level; flux buffers are double (x, y) / triple (z) buffered.
#pragma once
#include "oc_tendency.h"

#ifndef OC_HOSTSIM
#include <cuda.h>
#endif

// prefetch distance of the velocity rings (MarchSpec::PFV) for tile height TY, kernel kind and ring-0 distance PF
#ifndef OC_MARCH_PFV
#define OC_MARCH_PFV(TY, KIND, PF) ((TY) == 16 ? (PF) + 1 : (PF))
#endif

namespace oc {

// ---------------------------------------------------------------------------------------------------------
// TMA source descriptor and primitives.  Product build: CUtensorMap + PTX.  OC_HOSTSIM (tests only): plain copies.
// ---------------------------------------------------------------------------------------------------------
#ifndef OC_HOSTSIM
template <class FT>
struct alignas(64) TileSrc {
    CUtensorMap map;
};

OC_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
OC_DEV void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
OC_DEV void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
OC_DEV void mbar_expect(uint64_t* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Potentially-blocking wait.  A lost TMA transaction or a protocol bug would otherwise hang the GPU until the watchdog: after
// ~2^26 failed polls (seconds) the kernel traps, which the host sees as a launch failure instead of a hang.
OC_DEV void mbar_wait(uint64_t* bar, int parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        ".reg .u32 cnt;\n\t"
        "mov.u32 cnt, 0;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "add.u32 cnt, cnt, 1;\n\t"
        "setp.gt.u32 P1, cnt, 0x4000000;\n\t"
        "@P1 trap;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
OC_DEV void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
OC_DEV void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
template <class FT>
OC_DEV void tile_issue(void* dst, const TileSrc<FT>* src, int c0, int c1, int c2, uint64_t* bar, int, int) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)(uintptr_t)&src->map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
#else
template <class FT>
struct TileSrc {
    const FT* base;
    int dim[3];
    long long stride[3];
};
inline void mbar_init(uint64_t*, int) {}
inline void mbar_fence_init() {}
inline void mbar_expect(uint64_t*, int) {}
inline void mbar_wait(uint64_t*, int) {}
inline void proxy_fence_async() {}
inline void mbar_arrive(uint64_t*) {}
template <class FT>
inline void tile_issue(void* dst, const TileSrc<FT>* src, int c0, int c1, int c2, uint64_t*, int bx, int by) {
    FT* d = (FT*)dst;
    for (int j = 0; j < by; ++j)
        for (int i = 0; i < bx; ++i) {
            int x = c0 + i, y = c1 + j, z = c2;
            bool in = x >= 0 && x < src->dim[0] && y >= 0 && y < src->dim[1] && z >= 0 && z < src->dim[2];
            d[j * bx + i] = in ? src->base[x * src->stride[0] + y * src->stride[1] + z * src->stride[2]] : FT(0);   // OOB fill = 0, like TMA
        }
}
#endif

// ---------------------------------------------------------------------------------------------------------
// rings
// ---------------------------------------------------------------------------------------------------------
template <int XO_, int BX_, int YO_, int BY_, int LO_, int HI_, int D_>
struct RingSpec {
    static constexpr int XO = XO_, BX = BX_, YO = YO_, BY = BY_, LO = LO_, HI = HI_, D = D_;
    static constexpr int LIVE = HI_ - LO_ + 1;
};

template <class FT, class RS>
struct Ring {
    static constexpr int BOX_BYTES = RS::BX * RS::BY * (int)sizeof(FT);
    static constexpr int SLOT_BYTES = ((BOX_BYTES + 127) / 128) * 128;
    static constexpr int SLOT = SLOT_BYTES / (int)sizeof(FT);
    static constexpr int BYTES = SLOT_BYTES * RS::D;
    static constexpr bool POW2 = (RS::D & (RS::D - 1)) == 0;
    static constexpr int STRIDE_Y = RS::BX;
    FT* s;
    int k;        // the level of the current iteration
    int sk;       // its ring slot (carried and advanced by the marching loop: no per-access modulo)
    OC_HD static int slot_of(int lev) { return POW2 ? ((lev + 64 * RS::D) & (RS::D - 1)) : ((lev + 64 * RS::D) % RS::D); }
    // slot of level lev = k + c, |c| < D (c folds to a constant after inlining)
    OC_HD int slot_rel(int lev) const {
#if defined(__CUDA_ARCH__)
        __builtin_assume(sk >= 0 && sk < RS::D);     // lets (lev - k) == 0 fold to sk
#endif
        int t = sk + (lev - k);
        if (POW2) return t & (RS::D - 1);
        t = t < 0 ? t + RS::D : t;
        return t >= RS::D ? t - RS::D : t;
    }
    OC_HD FT operator()(int ii, int jj, int lev) const { return s[slot_rel(lev) * SLOT + (jj - RS::YO) * RS::BX + (ii - RS::XO)]; }
    OC_HD const FT* ptr(int ii, int jj, int lev) const { return s + slot_rel(lev) * SLOT + (jj - RS::YO) * RS::BX + (ii - RS::XO); }
    OC_HD FT* slot(int lev) const { return s + slot_of(lev) * SLOT; }     // absolute (TMA issue, one thread)
    OC_HD static int next_slot(int sk) { return sk + 1 == RS::D ? 0 : sk + 1; }
};

struct MarchSlots {
    int sk[4];    // ring slot of the current level, per ring
};
template <class FT, int CPT>
struct MarchStateT {
    MarchSlots sl;
    FT fz_prev[CPT];   // this thread's upper z-face flux of the previous level (= lower flux of the current one), per owned cell
    FT dfz[CPT];       // δz of the z-face fluxes of the level whose divergence is formed next
    // global-memory operands of the divergence phase, loaded one flux evaluation (~400 instructions) ahead of their use:
    // measured (ncu source page, C4): their consumers carried most of the long-scoreboard stalls
    FT gm[CPT];        // G⁻ of the cell being finished
    FT ph0[CPT], ph1[CPT];   // pHY′ at the cell and at its lower x / y neighbour (u and v kernels)
    int o;        // global offset of this thread's FIRST cell at the level being finished (advanced by one plane per iteration)
    int cell;     // bit h: this thread owns its h-th cell (thread row < TR, inside the domain): loop invariant
    int nit;      // iterations of this block (k_end - k_begin + 1): loop invariant, kept here so that no phase recomputes it
};

// Loads for iteration it + PF are issued in step<1>(it), after every thread has finished iteration it-1; they overwrite the
// slot of level k + PF + HI - D, which must lie below everything iteration it still reads: D >= LIVE + PF, and one more for a
// ring whose level k-1 is read by the divergence phase of iteration it (the Coriolis operand of the u / v kernels).
enum { MARCH_NBAR = 4 };     // data barriers per ring group (velocity rings, ring 0); the prefetch distances and the number of flux stages NSYNC are per kernel

// Which planes each kernel stages.  `FIELD` names the global field: -1 = the stepped field itself (ψ or c), 0/1/2 = u/v/w.
// E = elements per 16 bytes; box x-origins are kept multiples of E (16-byte aligned box rows).
// PF = prefetch distance in levels; ring depth D = LIVE + PF (one more for a ring whose level k-1 is read by the divergence phase).
// Two prefetch distances: PF for ring 0 (the stepped field — its newest plane, level k+3, is first read by the z-face flux at the END
// of the iteration, so its barrier is waited for there) and PFV for the velocity rings 1 … 3 (their newest plane is read by the in-plane
// fluxes at the START of the iteration; the planes are small — no halo rows — so a deeper prefetch is cheap).  Measured reason
// (profiles/r02b_ncu_march_c3_by_line.txt): with one distance of 1 the warps polled the data barrier 12 times per level — the planes
// the iteration starts with arrived late.
template <int KIND, int TX, int TY, int E, int PF_, int PFV_>
struct MarchSpec;
template <int TX, int TY, int E, int PF_, int PFV_>
struct MarchSpec<KIND_C, TX, TY, E, PF_, PFV_> {
    static constexpr int PF = PF_, PFV = PFV_;
    static constexpr int NR = 4;
    using R0 = RingSpec<-4, TX + 8, -3, TY + 6, -2, 3, 6 + PF>;   // c: WENO-5 radius in x, y, z
    using R1 = RingSpec<0, TX + 4, 0, TY, 0, 0, 1 + PFV>;         // u at the x-faces of level k
    using R2 = RingSpec<0, TX, 0, TY + 1, 0, 0, 1 + PFV>;         // v at the y-faces
    using R3 = RingSpec<0, TX, 0, TY, 1, 1, 1 + PFV>;             // w at the upper z-face (level k+1)
    static constexpr int F1 = 0, F2 = 1, F3 = 2;
};
template <int TX, int TY, int E, int PF_, int PFV_>
struct MarchSpec<KIND_U, TX, TY, E, PF_, PFV_> {
    static constexpr int PF = PF_, PFV = PFV_;
    static constexpr int NR = 3;
    using R0 = RingSpec<-4, TX + 8, -3, TY + 6, -2, 3, 6 + PF>;       // u
    using R1 = RingSpec<-E, TX + 2 * E, 0, TY + 1, 0, 0, 2 + PFV>;    // v[i-2..i+1, j0..j0+TY] (Centered-4 along x, Coriolis, τ12); +1: level k-1 is read by the divergence phase
    using R2 = RingSpec<-E, TX + 2 * E, 0, TY, 1, 1, 1 + PFV>;        // w[i-2..i+1] at level k+1
    using R3 = RingSpec<0, 4, 0, 1, 0, 0, 1>;
    static constexpr int F1 = 1, F2 = 2, F3 = -2;
};
template <int TX, int TY, int E, int PF_, int PFV_>
struct MarchSpec<KIND_V, TX, TY, E, PF_, PFV_> {
    static constexpr int PF = PF_, PFV = PFV_;
    static constexpr int NR = 3;
    using R0 = RingSpec<-4, TX + 8, -3, TY + 6, -2, 3, 6 + PF>;   // v
    using R1 = RingSpec<0, TX + 4, -2, TY + 3, 0, 0, 2 + PFV>;    // u[i0..i0+TX, j-2..j+1]; +1: also read by the divergence phase (Coriolis)
    using R2 = RingSpec<0, TX, -2, TY + 3, 1, 1, 1 + PFV>;        // w[j-2..j+1] at level k+1
    using R3 = RingSpec<0, 4, 0, 1, 0, 0, 1>;
    static constexpr int F1 = 0, F2 = 2, F3 = -2;
};
template <int TX, int TY, int E, int PF_, int PFV_>
struct MarchSpec<KIND_W, TX, TY, E, PF_, PFV_> {
    static constexpr int PF = PF_, PFV = PFV_;
    static constexpr int NR = 3;
    using R0 = RingSpec<-4, TX + 8, -3, TY + 6, -2, 3, 6 + PF>;   // w
    using R1 = RingSpec<0, TX + 4, 0, TY, -2, 1, 4 + PFV>;        // u[k-2..k+1] at the x-faces (Centered-4 along z)
    using R2 = RingSpec<0, TX, 0, TY + 1, -2, 1, 4 + PFV>;        // v[k-2..k+1] at the y-faces
    using R3 = RingSpec<0, 4, 0, 1, 0, 0, 1>;
    static constexpr int F1 = 0, F2 = 1, F3 = -2;
};

// ---------------------------------------------------------------------------------------------------------
// compile-time reconstruction coefficients: exact rationals rounded to FT, the last one of each set is
// 1 - sum(others) in FT arithmetic (src/Advection/reconstruction_coefficients.jl:49-64).  p/q with small integers is
// correctly rounded by FT(p)/FT(q).  Model<FT> checks this table against make_coefficients() at construction.
// ---------------------------------------------------------------------------------------------------------
template <class FT>
struct AdvConst {
    static constexpr FT p00 = FT(1) / FT(3), p01 = FT(5) / FT(6), p02 = FT(1) - (p00 + p01);      // WENO{3} coeff_p(r = 0)
    static constexpr FT p10 = FT(-1) / FT(6), p11 = FT(5) / FT(6), p12 = FT(1) - (p10 + p11);     // r = 1
    static constexpr FT p20 = FT(1) / FT(3), p21 = FT(-7) / FT(6), p22 = FT(1) - (p20 + p21);     // r = 2
    static constexpr FT c50 = FT(3) / FT(10), c51 = FT(3) / FT(5), c52 = FT(1) / FT(10);          // C★
    static constexpr FT q00 = FT(1) / FT(2), q01 = FT(1) - q00;                                   // WENO{2} coeff_p
    static constexpr FT q10 = FT(-1) / FT(2), q11 = FT(1) - q10;
    static constexpr FT c30 = FT(2) / FT(3), c31 = FT(1) / FT(3);
    static constexpr FT k0 = FT(-1) / FT(12), k1 = FT(7) / FT(12), k2 = FT(7) / FT(12), k3 = FT(1) - ((k0 + k1) + k2);
    static constexpr FT c40 = k3, c41 = k2, c42 = k1, c43 = k0;                                   // Centered(4), stencil order ψ[i-2..i+1]
    static constexpr FT eps = (FT)1e-8f;
};

// newton_div(Float32, a, b) with the GPU meaning of Base.FastMath.inv_fast: the approximate reciprocal (MUFU.RCP)
// (src/Utils/newton_div.jl:8-23)
OC_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
OC_HD double newton_div_fast(double a, double b) {
    double invd = (double)rcp_fast((float)b);
    double x = a * invd;
    return oc_fma(oc_fma(x, -b, a), invd, x);
}
OC_HD float newton_div_fast(float a, float b) { return a * rcp_fast(b); }

OC_HD float oc_fmaf(float a, float b, float c) {
#ifdef OC_HOSTSIM
    return std::fma(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
OC_HD double fmaT(double a, double b, double c) { return oc_fma(a, b, c); }
OC_HD float fmaT(float a, float b, float c) { return oc_fmaf(a, b, c); }

// 1/x for x >= 1 (a sum of WENO α's): Float32 reciprocal + one Newton step (relative error ~1e-14 in Float64)
OC_HD double rcp_newton(double x) {
    double r = (double)rcp_fast((float)x);
    return oc_fma(r, oc_fma(-x, r, 1.0), r);
}
OC_HD float rcp_newton(float x) { return rcp_fast(x); }

// 1/x in Float64 for any normal x > 0: MUFU.RCP64H seed (≥ 20 good bits, measured through the parity tests) + one cubic
// refinement step (error ~ seed³ ≲ 1e-18·…; the value only needs ~1e-14).  No Float32 round trip, so no exponent-range limit.
OC_HD double rcp_full(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);                 // r (1 + e + e²): cubic convergence, error ~ seed³
    r = fma(r, e, r);
#ifdef OC_RCP_FULL
    e = fma(-x, r, 1.0);
    r = fma(e, r, r);
#endif
    return r;
#else
    return 1.0 / x;
#endif
}

OC_HD double absT(double x) {
#ifdef OC_HOSTSIM
    return std::fabs(x);
#else
    return fabs(x);
#endif
}
OC_HD float absT(float x) {
#ifdef OC_HOSTSIM
    return std::fabs(x);
#else
    return fabsf(x);
#endif
}

OC_HD double rcp_den(double x) { return rcp_full(x); }
OC_HD float rcp_den(float x) { return rcp_fast(x); }

// WENO(order=5) value from the upwind-ordered stencil q0..q4 (weno_interpolants.jl:172-174,204-216,261,290-337,500).
// Algebraically the reference's formula, arranged to minimise FP64 instructions:
//   β_r = ψ1(C1ψ1+C2ψ2+C3ψ3)+ψ2(C4ψ2+C5ψ3)+C6ψ3² = 13/4 (second difference)² + 3/4 (one-sided first difference)²;
//   everything is carried divided by 3/4 (ε too), which leaves τ/(β+ε) unchanged;
//   Σ ω_r p_r with w_r = 1 + (τ/(β_r+ε))², ω_r = C★_r w_r / Σ C★ w (the coeff_p(r) sum to 1, so do the ω): written around the middle
//   candidate, q2 + (p_1 - q2) + ω_0 (p_0 - p_1) + ω_2 (p_2 - p_1), the differences of the candidates are third differences of q
//   with small-integer factors (see the body).
// Float64: the three weights are formed without divisions, w_r ∝ (b_r² + τ²)·(b_s b_t)² (one common factor
// (b0 b1 b2)² cancels in the ratio), leaving ONE reciprocal per face.  Float32 keeps newton_div (a single MUFU.RCP).
// 47 FP64 instructions (51 before round 2's rearrangement of the weights and of the final sum).
template <class FT>
OC_HD FT weno5_value_c(FT q0, FT q1, FT q2, FT q3, FT q4) {
    using K = AdvConst<FT>;
    constexpr FT c133 = FT(13) / FT(3);
    constexpr FT epss = K::eps * (FT(4) / FT(3));
    const FT e1 = q1 - q0, e2 = q2 - q1, e3 = q3 - q2, e4 = q4 - q3;
    const FT d0 = e4 - e3, d1 = e3 - e2, d2 = e2 - e1;
    const FT g0 = fmaT(FT(-3), e3, e4), g1 = e2 + e3, g2 = fmaT(FT(3), e2, -e1);
    const FT b0 = fmaT(c133 * d0, d0, fmaT(g0, g0, epss));
    const FT b1 = fmaT(c133 * d1, d1, fmaT(g1, g1, epss));
    const FT b2 = fmaT(c133 * d2, d2, fmaT(g2, g2, epss));
    const FT tau = b0 - b2;                                  // only τ² is used
    FT w0, w1, w2;
    if (sizeof(FT) == 8) {
        // w_r ∝ (b_r² + τ²) b_s² b_t² = P + τ² b_s² b_t²  with  P = b_0² b_1² b_2²
        const FT T2 = tau * tau;
        const FT B0 = b0 * b0, B1 = b1 * b1, B2 = b2 * b2;
        const FT m12 = B1 * B2, m02 = B0 * B2, m01 = B0 * B1;
        const FT P = B0 * m12;
        w0 = fmaT(T2, m12, P); w1 = fmaT(T2, m02, P); w2 = fmaT(T2, m01, P);
    } else {
        const FT at = absT(tau);
        const FT r0 = newton_div_fast(at, b0), r1 = newton_div_fast(at, b1), r2 = newton_div_fast(at, b2);
        w0 = fmaT(r0, r0, FT(1)); w1 = fmaT(r1, r1, FT(1)); w2 = fmaT(r2, r2, FT(1));
    }
    // Σ ω_r p_r = q2 + (p_1 - q2) + ω_0 (p_0 - p_1) + ω_2 (p_2 - p_1),  ω_r = C★_r w_r / Σ C★ w  (the ω sum to 1):
    //   p_1 - q2 = (e2 + 2 e3)/6,  p_0 - p_1 = -(d0 - d1)/6,  p_2 - p_1 = -2 (d1 - d2)/6,  C★ = (3, 6, 1)/10
    //   = q2 + [ (e2 + 2 e3) - (3 w0 (d0 - d1) + 2 w2 (d1 - d2)) / (3 w0 + 6 w1 + w2) ] / 6
    const FT x = d0 - d1, y = d1 - d2;
    const FT lin = fmaT(FT(2), e3, e2);
    const FT w03 = FT(3) * w0;
    const FT den = fmaT(FT(6), w1, w03 + w2);
    const FT num = fmaT(w2, y + y, w03 * x);
    return fmaT(fmaT(-num, rcp_den(den), lin), FT(1) / FT(6), q2);
}

template <class FT>
OC_HD FT weno3_value_c(FT q0, FT q1, FT q2) {
    using K = AdvConst<FT>;
    FT b0 = q1 * (FT(1) * q1 + FT(-2) * q2) + q2 * q2 * FT(1);
    FT b1 = q0 * (FT(1) * q0 + FT(-2) * q1) + q1 * q1 * FT(1);
    FT tau = absT(b0 - b1);
    FT r0 = newton_div_fast(tau, b0 + K::eps);
    FT r1 = newton_div_fast(tau, b1 + K::eps);
    FT a0 = K::c30 * (FT(1) + r0 * r0);
    FT a1 = K::c31 * (FT(1) + r1 * r1);
    FT rs = FT(1) / (a0 + a1);
    FT p0 = K::q00 * q1 + K::q01 * q2;
    FT p1 = K::q10 * q0 + K::q11 * q1;
    return (a0 * rs) * p0 + (a1 * rs) * p1;
}

// ---------------------------------------------------------------------------------------------------------
// reconstructions on ring accessors (same arithmetic, in the same order, as oc_advection.h)
// ---------------------------------------------------------------------------------------------------------
template <int DIR, class FT, class A>
OC_HD FT rd(const A& a, int ii, int jj, int lev, int n) {
    if (DIR == 0) return a(ii + n, jj, lev);
    if (DIR == 1) return a(ii, jj + n, lev);
    return a(ii, jj, lev + n);
}

// UpwindBiased(order=5) / (order=3) value from the upwind-ordered stencil: the exact rationals (1/30, -13/60, 47/60, 9/20, -1/20) and
// (-1/6, 5/6, 1/3) of uniform_reconstruction_coefficients (reconstruction_coefficients.jl:87-89) with the common denominator pulled
// out — integer constants, one multiplication by 1/60 (1/6)
template <class FT>
OC_HD FT upwind5_value_c(FT q0, FT q1, FT q2, FT q3, FT q4) {
    return (FT(1) / FT(60)) * fmaT(FT(2), q0, fmaT(FT(-13), q1, fmaT(FT(47), q2, fmaT(FT(27), q3, FT(-3) * q4))));
}
template <class FT>
OC_HD FT upwind3_value_c(FT q0, FT q1, FT q2) {
    return (FT(1) / FT(6)) * fmaT(FT(-1), q0, fmaT(FT(5), q1, FT(2) * q2));
}

// _biased_interpolate at face (ii,jj,lev) along DIR; f = global face index along DIR.  LIN = false: the WENO(order=5) chain
// (WENO5 -> WENO3 -> UpwindBiased1); LIN = true: the UpwindBiased(order=5) chain (UB5 -> UB3 -> UB1) — the same stencils and order
// windows (topologically_conditional_interpolation.jl:46-52), linear weights.
// WIN = false: the dimension is not Bounded — always the high-order branch, no window logic.
template <int DIR, bool WIN, bool LIN, class FT, class A>
OC_HD FT t_weno5_biased(const A& a, int ii, int jj, int lev, bool left, int f, const OrderWindow& w) {
    if (!WIN || (f >= w.lo_hi && f <= w.hi_hi)) {
        if (DIR < 2) {
            // in-plane: pick the upwind stencil by address (5 loads) instead of loading 6 values and selecting
            const int st = DIR == 0 ? 1 : A::STRIDE_Y;
            const FT* c = a.ptr(ii, jj, lev);
            const FT* ctr = left ? c - st : c;
            const int sg = left ? st : -st;
            if (LIN) return upwind5_value_c<FT>(ctr[-2 * sg], ctr[-sg], ctr[0], ctr[sg], ctr[2 * sg]);
            return weno5_value_c<FT>(ctr[-2 * sg], ctr[-sg], ctr[0], ctr[sg], ctr[2 * sg]);
        }
        FT m3 = rd<DIR, FT>(a, ii, jj, lev, -3), m2 = rd<DIR, FT>(a, ii, jj, lev, -2), m1 = rd<DIR, FT>(a, ii, jj, lev, -1);
        FT p0 = rd<DIR, FT>(a, ii, jj, lev, 0), p1 = rd<DIR, FT>(a, ii, jj, lev, 1), p2 = rd<DIR, FT>(a, ii, jj, lev, 2);
        if (LIN) return upwind5_value_c<FT>(left ? m3 : p2, left ? m2 : p1, left ? m1 : p0, left ? p0 : m1, left ? p1 : m2);
        return weno5_value_c<FT>(left ? m3 : p2, left ? m2 : p1, left ? m1 : p0, left ? p0 : m1, left ? p1 : m2);
    } else if (f >= w.lo_mid && f <= w.hi_mid) {
        FT m2 = rd<DIR, FT>(a, ii, jj, lev, -2), m1 = rd<DIR, FT>(a, ii, jj, lev, -1);
        FT p0 = rd<DIR, FT>(a, ii, jj, lev, 0), p1 = rd<DIR, FT>(a, ii, jj, lev, 1);
        if (LIN) return upwind3_value_c<FT>(left ? m2 : p1, left ? m1 : p0, left ? p0 : m1);
        return weno3_value_c<FT>(left ? m2 : p1, left ? m1 : p0, left ? p0 : m1);
    }
    return left ? rd<DIR, FT>(a, ii, jj, lev, -1) : rd<DIR, FT>(a, ii, jj, lev, 0);
}

// _symmetric_interpolate of A·q (Centered(4) → Centered(2) near walls)
template <int DIR, bool WIN, class FT, class A>
OC_HD FT t_weno5_symmetric(const A& a, int ii, int jj, int lev, FT area, int f, const OrderWindow& w) {
    if (!WIN || (f >= w.lo_hi && f <= w.hi_hi)) {
        // Centered(4): (-1, 7, 7, -1)/12
        const FT in = rd<DIR, FT>(a, ii, jj, lev, -1) + rd<DIR, FT>(a, ii, jj, lev, 0);
        const FT out = rd<DIR, FT>(a, ii, jj, lev, -2) + rd<DIR, FT>(a, ii, jj, lev, 1);
        return (area * (FT(1) / FT(12))) * fmaT(FT(7), in, -out);
    }
    return FT(0.5) * (area * rd<DIR, FT>(a, ii, jj, lev, -1)) + FT(0.5) * (area * rd<DIR, FT>(a, ii, jj, lev, 0));
}

// The same along z on a vertically stretched grid: a_n = h·Δzᶜ[lev+n] belongs to the advecting velocity's own point
// (Ax_qᶠᶜᶜ / Ay_qᶜᶠᶜ inside ℑzᵃᵃᶠ: upwind_biased_advective_fluxes.jl:79-91); dz points at Δzᶜ of level lev.
template <bool WIN, class FT, class A>
OC_HD FT t_weno5_symmetric_z(const A& a, int ii, int jj, int lev, FT h, const FT* dz, int f, const OrderWindow& w) {
    using K = AdvConst<FT>;
    if (!WIN || (f >= w.lo_hi && f <= w.hi_hi)) {
        FT r = K::c40 * ((h * dz[-2]) * a(ii, jj, lev - 2));
        r = r + K::c41 * ((h * dz[-1]) * a(ii, jj, lev - 1));
        r = r + K::c42 * ((h * dz[0]) * a(ii, jj, lev));
        r = r + K::c43 * ((h * dz[1]) * a(ii, jj, lev + 1));
        return r;
    }
    return FT(0.5) * ((h * dz[-1]) * a(ii, jj, lev - 1)) + FT(0.5) * ((h * dz[0]) * a(ii, jj, lev));
}

// ---------------------------------------------------------------------------------------------------------
// the kernel.  BND = bit mask of possibly-Bounded dimensions: 0 (none: all wall logic compiled out), 4 (z only), 7 (generic).
//              CLO = 0: constant ν, κ (possibly 0: no closure); 1: generic (AMD eddy fields, closure tuples).
//              ADV = ADV_CENTERED2, ADV_WENO5 (the BASELINE schemes) or ADV_UPWIND5 (the WENO-5 stencils with linear weights: the same
//              data movement with a tenth of the FP64 work — the measurement that separates the HBM bound from the FP64 bound).
// One thread per y-face: THREADS = TX·(TY+1); the x-faces and z-faces map onto the same threads so that every
// warp evaluates at most three fluxes per level (no tail warps in front of the barrier).
// ---------------------------------------------------------------------------------------------------------
template <class FT, int ADV, int KIND, int BND, int CLO, int TY_ = 8, int STR = 0>
struct MarchKernel {
    static constexpr int TX = 32, TY = TY_;
    // CPT cells per thread along y: thread row r (0 … TR-1) owns the cells of tile rows r + h·TR, h = 0 … CPT-1.  The per-level fixed
    // costs of a thread (barrier wait / arrive and polling, ring-slot arithmetic, constant reloads, loop control) are paid once per
    // CPT cells — the kernel is bound by instruction issue, and ≈ 40 % of its non-FP64 instructions are such per-level costs.
    static constexpr int TR = 8;
    static constexpr int CPT = TY_ / TR;
    static_assert(CPT * TR == TY_ && (CPT == 1 || CPT == 2), "tile heights 8 (one cell per thread) and 16 (two)");
    static constexpr int THREADS = TX * (TR + 1);
    // 32×8 tiles, one cell per thread: three 288-thread CTAs per SM (the w kernel needs PF = 1 for that: with PF = 2 its three
    // 4-to-6-level rings fit only two, measured 3.5 vs 3.0 ms).  32×16 tiles, two cells per thread: two 288-thread CTAs per SM (1024
    // instead of 768 cells in flight per SM), PF = 1, three flux stages (≈ 100 KB of shared memory per CTA; the w kernel's deeper rings
    // would leave room for one CTA only, so it stays on 32×8 tiles).
    static constexpr int MIN_BLOCKS = TY_ == 8 ? 3 : 2;
    static constexpr int PFK = sizeof(FT) == 4 ? 2 : (TY_ == 8 ? (KIND == KIND_W ? 1 : 2) : 1);      // Float32 planes are half the size: room for 2
    // velocity rings one level deeper where the CTAs per SM still fit: the 32×16-tile kernels (Float64 tracer kernel: 98.9 -> 111.6 KB of
    // the 113 KB two CTAs per SM may use; checked by the static_assert below).  The 32×8-tile w kernel (75.3 KB of 75 KB for three CTAs) cannot.
    static constexpr int PFVK = OC_MARCH_PFV(TY_, KIND, PFK);
    static constexpr int NSYNC = TY_ == 8 ? 4 : 3;
    static constexpr int COMP = KIND == KIND_C ? -1 : KIND;
    static constexpr bool WIN = BND != 0;                 // any wall logic at all
    // BND is a bit mask of the dimensions that MAY be Bounded (7 = generic): order-reduction windows exist only there
    template <int D> static constexpr bool WINV = ((BND >> D) & 1) != 0;
    using SP = MarchSpec<KIND, TX, TY, 16 / (int)sizeof(FT), PFK, PFVK>;
    static_assert(SP::PF < MARCH_NBAR && SP::PFV < MARCH_NBAR, "one barrier per iteration in flight");
    using G0 = Ring<FT, typename SP::R0>;
    using G1 = Ring<FT, typename SP::R1>;
    using G2 = Ring<FT, typename SP::R2>;
    using G3 = Ring<FT, typename SP::R3>;
    static constexpr int NR = SP::NR;
    static constexpr int NFX = (TX + 1) * TY, NFY = TX * (TY + 1), NFZ = TX * TY;
    static constexpr int NFXP = ((NFX + 15) / 16) * 16, NFYP = ((NFY + 15) / 16) * 16;
    // shared-memory map (bytes)
    static constexpr size_t OFF_BAR = 0;
    static constexpr size_t OFF_R0 = 128;
    static constexpr size_t OFF_R1 = OFF_R0 + G0::BYTES;
    static constexpr size_t OFF_R2 = OFF_R1 + G1::BYTES;
    static constexpr size_t OFF_R3 = OFF_R2 + G2::BYTES;
    static constexpr size_t OFF_FX = OFF_R3 + (NR > 3 ? G3::BYTES : 0);
    static constexpr size_t OFF_FY = OFF_FX + sizeof(FT) * NSYNC * NFXP;
    static constexpr size_t SMEM = OFF_FY + sizeof(FT) * NSYNC * NFYP;
    // 228 KB of shared memory per SM, 1 KB reserved per resident CTA
    static_assert(SMEM <= (233472 - MIN_BLOCKS * 1024) / MIN_BLOCKS, "the kernel's shared memory must allow MIN_BLOCKS CTAs per SM");
    // bytes per iteration / of the first iteration, per ring group (C: ring 0, V: rings 1 … 3)
    static constexpr int LEVEL_BYTES_C = G0::BOX_BYTES;
    static constexpr int LEVEL_BYTES_V = G1::BOX_BYTES + G2::BOX_BYTES + (NR > 3 ? G3::BOX_BYTES : 0);
    static constexpr int FIRST_BYTES_C = G0::BOX_BYTES * SP::R0::LIVE;
    static constexpr int FIRST_BYTES_V = G1::BOX_BYTES * SP::R1::LIVE + G2::BOX_BYTES * SP::R2::LIVE + (NR > 3 ? G3::BOX_BYTES * SP::R3::LIVE : 0);

    TendencyArgs<FT> a;
    TileSrc<FT> src[4];   // staged fields, ring order
    int xpad;             // TMA coordinate of interior index i = 0 (j = 0 ↔ H[1], k = 0 ↔ H[2])
    int KC;               // z-levels per chunk (grid.z chunks)
    int by0;              // first tile row of this launch (interior / boundary-strip split of distributed models: oc_model_impl.h tendencies())

    OC_HD int k_begin(const Block& b) const { return b.z * KC; }
    OC_HD int k_end(const Block& b) const { int e = (b.z + 1) * KC; return e < a.g.N[2] ? e : a.g.N[2]; }
    OC_HD int iterations(const Block& b) const { return k_end(b) - k_begin(b) + 1; }   // + the z-flux-only pre-iteration

    // ring views for the iteration at level kk with slot state st (`Ctx` bundles what the flux functions need)
    struct Ctx {
        char* smem;
        int k;
        MarchSlots st;
    };
    OC_HD G0 r0(const Ctx& c) const { return G0{reinterpret_cast<FT*>(c.smem + OFF_R0), c.k, c.st.sk[0]}; }
    OC_HD G1 r1(const Ctx& c) const { return G1{reinterpret_cast<FT*>(c.smem + OFF_R1), c.k, c.st.sk[1]}; }
    OC_HD G2 r2(const Ctx& c) const { return G2{reinterpret_cast<FT*>(c.smem + OFF_R2), c.k, c.st.sk[2]}; }
    OC_HD G3 r3(const Ctx& c) const { return G3{reinterpret_cast<FT*>(c.smem + OFF_R3), c.k, c.st.sk[3]}; }
    OC_HD Ctx raw(char* smem) const { return Ctx{smem, 0, MarchSlots{{0, 0, 0, 0}}}; }

    // ---- metrics.  STR = 1: vertically stretched grid (z Bounded, level tables in Geom); STR = 0: these fold to the constants
    // of the regular grid, so the regular kernels carry no trace of the stretched path.  zf: the point is Face-located in z.
    static constexpr bool ZS = STR != 0;
    static_assert(!ZS || WINV<2>, "a stretched z is Bounded");
    OC_HD FT m_dz(bool zf, int lev) const { return zf ? a.g.dzf[lev] : a.g.dzc[lev]; }
    OC_HD FT m_area(int D, bool zf, int lev) const { return (ZS && D != 2) ? a.g.d[D == 0 ? 1 : 0] * m_dz(zf, lev) : a.g.A[D]; }
    OC_HD FT m_rdz(bool zf, int lev) const { return ZS ? (zf ? a.g.rdzf[lev] : a.g.rdzc[lev]) : a.g.rd[2]; }
    OC_HD FT m_rV(bool zf, int lev) const { return ZS ? (zf ? a.g.rVf[lev] : a.g.rVc[lev]) : a.g.rV; }
    OC_HD FT m_vol(bool zf, int lev) const { return ZS ? a.g.A[2] * m_dz(zf, lev) : a.g.V; }

    // ---- loads -----------------------------------------------------------------------------------------------
    template <class G, class RS>
    OC_DEV void issue_level(const G& ring, const TileSrc<FT>* s, int i0, int j0, int lev, uint64_t* bar) const {
        tile_issue<FT>(ring.slot(lev), s, i0 + RS::XO + xpad, j0 + RS::YO + a.g.H[1], lev + a.g.H[2], bar, RS::BX, RS::BY);
    }
    // the new level every ring of a group needs for the iteration at level k
    OC_HD uint64_t* bar_c(char* smem, int it) const { return reinterpret_cast<uint64_t*>(smem + OFF_BAR) + (it % MARCH_NBAR); }
    OC_HD uint64_t* bar_v(char* smem, int it) const { return reinterpret_cast<uint64_t*>(smem + OFF_BAR) + MARCH_NBAR + (it % MARCH_NBAR); }
    OC_DEV void issue_iteration_c(char* smem, int i0, int j0, int k, uint64_t* bar) const {
        const Ctx c = raw(smem);
        issue_level<G0, typename SP::R0>(r0(c), &src[0], i0, j0, k + SP::R0::HI, bar);
    }
    OC_DEV void issue_iteration_v(char* smem, int i0, int j0, int k, uint64_t* bar) const {
        const Ctx c = raw(smem);
        issue_level<G1, typename SP::R1>(r1(c), &src[1], i0, j0, k + SP::R1::HI, bar);
        issue_level<G2, typename SP::R2>(r2(c), &src[2], i0, j0, k + SP::R2::HI, bar);
        if (NR > 3) issue_level<G3, typename SP::R3>(r3(c), &src[3], i0, j0, k + SP::R3::HI, bar);
    }

    OC_DEV void begin0(const Block&, int tid, char* smem) const {
        if (tid == 0) {
            uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
            for (int n = 0; n < 2 * MARCH_NBAR; ++n) mbar_init(bar + n, 1);
            for (int n = 0; n < NSYNC; ++n) mbar_init(bar + 2 * MARCH_NBAR + n, THREADS / 32);   // one arrival per warp
            mbar_fence_init();
        }
    }
    typedef MarchStateT<FT, CPT> State;
    OC_DEV void begin1(const Block& b, int tid, char* smem, State& stt) const {
        MarchSlots& st = stt.sl;
        {
            const int lane_ = tid & (TX - 1), row_ = tid / TX;
            const int ci = b.x * TX + lane_, cj = (b.y + by0) * TY + row_;
            stt.cell = 0;
            for (int h = 0; h < CPT; ++h) {
                stt.fz_prev[h] = FT(0); stt.dfz[h] = FT(0); stt.gm[h] = FT(0); stt.ph0[h] = FT(0); stt.ph1[h] = FT(0);
                if (row_ < TR && ci < a.g.N[0] && cj + h * TR < a.g.N[1]) stt.cell |= 1 << h;
            }
            stt.o = a.g.idx(ci, cj, k_begin(b) - 2);        // iteration it finishes level k_begin - 2 + it
            stt.nit = iterations(b);
        }
        {   // ring slots of the first level (kf = k_begin - 1)
            const int kf0 = k_begin(b) - 1;
            st.sk[0] = G0::slot_of(kf0); st.sk[1] = G1::slot_of(kf0); st.sk[2] = G2::slot_of(kf0); st.sk[3] = G3::slot_of(kf0);
        }
        if (tid != 0) return;
        const int i0 = b.x * TX, j0 = (b.y + by0) * TY, kf = k_begin(b) - 1, n = iterations(b);
        // iteration 0 (level kf): every live level of every ring
        const Ctx c = raw(smem);
        uint64_t* bv = bar_v(smem, 0);
        mbar_expect(bv, FIRST_BYTES_V);
        for (int l = SP::R1::LO; l < SP::R1::HI; ++l) issue_level<G1, typename SP::R1>(r1(c), &src[1], i0, j0, kf + l, bv);
        for (int l = SP::R2::LO; l < SP::R2::HI; ++l) issue_level<G2, typename SP::R2>(r2(c), &src[2], i0, j0, kf + l, bv);
        if (NR > 3)
            for (int l = SP::R3::LO; l < SP::R3::HI; ++l) issue_level<G3, typename SP::R3>(r3(c), &src[3], i0, j0, kf + l, bv);
        issue_iteration_v(smem, i0, j0, kf, bv);
        uint64_t* bc = bar_c(smem, 0);
        mbar_expect(bc, FIRST_BYTES_C);
        for (int l = SP::R0::LO; l < SP::R0::HI; ++l) issue_level<G0, typename SP::R0>(r0(c), &src[0], i0, j0, kf + l, bc);
        issue_iteration_c(smem, i0, j0, kf, bc);
        for (int it = 1; it < SP::PFV && it < n; ++it) {
            mbar_expect(bar_v(smem, it), LEVEL_BYTES_V);
            issue_iteration_v(smem, i0, j0, kf + it, bar_v(smem, it));
        }
        for (int it = 1; it < SP::PF && it < n; ++it) {
            mbar_expect(bar_c(smem, it), LEVEL_BYTES_C);
            issue_iteration_c(smem, i0, j0, kf + it, bar_c(smem, it));
        }
    }

    // ---- closure fluxes (mirror of TendencyKernel::viscous_flux / diffusive_flux) ---------------------------------
    // velocity component C at tile-local (ii,jj) and level lev, displaced by n along direction DIR
    template <int C, int DIR>
    OC_HD FT vel(const Ctx& smem, int ii, int jj, int lev, int n) const {
        if (C == COMP) return rd<DIR, FT>(r0(smem), ii, jj, lev, n);
        if (C == SP::F1) return rd<DIR, FT>(r1(smem), ii, jj, lev, n);
        return rd<DIR, FT>(r2(smem), ii, jj, lev, n);
    }

    OC_HD FT nu_ff(int o, int d1, int d2) const {
        const Geom<FT>& g = a.g;
        int s1 = g.st(d1), s2 = g.st(d2);
        const FT* n = a.nu_e + o;
        // = ½(½(a+b) + ½(c+d)) bit for bit (scaling by powers of two is exact)
        return FT(0.25) * ((n[-s1 - s2] + n[-s2]) + (n[-s1] + n[0]));
    }

    // νₑ at the flux point of τ_{COMP,D} (ccc for D == COMP, else the edge that is Face in D and COMP), κₑ at the D-face
    // (abstract_scalar_diffusivity_closure.jl:310-332).  Called BEFORE the advective flux is evaluated so that the global loads
    // are in flight during the reconstruction.
    template <int D>
    OC_HD FT eddy_coefficient(int i, int j, int lev) const {
        const Geom<FT>& g = a.g;
        const int o = g.idx(i, j, lev);
        if (KIND == KIND_C) {
            if (!a.kappa_e) return FT(0);
            return FT(0.5) * (a.kappa_e[o - g.st(D)] + a.kappa_e[o]);
        }
        if (!a.nu_e) return FT(0);
        if (D == COMP) return a.nu_e[o];
        return nu_ff(o, D < COMP ? D : COMP, D < COMP ? COMP : D);
    }

    template <int D>
    OC_HD FT viscous_flux(const Ctx& smem, int ii, int jj, int lev, FT nu) const {
        const Geom<FT>& g = a.g;
        FT sig;
        if (D == COMP) {
            sig = (vel<COMP, D>(smem, ii, jj, lev, 1) - vel<COMP, D>(smem, ii, jj, lev, 0)) * (D == 2 ? m_rdz(false, lev) : g.rd[D]);
        } else {
            constexpr int lo = D < COMP ? D : COMP, hi = D < COMP ? COMP : D;
            FT dl = (vel<lo, hi>(smem, ii, jj, lev, 0) - vel<lo, hi>(smem, ii, jj, lev, -1)) * (hi == 2 ? m_rdz(true, lev) : g.rd[hi]);
            FT dh = (vel<hi, lo>(smem, ii, jj, lev, 0) - vel<hi, lo>(smem, ii, jj, lev, -1)) * g.rd[lo];
            sig = FT(0.5) * (dl + dh);
        }
        const FT AD = m_area(D, COMP == 2, lev);
        if (CLO == 0) return (FT(-2) * a.nu * AD) * sig;
        FT flux = FT(0);
        if (a.has_scalar) flux = AD * (FT(-2) * (a.nu * sig));
        if (a.nu_e) {
            FT f2 = AD * (FT(-2) * (nu * sig));
            flux = a.has_scalar ? flux + f2 : f2;
        }
        return flux;
    }

    template <int D>
    OC_HD FT diffusive_flux(const Ctx& smem, int ii, int jj, int lev, FT kap) const {
        const Geom<FT>& g = a.g;
        G0 c = r0(smem);
        FT grad = (rd<D, FT>(c, ii, jj, lev, 0) - rd<D, FT>(c, ii, jj, lev, -1)) * (D == 2 ? m_rdz(true, lev) : g.rd[D]);
        const FT AD = m_area(D, false, lev);
        if (CLO == 0) return -(a.kappa * AD) * grad;   // (folded with the advective part in total_flux for CLO == 0)
        FT flux = FT(0);
        if (a.has_scalar) flux = AD * (-(a.kappa * grad));
        if (a.kappa_e) {
            FT f2 = AD * (-(kap * grad));
            flux = a.has_scalar ? flux + f2 : f2;
        }
        return flux;
    }

    // ---- advective flux through the faces normal to D at flux index (ii, jj, lev); id = global index along D, ic along COMP
    template <int D>
    OC_HD FT advective_flux(const Ctx& smem, int ii, int jj, int lev, int id, int ic) const {
        const Geom<FT>& g = a.g;
        // Centered: the area at the flux point; upwind schemes: the area of the advecting velocity's own point (they differ only
        // for the x / y fluxes of w on a stretched grid, which take the per-level areas inside the z interpolation below)
        const FT A = m_area(D, COMP == 2 && ADV == 0, lev);
        if (KIND == KIND_C) {
            FT u = D == 0 ? r1(smem)(ii, jj, lev) : (D == 1 ? r2(smem)(ii, jj, lev) : r3(smem)(ii, jj, lev));
            G0 c = r0(smem);
            if (ADV == 0) {
                return (A * u) * (FT(0.5) * rd<D, FT>(c, ii, jj, lev, -1) + FT(0.5) * rd<D, FT>(c, ii, jj, lev, 0));
            } else {
                OrderWindow w;
                if (WINV<D>) w = order_window(g.wlo[D] != 0, g.whi[D] != 0, false, g.N[D]);
                FT cr = t_weno5_biased<D, WINV<D>, ADV == ADV_UPWIND5, FT>(c, ii, jj, lev, u > FT(0), id, w);
                return A * u * cr;
            }
        } else {
            constexpr int CC = COMP < 0 ? 0 : COMP;
            G0 psi = r0(smem);
            if (D == CC) {
                // centre-type: the face-type stencils evaluated at face id+1
                if (ADV == 0) {
                    FT ut = FT(0.5) * rd<D, FT>(psi, ii, jj, lev, 0) + FT(0.5) * rd<D, FT>(psi, ii, jj, lev, 1);
                    return A * ut * ut;
                } else {
                    OrderWindow w;
                    if (WINV<D>) w = order_window(g.wlo[D] != 0, g.whi[D] != 0, true, g.N[D]);
                    const int i1 = ii + (D == 0), j1 = jj + (D == 1), l1 = lev + (D == 2);
                    FT ut = t_weno5_symmetric<D, WINV<D>, FT>(psi, i1, j1, l1, A, id + 1, w);
                    FT pr = t_weno5_biased<D, WINV<D>, ADV == ADV_UPWIND5, FT>(psi, i1, j1, l1, ut > FT(0), id + 1, w);
                    return ut * pr;
                }
            } else {
                // face-type: advecting velocity U[D] interpolated along CC, ψ reconstructed along D
                if (ADV == 0) {
                    FT ut = FT(0.5) * vel<D, CC>(smem, ii, jj, lev, -1) + FT(0.5) * vel<D, CC>(smem, ii, jj, lev, 0);
                    FT pt = FT(0.5) * rd<D, FT>(psi, ii, jj, lev, -1) + FT(0.5) * rd<D, FT>(psi, ii, jj, lev, 0);
                    return A * ut * pt;
                } else {
                    OrderWindow wc, wd;
                    if (WINV<CC>) wc = order_window(g.wlo[CC] != 0, g.whi[CC] != 0, false, g.N[CC]);
                    if (WINV<D>) wd = order_window(g.wlo[D] != 0, g.whi[D] != 0, false, g.N[D]);
                    FT ut;
                    if (ZS && CC == 2) {
                        const FT h = g.d[D == 0 ? 1 : 0];
                        if (D == SP::F1) ut = t_weno5_symmetric_z<WINV<CC>, FT>(r1(smem), ii, jj, lev, h, g.dzc + lev, ic, wc);
                        else ut = t_weno5_symmetric_z<WINV<CC>, FT>(r2(smem), ii, jj, lev, h, g.dzc + lev, ic, wc);
                    } else if (D == SP::F1) ut = t_weno5_symmetric<CC, WINV<CC>, FT>(r1(smem), ii, jj, lev, A, ic, wc);
                    else ut = t_weno5_symmetric<CC, WINV<CC>, FT>(r2(smem), ii, jj, lev, A, ic, wc);
                    FT pr = t_weno5_biased<D, WINV<D>, ADV == ADV_UPWIND5, FT>(psi, ii, jj, lev, ut > FT(0), id, wd);
                    return ut * pr;
                }
            }
        }
    }

    // i, j, k: global flux index; (ii, jj): the same, tile-local
    template <int D>
    OC_HD FT total_flux(const Ctx& smem, int ii, int jj, int i, int j, int k) const {
        const int id = D == 0 ? i : (D == 1 ? j : k);
        const int ic = COMP == 0 ? i : (COMP == 1 ? j : k);
        FT eddy = FT(0);
        if (CLO != 0) eddy = eddy_coefficient<D>(i, j, k);
        FT F = advective_flux<D>(smem, ii, jj, k, id, ic);
        if (CLO == 0 && KIND == KIND_C) {
            // constant κ: F - (κ A / Δ)(c[0] - c[-1]) with the constant folded (one subtraction and one FMA)
            G0 c = r0(smem);
            const FT dc = rd<D, FT>(c, ii, jj, k, 0) - rd<D, FT>(c, ii, jj, k, -1);
            return fmaT(-(a.kappa * m_area(D, false, k) * (D == 2 ? m_rdz(true, k) : a.g.rd[D])), dc, F);
        }
        if (CLO == 0 || a.has_scalar || a.nu_e || a.kappa_e) {
            if constexpr (KIND == KIND_C) F = F + diffusive_flux<D>(smem, ii, jj, k, eddy);
            else F = F + viscous_flux<D>(smem, ii, jj, k, eddy);
        }
        return F;
    }

    // ---- one level, software-pipelined ----------------------------------------------------------------------------------
    // The marching loop (oc_exec.h: march_entry) runs, for it = 0 … n:
    //     step<0>(it)   wait for the TMA data of level k(it); x- and y-face fluxes of the level -> stage it % 4
    //     -- wait until every thread has ARRIVED for iteration it-1 (split mbarrier: nobody blocks at the arrive) --
    //     step<1>(it)   thread 0 issues the loads of iteration it+2; divergence / substep / stores of level k(it-1)
    //     step<2>(it)   upper z-face flux of level k(it) (kept in registers: the same thread owns the cell above and below)
    //     -- arrive for iteration it --
    // so a warp only stalls when another warp is a whole level behind.  Hazards: flux stage it % 4 is rewritten in
    // step<0>(it+4), which is only reached after the wait for iteration it+2, i.e. after every thread finished
    // step<1>(it+1) (the last reader).  Ring slots: the loads issued in step<1>(it) overwrite levels that were last read in
    // iteration it-1 (see MarchSpec), whose arrivals have all been observed.
    OC_HD uint64_t* sync_bar(char* smem, int it) const { return reinterpret_cast<uint64_t*>(smem + OFF_BAR) + 2 * MARCH_NBAR + (it % NSYNC); }
    OC_HD static int sync_parity(int it) { return (it / NSYNC) & 1; }

    OC_DEV void sync_wait(char* smem, int it) const { mbar_wait(sync_bar(smem, it), sync_parity(it)); }
    // one arrival per warp (after __syncwarp, so the elected lane's release covers the whole warp's flux stores): every arrival
    // wakes the sleeping waiters, so 9 arrivals per level instead of 288 cut the re-polling by an order of magnitude
    OC_DEV void sync_arrive(char* smem, int it) const {
#if defined(__CUDA_ARCH__)
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(sync_bar(smem, it));
#else
        mbar_arrive(sync_bar(smem, it));
#endif
    }

    template <int PHASE>
    OC_DEV void step(const Block& b, int tid, char* smem, int it, State& stt) const {
        const Geom<FT>& g = a.g;
        const int i0 = b.x * TX, j0 = (b.y + by0) * TY;
        const int nit = stt.nit;
        const int k = k_begin(b) - 1 + it;                   // level of this iteration (it = 0: z-flux only)
        const Ctx cx{smem, k, stt.sl};
        constexpr int shx = COMP == 0 ? -1 : 0, shy = COMP == 1 ? -1 : 0, shz = COMP == 2 ? -1 : 0;
        const int lane = tid & (TX - 1), row = tid / TX;     // row = 0 … TR
        if (PHASE == 0) {
            if (it >= 2) {
                // operands of step<1>(it), which finishes level k-1 for this thread's cells — also in the last iteration (it == nit)
#pragma unroll
                for (int h = 0; h < CPT; ++h) {
                    if (!((stt.cell >> h) & 1)) continue;
                    const int op = stt.o + h * TR * g.sy;
                    if (a.mode == STEP_RK3 || (a.mode == STEP_AB2 && !a.ab2_euler)) stt.gm[h] = a.Gm[op];
                    if ((KIND == KIND_U || KIND == KIND_V) && a.pHY) {
                        stt.ph0[h] = a.pHY[op];
                        stt.ph1[h] = a.pHY[op - (KIND == KIND_U ? 1 : g.sy)];
                    }
                }
            }
            if (it >= nit) return;
            mbar_wait(bar_v(smem, it), (it / MARCH_NBAR) & 1);       // the velocity planes of this level
            if (it == 0) return;
            FT* fx = reinterpret_cast<FT*>(smem + OFF_FX) + (it % NSYNC) * NFXP;
            FT* fy = reinterpret_cast<FT*>(smem + OFF_FY) + (it % NSYNC) * NFYP;
            // y-faces: thread row r takes the face rows r + h·(TR+1) ≤ TY.  CLO == 0: every operand is in shared memory, so
            // out-of-range faces of partial tiles are evaluated too (on zero-filled / neighbouring data) and discarded — no
            // divergent branch
#pragma unroll
            for (int h = 0; h < CPT; ++h) {
                const int fr = row + h * (TR + 1);
                if (h > 0 && fr > TY) continue;
                const bool ok = i0 + lane < g.N[0] && j0 + fr <= g.N[1];
                FT F = FT(0);
                if (CLO == 0 || ok) F = total_flux<1>(cx, lane, fr + shy, i0 + lane, j0 + fr + shy, k);
                fy[fr * TX + lane] = (CLO == 0 || ok) ? F : FT(0);      // faces outside the grid are never consumed
            }
            // x-faces: thread rows 0 … TR-1 take the faces s = 0 … TX-1 of their tile rows; the last warp takes the TY faces s = TX
            if (row < TR) {
#pragma unroll
                for (int h = 0; h < CPT; ++h) {
                    const int s = lane, jj = row + h * TR;
                    const bool ok = j0 + jj < g.N[1] && i0 + s <= g.N[0];
                    FT F = FT(0);
                    if (CLO == 0 || ok) F = total_flux<0>(cx, s + shx, jj, i0 + s + shx, j0 + jj, k);
                    fx[jj * (TX + 1) + s] = (CLO == 0 || ok) ? F : FT(0);
                }
            } else if (lane < TY) {
                const int s = TX, jj = lane;
                const bool ok = j0 + jj < g.N[1] && i0 + s <= g.N[0];
                FT F = FT(0);
                if (CLO == 0 || ok) F = total_flux<0>(cx, s + shx, jj, i0 + s + shx, j0 + jj, k);
                fx[jj * (TX + 1) + s] = (CLO == 0 || ok) ? F : FT(0);
            }
        } else if (PHASE == 2) {
            if (it >= nit) return;
            mbar_wait(bar_c(smem, it), (it / MARCH_NBAR) & 1);       // ring 0's newest plane (level k+3): first read here
            if (row < TR) {   // upper z-faces of this thread's cells
#pragma unroll
                for (int h = 0; h < CPT; ++h) {
                    const int jj = row + h * TR;
                    const bool ok = i0 + lane < g.N[0] && j0 + jj < g.N[1] && k + 1 <= g.N[2];
                    FT F = FT(0);
                    if (CLO == 0 || ok) F = total_flux<2>(cx, lane, jj, i0 + lane, j0 + jj, k + 1 + shz);
                    F = (CLO == 0 || ok) ? F : FT(0);
                    stt.dfz[h] = F - stt.fz_prev[h];
                    stt.fz_prev[h] = F;
                }
            }
            // advance the ring slots to the next level
            stt.sl.sk[0] = G0::next_slot(stt.sl.sk[0]); stt.sl.sk[1] = G1::next_slot(stt.sl.sk[1]);
            stt.sl.sk[2] = G2::next_slot(stt.sl.sk[2]); stt.sl.sk[3] = G3::next_slot(stt.sl.sk[3]);
        } else {
            const int o_first = stt.o;
            stt.o = o_first + g.sz;
            if (tid == 0) {
                const int lit = it + SP::PF, litv = it + SP::PFV;
                if (lit < nit || litv < nit) proxy_fence_async();
                if (litv < nit) {
                    mbar_expect(bar_v(smem, litv), LEVEL_BYTES_V);
                    issue_iteration_v(smem, i0, j0, k + SP::PFV, bar_v(smem, litv));
                }
                if (lit < nit) {
                    mbar_expect(bar_c(smem, lit), LEVEL_BYTES_C);
                    issue_iteration_c(smem, i0, j0, k + SP::PF, bar_c(smem, lit));
                }
            }
            if (it < 2 || !stt.cell) return;                 // levels start at iteration 1; their divergence is formed one iteration later
            const int kc = k - 1;                            // the level being finished
            const FT* fx = reinterpret_cast<const FT*>(smem + OFF_FX) + ((it - 1) % NSYNC) * NFXP;
            const FT* fy = reinterpret_cast<const FT*>(smem + OFF_FY) + ((it - 1) % NSYNC) * NFYP;
#pragma unroll
            for (int h = 0; h < CPT; ++h) {
                if (!((stt.cell >> h) & 1)) continue;
                const int ii = lane, jj = row + h * TR;
                const int i = i0 + ii, j = j0 + jj;
                const int o = o_first + h * TR * g.sy;
                const FT u0 = r0(cx)(ii, jj, kc);
                if (WIN && COMP >= 0) {
                    // exclude_periphery: the wall face of a wall-normal velocity is not stepped (kernel_launching.jl:145-146)
                    const int ic = COMP == 0 ? i : (COMP == 1 ? j : kc);
                    if (g.wlo[COMP < 0 ? 0 : COMP] && ic == 0 && g.N[COMP < 0 ? 0 : COMP] > 1) {
                        if (a.mode != STEP_NONE) a.Unew[o] = u0;
                        continue;
                    }
                }
                const FT dFx = fx[jj * (TX + 1) + ii + 1] - fx[jj * (TX + 1) + ii];
                const FT dFy = fy[(jj + 1) * TX + ii] - fy[jj * TX + ii];
                const FT dFz = stt.dfz[h];
                FT G = -(m_rV(COMP == 2, kc) * (dFx + dFy + dFz));
                if ((KIND == KIND_U || KIND == KIND_V) && a.has_coriolis) {
                    // FPlane (f_plane.jl:50-52) and BetaPlane (beta_plane.jl:56-72: the same with f = f₀ + β·ynode of the velocity point,
                    // one more FMA); the other horizontal component is ring 1
                    FT fj = a.f;
                    if (a.has_coriolis == 2) fj = a.f + a.cor_beta * (a.cor_y0 + (FT(j) + (KIND == KIND_U ? FT(0.5) : FT(0))) * g.d[1]);
                    G1 q = r1(cx);
                    FT num, cnt = FT(1);
                    if (KIND == KIND_U) {
                        num = FT(0.25) * ((q(ii - 1, jj, kc) + q(ii, jj, kc)) + (q(ii - 1, jj + 1, kc) + q(ii, jj + 1, kc)));
                        if (WIN) {
                            int ax0 = !(g.wlo[0] && (i - 1 < 0)), ax1 = 1;
                            int ay0 = !(g.wlo[1] && (j < 1)), ay1 = !(g.whi[1] && (j + 1 > g.N[1] - 1));
                            cnt = FT(0.5) * (FT(0.5) * FT(ax0 * ay0 + ax1 * ay0) + FT(0.5) * FT(ax0 * ay1 + ax1 * ay1));
                        }
                        FT val = cnt == FT(0) ? FT(0) : num / cnt;
                        G = G - (-fj * val);
                    } else {
                        num = FT(0.25) * ((q(ii, jj - 1, kc) + q(ii + 1, jj - 1, kc)) + (q(ii, jj, kc) + q(ii + 1, jj, kc)));
                        if (WIN) {
                            int ax0 = !(g.wlo[0] && (i < 1)), ax1 = !(g.whi[0] && (i + 1 > g.N[0] - 1));
                            int ay0 = !(g.wlo[1] && (j - 1 < 0)), ay1 = 1;
                            cnt = FT(0.5) * (FT(0.5) * FT(ax0 * ay0 + ax1 * ay0) + FT(0.5) * FT(ax0 * ay1 + ax1 * ay1));
                        }
                        FT val = cnt == FT(0) ? FT(0) : num / cnt;
                        G = G - (fj * val);
                    }
                }
                if ((KIND == KIND_U || KIND == KIND_V) && a.pHY) G = G - (stt.ph0[h] - stt.ph1[h]) * g.rd[KIND == KIND_U ? 0 : 1];
                if (WIN && a.add_flux_bcs) {
                    const int ijk[3] = {i, j, kc};
                    for (int d = 0; d < 3; ++d) {
                        const FT Ad = m_area(d, COMP == 2, kc), Vd = m_vol(COMP == 2, kc);
                        if (a.fbc.on[2 * d] && ijk[d] == 0) G = G + a.fbc.val[2 * d] * Ad / Vd;
                        if (a.fbc.on[2 * d + 1] && ijk[d] == g.N[d] - 1) G = G - a.fbc.val[2 * d + 1] * Ad / Vd;
                    }
                }
                a.Gn[o] = G;
                if (a.mode == STEP_RK3_FIRST) {
                    a.Unew[o] = u0 + a.ca * G;
                } else if (a.mode == STEP_RK3) {
                    a.Unew[o] = u0 + a.dt * (a.ca * G + a.cb * stt.gm[h]);
                } else if (a.mode == STEP_AB2) {
                    FT Gu = a.ab2_euler ? a.ca * G : a.ca * G - a.cb * stt.gm[h];
                    a.Unew[o] = u0 + a.dt * Gu;
                }
            }
        }
    }
};

}  // namespace oc
