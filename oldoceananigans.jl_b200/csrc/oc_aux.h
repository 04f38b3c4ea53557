// oc_aux.h — auxiliary-field kernels of update_state!: hydrostatic pressure anomaly, AMD and Smagorinsky(-Lilly) eddy
// viscosity / diffusivities.
//
// Replaces _update_hydrostatic_pressure! (src/Models/NonhydrostaticModels/update_hydrostatic_pressure.jl:12-49)
// and _compute_AMD_viscosity! / _compute_AMD_diffusivity!
// (src/TurbulenceClosures/turbulence_closure_implementations/anisotropic_minimum_dissipation.jl:154-351 with
//  velocity_tracer_gradients.jl:126-250).
#pragma once
#include "oc_common.h"

namespace oc {

template <class FT>
struct HydrostaticPressureKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* pHY;
    const FT* bT;
    const FT* bS;
    int buoyancy;          // 1 tracer, 2 seawater linear
    FT grav, alpha, beta;
    int tilted;            // BuoyancyForce with a gravity_unit_vector: z_dot_g_b = ĝ_z ℑzᶠ b   g_dot_b.jl:3
    FT gz;
    int ilo, ni, jlo, nj;  // column range: 0:N+1 (i.e. -1..N) unless Flat (p_kernel_parameters :41-49)

    OC_HD FT b_at(int o) const {
        if (buoyancy == 1) return bT[o];
        return grav * (alpha * bT[o] - beta * bS[o]);
    }
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int ii = b.x * nt + tid;
        if (ii >= ni || b.y >= nj) return;
        int i = ilo + ii, j = jlo + b.y;
        int Nz = g.N[2];
        int o = g.idx(i, j, Nz);                       // k = Nz (reference index Nz+1): the halo/periodic image above the top
        FT bup = b_at(o);
        FT p = FT(0);
#pragma unroll 8
        for (int k = Nz - 1; k >= 0; --k) {
            o -= g.sz;
            FT bk = b_at(o);
            FT bf = FT(0.5) * (bk + bup);              // z_dot_g_bᶜᶜᶠ(k+1) = ℑzᵃᵃᶠ b   g_dot_b.jl:4
            if (tilted) bf = gz * bf;
            const FT dzf = g.dz_at(true, k + 1);       // Δzᶜᶜᶠ of the face above cell k   update_hydrostatic_pressure.jl:15-19
            p = (k == Nz - 1) ? -bf * dzf : p - bf * dzf;
            pHY[o] = p;
            bup = bk;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// AMD.  One thread per cell; gradients are re-derived from u, v, w (stencil radius 1 around the cell).
// The helper struct mirrors the reference's operator names so each term can be checked line by line.
// ---------------------------------------------------------------------------------------------------------
template <class FT, bool STR = false, bool CB = false>
struct AmdKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;      // 32 (x) × 8 (y) cells per CTA: the 27-point neighbourhoods share L1 lines in x AND y
    static constexpr int MIN_BLOCKS = 3;     // cap registers at 80: the kernel is latency-bound at the 110 it would otherwise take
    Geom<FT> g;
    const FT* u;
    const FT* v;
    const FT* w;
    FT* nu_e;
    FT Cnu;
    int ntr;
    const FT* c[8];
    FT* kappa_e[8];
    FT Ckappa[8];
    FT ratios[6];      // Δᶠa/Δᶠb   (:224-226)
    FT delta2;         // δ² = 3 / (1/Δᶠx² + 1/Δᶠy² + 1/Δᶠz²)   (:166,190)
    // stretched grids: per-level tables of the constants that contain divisions by Δzᶜ[k], formed once on the host in the same
    // FT arithmetic (Model::build_z_tables): kxw = (Δᶠx/Δᶠz)/Δx, kzu = (Δᶠz/Δᶠx)/Δzᶠ, kyw = (Δᶠy/Δᶠz)/Δy, kzv = (Δᶠz/Δᶠy)/Δzᶠ,
    // kcz = Δᶠz/Δzᶠ, d2 = δ²; indexable like Geom::dzc
    const FT* lv_kxw; const FT* lv_kzu; const FT* lv_kyw; const FT* lv_kzv; const FT* lv_kcz; const FT* lv_d2;
    // CB variant only — the buoyancy modification (Cb ≠ nothing, :62-68): b = buoyancy_perturbationᶜᶜᶜ from the tracers
    FT Cb;
    int buoyancy;        // 1 tracer b, 2 seawater linear (CB is never instantiated without buoyancy: the term is zero)
    const FT* bT;
    const FT* bS;
    FT grav, alpha, beta;
    OC_HD FT b_at(int p) const {                                                          // buoyancy_tracer.jl:12 ; seawater_buoyancy.jl:203-207
        if (buoyancy == 1) return bT[p];
        return grav * (alpha * bT[p] - beta * bS[p]);
    }
    void set_consts() {
        FT fx = FT(2) * g.d[0], fy = FT(2) * g.d[1], fz = FT(2) * g.d[2];
        ratios[0] = fx / fy; ratios[1] = fy / fx; ratios[2] = fx / fz; ratios[3] = fz / fx; ratios[4] = fy / fz; ratios[5] = fz / fy;
        delta2 = FT(3) / (FT(1) / (fx * fx) + FT(1) / (fy * fy) + FT(1) / (fz * fz));
    }

    // Every normalised gradient is evaluated ONCE at the four corners its interpolations need (24 evaluations per cell).
    // The reference's nested ℑ(…) expressions (anisotropic_minimum_dissipation.jl:240-351) are then assembled from 15 corner
    // sums instead of 30 separately interpolated products — an exact algebraic identity (ℑ of four corners = ¼ Σ; products with
    // Σ₁₂ = ½(∂y u + ∂x v) expand into the sums of squares and one cross sum per plane):
    //   A1 = Σ ∂x v, B1 = Σ ∂y u, A2 = Σ (∂x v)², B2 = Σ (∂y u)², AB = Σ ∂x v ∂y u over the xy corners; C·, D· (∂x w, ∂z u) over xz;
    //   E·, F· (∂y w, ∂z v) over yz (all normalised gradients);
    //   q = S11² + S22² + S33² + ¼ (A2 + B2 + C2 + D2 + E2 + F2)                                              (:285-306)
    //   r = S11³ + S22³ + S33³ + ¼ [S11 (Sxy + Sxz) + S22 (Sxy + Syz) + S33 (Sxz + Syz)]
    //       + (1/64) [A1 C1 (E1 + F1) + B1 E1 (C1 + D1) + D1 F1 (A1 + B1)],   Sxy = A2 + B2 + AB, …              (:240-279)
    //   σ = ½ (X2 + Y2 + Z2),  θ = ½ (S11 X2 + S22 Y2 + S33 Z2) + (1/16) [X1 Y1 (A1 + B1) + X1 Z1 (C1 + D1) + Y1 Z1 (G1 + F1)]
    //       with X1 = Σ ∂x c, X2 = Σ (∂x c)² over the two x-faces, …, and G1 = Σ ∂y w over the XZ corners (sic, :336)   (:322-351)
    // Measured (ncu): the kernel is bound by the FP64 pipe (60 %); this form needs ≈ 300 FP64 instructions per cell instead of 456.
    // Differences to the reference's order of operations are rounding-level (≈ 1e-16 relative per term).
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        (void)nt;
        int i = b.x * 32 + (tid & 31), j = b.y * 8 + (tid >> 5), k = b.z;
        if (i >= g.N[0] || j >= g.N[1]) return;
        const int o = g.idx(i, j, k);
        const int sx = 1, sy = g.sy, sz = g.sz;
        const FT fx = FT(2) * g.d[0], fy = FT(2) * g.d[1];
        const FT d2 = STR ? lv_d2[k] : delta2;                                             // δ² (:166,190)
        // combined constants: (filter-width ratio) × (reciprocal spacing); [0] on the cell's level, [1] on the level above
        const FT kxv = ratios[0] * g.rd[0], kyu = ratios[1] * g.rd[1];
        FT kxw[2], kzu[2], kyw[2], kzv[2], kcz[2];
        for (int n = 0; n < 2; ++n) {
            if (STR) { kxw[n] = lv_kxw[k + n]; kzu[n] = lv_kzu[k + n]; kyw[n] = lv_kyw[k + n]; kzv[n] = lv_kzv[k + n]; kcz[n] = lv_kcz[k + n]; }
            else {
                kxw[n] = ratios[2] * g.rd[0]; kzu[n] = ratios[3] * g.rd[2]; kyw[n] = ratios[4] * g.rd[1]; kzv[n] = ratios[5] * g.rd[2];
                kcz[n] = (FT(2) * g.d[2]) * g.rd[2];
            }
        }
        const FT rdzc = STR ? g.rdzc[k] : g.rd[2];
        const int cxy[4] = {o, o + sx, o + sy, o + sx + sy};
        const int cxz[4] = {o, o + sx, o + sz, o + sx + sz};
        const int cyz[4] = {o, o + sy, o + sz, o + sy + sz};
        FT A1 = FT(0), B1 = FT(0), A2 = FT(0), B2 = FT(0), AB = FT(0);
        FT C1 = FT(0), D1 = FT(0), C2 = FT(0), D2 = FT(0), CD = FT(0);
        FT E1 = FT(0), F1 = FT(0), E2 = FT(0), F2 = FT(0), EF = FT(0);
        FT G1 = FT(0);
        for (int n = 0; n < 4; ++n) {
            const int up = n >> 1;                      // corners 2, 3 of the xz / yz sets are on the level above
            const int pxy = cxy[n], pxz = cxz[n], pyz = cyz[n];
            const FT dxv = (v[pxy] - v[pxy - sx]) * kxv, dyu = (u[pxy] - u[pxy - sy]) * kyu;           // ffc
            const FT dxw = (w[pxz] - w[pxz - sx]) * kxw[up], dzu = (u[pxz] - u[pxz - sz]) * kzu[up];   // fcf
            const FT dyw = (w[pyz] - w[pyz - sy]) * kyw[up], dzv = (v[pyz] - v[pyz - sz]) * kzv[up];   // cff
            A1 += dxv; B1 += dyu; A2 += dxv * dxv; B2 += dyu * dyu; AB += dxv * dyu;
            C1 += dxw; D1 += dzu; C2 += dxw * dxw; D2 += dzu * dzu; CD += dxw * dzu;
            E1 += dyw; F1 += dzv; E2 += dyw * dyw; F2 += dzv * dzv; EF += dyw * dzv;
            if (ntr > 0) G1 += (w[pxz] - w[pxz - sy]) * kyw[up];                                       // ∂y w at the xz corners (:336)
        }
        const FT S11 = (u[o + sx] - u[o]) * g.rd[0], S22 = (v[o + sy] - v[o]) * g.rd[1], S33 = (w[o + sz] - w[o]) * rdzc;   // ccc
        const FT q11 = S11 * S11, q22 = S22 * S22, q33 = S33 * S33;
        const FT q = (q11 + q22 + q33) + FT(0.25) * (((A2 + B2) + (C2 + D2)) + (E2 + F2));
        FT nu = FT(0);
        if (q != FT(0)) {
            const FT Sxy = A2 + B2 + AB, Sxz = C2 + D2 + CD, Syz = E2 + F2 + EF;
            const FT cubes = S11 * q11 + S22 * q22 + S33 * q33;
            const FT mixed = S11 * (Sxy + Sxz) + S22 * (Sxy + Syz) + S33 * (Sxz + Syz);
            const FT triple = A1 * C1 * (E1 + F1) + B1 * E1 * (C1 + D1) + D1 * F1 * (A1 + B1);
            const FT r = cubes + FT(0.25) * mixed + FT(0.015625) * triple;
            if (CB) {
                // Cb ζ = Cb (ℑxz norm_∂x_w · Δᶠx ℑx ∂x b + ℑyz norm_∂y_w · Δᶠy ℑy ∂y b + ∂z w · Δᶠz ℑz ∂z b) / Δᶠz        (:168, :310-323)
                const FT b0 = b_at(o);
                const FT bx = FT(0.5) * ((b0 - b_at(o - sx)) * g.rd[0] + (b_at(o + sx) - b0) * g.rd[0]);
                const FT by = FT(0.5) * ((b0 - b_at(o - sy)) * g.rd[1] + (b_at(o + sy) - b0) * g.rd[1]);
                const FT rzf0 = STR ? g.rdzf[k] : g.rd[2], rzf1 = STR ? g.rdzf[k + 1] : g.rd[2];
                const FT bz = FT(0.5) * ((b0 - b_at(o - sz)) * rzf0 + (b_at(o + sz) - b0) * rzf1);
                const FT fz = FT(2) * (STR ? g.dzc[k] : g.d[2]);
                const FT wb = (FT(0.25) * C1) * fx * bx + (FT(0.25) * E1) * fy * by + S33 * fz * bz;
                nu = -Cnu * d2 * (r - Cb * wb / fz) / q;
            } else
            nu = -Cnu * d2 * r / q;                                                        // Cb = nothing: no buoyancy term (:168,281)
        }
        nu_e[o] = oc_max<FT>(FT(0), nu);
        if (ntr == 0) return;
        const FT kcx = fx * g.rd[0], kcy = fy * g.rd[1];
        for (int tr = 0; tr < ntr; ++tr) {
            const FT* cc = c[tr];
            const FT c0 = cc[o];
            const FT cx0 = (c0 - cc[o - sx]) * kcx, cx1 = (cc[o + sx] - c0) * kcx;        // norm_∂x_c at the two x-faces (fcc)
            const FT cy0 = (c0 - cc[o - sy]) * kcy, cy1 = (cc[o + sy] - c0) * kcy;
            const FT cz0 = (c0 - cc[o - sz]) * kcz[0], cz1 = (cc[o + sz] - c0) * kcz[1];  // ccf: the upper face is on the level above
            const FT X1 = cx0 + cx1, Y1 = cy0 + cy1, Z1 = cz0 + cz1;
            const FT X2 = cx0 * cx0 + cx1 * cx1, Y2 = cy0 * cy0 + cy1 * cy1, Z2 = cz0 * cz0 + cz1 * cz1;
            const FT sigma = FT(0.5) * (X2 + Y2 + Z2);                                     // norm_θᵢ²ᶜᶜᶜ :349-351
            FT kap = FT(0);
            if (sigma != FT(0)) {
                const FT theta = FT(0.5) * (S11 * X2 + S22 * Y2 + S33 * Z2)
                               + FT(0.0625) * (X1 * Y1 * (A1 + B1) + X1 * Z1 * (C1 + D1) + Y1 * Z1 * (G1 + F1));   // :322-347
                kap = -Ckappa[tr] * d2 * theta / sigma;                                    // :191
            }
            kappa_e[tr][o] = oc_max<FT>(FT(0), kap);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// Smagorinsky / SmagorinskyLilly eddy viscosity.  Replaces _compute_smagorinsky_viscosity!
// (src/TurbulenceClosures/turbulence_closure_implementations/Smagorinskys/smagorinsky.jl:92-108) with
// ΣᵢⱼΣᵢⱼᶜᶜᶜ (Smagorinskys/scale_invariant_operators.jl:10-13; velocity_tracer_gradients.jl:25-46,78) and, for the
// LillyCoefficient, square_smagorinsky_coefficient / stability (Smagorinskys/lilly_coefficient.jl:114-135) with ∂z_b
// (buoyancy_tracer.jl:16, seawater_buoyancy.jl:219-224).  One thread per cell, the AmdKernel's CTA shape; the off-diagonal
// strains are evaluated once at each of the four corners their interpolation needs.  κₑ[t] = νₑ / Pr[t] is stored per tracer
// (the reference divides the interpolated νₑ, smagorinsky.jl:154-156: the same value up to one rounding), so that the tendency
// kernels read eddy diffusivities exactly as they do for AMD.
// ---------------------------------------------------------------------------------------------------------
template <class FT, bool STR = false>
struct SmagorinskyKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 3;
    Geom<FT> g;
    const FT* u;
    const FT* v;
    const FT* w;
    FT* nu_e;
    int ntr;
    FT* kappa_e[8];
    FT rPr[8];           // 1 / Pr[tracer]
    FT cs2;              // C² (smagorinsky.jl:110; lilly_coefficient.jl:133)
    int lilly;           // 1: multiply by ς(N², Σ², Cb)
    FT Cb;
    int buoyancy;        // 0 none (∂z_b = 0, no_buoyancy.jl:9), 1 tracer b, 2 seawater linear
    const FT* bT;
    const FT* bS;
    FT grav, alpha, beta;
    FT df2;              // Δᶠ² = cbrt(Δx Δy Δz)², regular grids
    const FT* lv_df2;    // per level on a stretched grid (Model::build_z_tables), indexable like Geom::dzc

    // ∂z_b at ccf, level k (face between cells k-1 and k); o = index of cell k
    OC_HD FT dzb(int o, int k) const {
        const FT r = STR ? g.rdzf[k] : g.rd[2];
        if (buoyancy == 1) return (bT[o] - bT[o - g.sz]) * r;
        return grav * (alpha * ((bT[o] - bT[o - g.sz]) * r) - beta * ((bS[o] - bS[o - g.sz]) * r));
    }

    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        (void)nt;
        int i = b.x * 32 + (tid & 31), j = b.y * 8 + (tid >> 5), k = b.z;
        if (i >= g.N[0] || j >= g.N[1]) return;
        const int o = g.idx(i, j, k);
        const int sx = 1, sy = g.sy, sz = g.sz;
        const FT rdx = g.rd[0], rdy = g.rd[1];
        FT rzf[2];                                      // Δz⁻¹ at the ccf / fcf / cff level of the cell and of the level above
        rzf[0] = STR ? g.rdzf[k] : g.rd[2];
        rzf[1] = STR ? g.rdzf[k + 1] : g.rd[2];
        const FT rdzc = STR ? g.rdzc[k] : g.rd[2];
        const int cxy[4] = {o, o + sx, o + sy, o + sx + sy};
        const int cxz[4] = {o, o + sx, o + sz, o + sx + sz};
        const int cyz[4] = {o, o + sy, o + sz, o + sy + sz};
        FT q12[4], q13[4], q23[4];
        for (int n = 0; n < 4; ++n) {
            const int up = n >> 1;
            const int pxy = cxy[n], pxz = cxz[n], pyz = cyz[n];
            const FT S12 = FT(0.5) * ((u[pxy] - u[pxy - sy]) * rdy + (v[pxy] - v[pxy - sx]) * rdx);          // ffc: ½(∂y u + ∂x v)
            const FT S13 = FT(0.5) * ((u[pxz] - u[pxz - sz]) * rzf[up] + (w[pxz] - w[pxz - sx]) * rdx);      // fcf: ½(∂z u + ∂x w)
            const FT S23 = FT(0.5) * ((v[pyz] - v[pyz - sz]) * rzf[up] + (w[pyz] - w[pyz - sy]) * rdy);      // cff: ½(∂z v + ∂y w)
            q12[n] = S12 * S12; q13[n] = S13 * S13; q23[n] = S23 * S23;
        }
        // ℑxyᶜᶜᵃ = ℑyᶜ(ℑxᶜ ·) and so on: the lower dimension is the inner operator (interpolation_operators.jl:45-56)
        const FT I12 = FT(0.5) * (FT(0.5) * (q12[0] + q12[1]) + FT(0.5) * (q12[2] + q12[3]));
        const FT I13 = FT(0.5) * (FT(0.5) * (q13[0] + q13[1]) + FT(0.5) * (q13[2] + q13[3]));
        const FT I23 = FT(0.5) * (FT(0.5) * (q23[0] + q23[1]) + FT(0.5) * (q23[2] + q23[3]));
        const FT S11 = (u[o + sx] - u[o]) * rdx, S22 = (v[o + sy] - v[o]) * rdy, S33 = (w[o + sz] - w[o]) * rdzc;
        const FT tr = S11 * S11 + S22 * S22 + S33 * S33;
        const FT S2 = tr + FT(2) * I12 + FT(2) * I13 + FT(2) * I23;
        // νₑ = ς c² Δᶠ² √(2Σ²) with ς = √(1 − min(1, Cb N²⁺ / Σ²)) (0 when Σ² = 0)  ==  c² Δᶠ² √(2 max(0, Σ² − Cb N²⁺)):
        // the same value with one square root and no division (both forms lose the same digits when Σ² ≈ Cb N²⁺)
        FT arg = S2;
        if (lilly && buoyancy) {
            const FT N2 = FT(0.5) * (dzb(o, k) + dzb(o + sz, k + 1));                      // ℑzᵃᵃᶜ ∂z_b   lilly_coefficient.jl:130
            arg = oc_max<FT>(FT(0), S2 - Cb * oc_max<FT>(FT(0), N2));
        }
        const FT d2 = STR ? lv_df2[k] : df2;
        const FT nu = cs2 * d2 * oc_sqrt<FT>(FT(2) * arg);
        nu_e[o] = nu;
        for (int t = 0; t < ntr; ++t) kappa_e[t][o] = nu * rPr[t];                         // κₑ = νₑ / Pr; rPr = 1 / Pr formed on the host
    }
};

// ---------------------------------------------------------------------------------------------------------
// On-device step diagnostics: cell_advection_timescale (src/Advection/cell_advection_timescale.jl:13-34,
// the TimeStepWizard's reduction, src/Simulations/time_step_wizard.jl:101-115), max|u|, max|v|, max|w| (progress messages) and
// the NaNChecker's hasnan(u) (src/Diagnostics/nan_checker.jl) in ONE pass — instead of full-field device-to-host copies.
// All reduced quantities are non-negative doubles, so their bit patterns order like unsigned integers.
// out[0] = min timescale, out[1..3] = max |u|,|v|,|w|, out[4] = NaN flag.
// ---------------------------------------------------------------------------------------------------------
template <class FT>
struct DiagnosticsKernel {
    static constexpr int PHASES = 2;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr size_t SMEM = sizeof(double) * 5 * THREADS;
    Geom<FT> g;
    const FT* u;
    const FT* v;
    const FT* w;
    unsigned long long* out;
    OC_HD static unsigned long long bits(double x) { unsigned long long b; memcpy(&b, &x, 8); return b; }
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char* smem) const {
        double* sh = reinterpret_cast<double*>(smem);
        if (PHASE == 0) {
            double tmin = 1.0e300, mu = 0.0, mv = 0.0, mw = 0.0, nan = 0.0;
            const int j = b.y, k = b.z;
            for (int i = b.x * nt + tid; i < g.N[0]; i += nt * 1) {
                const int o = g.idx(i, j, k);
                const FT uu = u[o], vv = v[o], ww = w[o];
                const FT au = oc_abs<FT>(uu), av = oc_abs<FT>(vv), aw = oc_abs<FT>(ww);
                FT inv = (g.flat[0] ? FT(0) : au * g.rd[0]) + (g.flat[1] ? FT(0) : av * g.rd[1]) + (g.flat[2] ? FT(0) : aw * g.rdz_at(true, k));   // Δz⁻¹ᶜᶜᶠ
                double tau = (double)(FT(1) / inv);
                if (tau < tmin) tmin = tau;             // NaN compares false: a NaN cell does not enter the minimum, it raises the flag
                if ((double)au > mu) mu = au;
                if ((double)av > mv) mv = av;
                if ((double)aw > mw) mw = aw;
                if (uu != uu) nan = 1.0;
            }
            sh[tid] = tmin; sh[nt + tid] = mu; sh[2 * nt + tid] = mv; sh[3 * nt + tid] = mw; sh[4 * nt + tid] = nan;
        } else {
            if (tid != 0) return;
            double tmin = 1.0e300, mu = 0.0, mv = 0.0, mw = 0.0, nan = 0.0;
            for (int t = 0; t < nt; ++t) {
                if (sh[t] < tmin) tmin = sh[t];
                if (sh[nt + t] > mu) mu = sh[nt + t];
                if (sh[2 * nt + t] > mv) mv = sh[2 * nt + t];
                if (sh[3 * nt + t] > mw) mw = sh[3 * nt + t];
                if (sh[4 * nt + t] > nan) nan = sh[4 * nt + t];
            }
#if defined(__CUDA_ARCH__)
            atomicMin(out + 0, bits(tmin));
            atomicMax(out + 1, bits(mu));
            atomicMax(out + 2, bits(mv));
            atomicMax(out + 3, bits(mw));
            if (nan > 0.0) atomicMax(out + 4, 1ull);
#else
            if (bits(tmin) < out[0]) out[0] = bits(tmin);
            if (bits(mu) > out[1]) out[1] = bits(mu);
            if (bits(mv) > out[2]) out[2] = bits(mv);
            if (bits(mw) > out[3]) out[3] = bits(mw);
            if (nan > 0.0) out[4] = 1ull;
#endif
        }
    }
};

// maximum(abs, interior(field)) of ONE field, on the device (8-byte result): what cell_diffusion_timescale needs for the eddy-viscosity
// closures (maximum(νₑ), maximum(κₑ): src/TurbulenceClosures/turbulence_closure_diagnostics.jl:57-69) and progress messages need for
// tracers.  |x| ≥ 0, so the bit patterns of the doubles order like unsigned integers; a NaN anywhere raises out[1].
template <class FT>
struct FieldMaxAbsKernel {
    static constexpr int PHASES = 2;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr size_t SMEM = sizeof(double) * 2 * THREADS;
    Geom<FT> g;
    const FT* f;
    int nx;                // interior extent along x (N, +1 for a Face-located field in a Bounded x)
    unsigned long long* out;
    OC_HD static unsigned long long bits(double x) { unsigned long long b; memcpy(&b, &x, 8); return b; }
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char* smem) const {
        double* sh = reinterpret_cast<double*>(smem);
        if (PHASE == 0) {
            double mx = 0.0, nan = 0.0;
            for (int i = tid; i < nx; i += nt) {
                const FT x = f[g.idx(i, b.y, b.z)];
                const double a = (double)oc_abs<FT>(x);
                if (a > mx) mx = a;
                if (x != x) nan = 1.0;
            }
            sh[tid] = mx; sh[nt + tid] = nan;
        } else {
            if (tid != 0) return;
            double mx = 0.0, nan = 0.0;
            for (int t = 0; t < nt; ++t) {
                if (sh[t] > mx) mx = sh[t];
                if (sh[nt + t] > nan) nan = sh[nt + t];
            }
#if defined(__CUDA_ARCH__)
            atomicMax(out + 0, bits(mx));
            if (nan > 0.0) atomicMax(out + 1, 1ull);
#else
            if (bits(mx) > out[0]) out[0] = bits(mx);
            if (nan > 0.0) out[1] = 1ull;
#endif
        }
    }
};

}  // namespace oc
