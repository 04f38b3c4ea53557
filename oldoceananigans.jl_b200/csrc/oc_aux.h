// oc_aux.h — auxiliary-field kernels of update_state!: hydrostatic pressure anomaly and AMD eddy
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
// STR = false: regular grid — the ratios with Δᶠz and the z-derivative metrics are constants (folded at compile time into the
// same arithmetic as before); STR = true: vertically stretched grid — they depend on the level of the evaluation point:
// every Δᶠz_{loc}(i,j,k′) is 2·Δzᶜᶜᶜ(k′) at the INDEX k′ of the point (:228-234), ∂z at fcf / cff uses Δz⁻¹ᶠ(k′), ∂z w at ccc Δz⁻¹ᶜ(k).
// `up` = 0: the point is on the cell's level k, 1: on the level above.
template <class FT, bool STR>
struct AmdPoint {
    const Geom<FT>& g;
    const FT* u;
    const FT* v;
    const FT* w;
    FT rxy, ryx;                       // Δᶠa/Δᶠb with Δᶠ = 2Δ   (:224-226)
    FT rxz_[2], rzx_[2], ryz_[2], rzy_[2], rdzf_[2], fz_[2], rdzc_;
    // the filter-width ratios are loop invariants with divisions: on a regular grid formed once on the host (set_consts), same FT
    // arithmetic; on a stretched grid formed here from the level tables
    OC_HD AmdPoint(const Geom<FT>& g_, const FT* u_, const FT* v_, const FT* w_, const FT* r, int k) : g(g_), u(u_), v(v_), w(w_) {
        rxy = r[0]; ryx = r[1];
        if (!STR) {
            rxz_[0] = r[2]; rzx_[0] = r[3]; ryz_[0] = r[4]; rzy_[0] = r[5]; rdzf_[0] = g.rd[2]; fz_[0] = FT(2) * g.d[2]; rdzc_ = g.rd[2];
        } else {
            const FT fx = FT(2) * g.d[0], fy = FT(2) * g.d[1];
            for (int n = 0; n < 2; ++n) {
                fz_[n] = FT(2) * g.dzc[k + n];
                rxz_[n] = fx / fz_[n]; rzx_[n] = fz_[n] / fx; ryz_[n] = fy / fz_[n]; rzy_[n] = fz_[n] / fy;
                rdzf_[n] = g.rdzf[k + n];
            }
            rdzc_ = g.rdzc[k];
        }
    }
    OC_HD FT rxz(int up) const { return rxz_[STR ? up : 0]; }
    OC_HD FT rzx(int up) const { return rzx_[STR ? up : 0]; }
    OC_HD FT ryz(int up) const { return ryz_[STR ? up : 0]; }
    OC_HD FT rzy(int up) const { return rzy_[STR ? up : 0]; }
    OC_HD FT rdzf(int up) const { return rdzf_[STR ? up : 0]; }
    OC_HD FT fz(int up) const { return fz_[STR ? up : 0]; }
    // normalised gradients at their natural locations (velocity_tracer_gradients.jl:126-154); o = linear index
    OC_HD FT dxu(int o) const { return (u[o + 1] - u[o]) * g.rd[0]; }                       // ccc
    OC_HD FT dyv(int o) const { return (v[o + g.sy] - v[o]) * g.rd[1]; }                    // ccc
    OC_HD FT dzw(int o) const { return (w[o + g.sz] - w[o]) * rdzc_; }                      // ccc
    OC_HD FT dxv(int o) const { return rxy * ((v[o] - v[o - 1]) * g.rd[0]); }               // ffc
    OC_HD FT dyu(int o) const { return ryx * ((u[o] - u[o - g.sy]) * g.rd[1]); }            // ffc
    OC_HD FT dxw(int o, int up) const { return rxz(up) * ((w[o] - w[o - 1]) * g.rd[0]); }       // fcf
    OC_HD FT dzu(int o, int up) const { return rzx(up) * ((u[o] - u[o - g.sz]) * rdzf(up)); }   // fcf
    OC_HD FT dyw(int o, int up) const { return ryz(up) * ((w[o] - w[o - g.sy]) * g.rd[1]); }    // cff
    OC_HD FT dzv(int o, int up) const { return rzy(up) * ((v[o] - v[o - g.sz]) * rdzf(up)); }   // cff
};

// ℑ of a functor F(o) from (Face,Face) in dims (d1<d2) to centre: ℑ_{d2}ᶜ(ℑ_{d1}ᶜ F)   interpolation_operators.jl:45-56
template <class FT, class F>
OC_HD FT interp2c(const F& fn, int o, int s1, int s2) {
    return FT(0.5) * (FT(0.5) * (fn(o) + fn(o + s1)) + FT(0.5) * (fn(o + s2) + fn(o + s1 + s2)));
}
template <class FT, class F>
OC_HD FT interp1c(const F& fn, int o, int s) {
    return FT(0.5) * (fn(o) + fn(o + s));
}

// 4-point interpolation of values already evaluated at the corners {o, o+s1, o+s2, o+s1+s2} — the arithmetic of interp2c
template <class FT>
OC_HD FT interp4(const FT* f) { return FT(0.5) * (FT(0.5) * (f[0] + f[1]) + FT(0.5) * (f[2] + f[3])); }

template <class FT, bool STR = false>
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
    void set_consts() {
        FT fx = FT(2) * g.d[0], fy = FT(2) * g.d[1], fz = FT(2) * g.d[2];
        ratios[0] = fx / fy; ratios[1] = fy / fx; ratios[2] = fx / fz; ratios[3] = fz / fx; ratios[4] = fy / fz; ratios[5] = fz / fy;
        delta2 = FT(3) / (FT(1) / (fx * fx) + FT(1) / (fy * fy) + FT(1) / (fz * fz));
    }

    // Every normalised gradient is evaluated ONCE at the four corners its interpolations need (24 evaluations per cell);
    // all 30 terms of AMD are then products of those — same operands, same order of operations as the reference's
    // nested ℑ(…) calls (anisotropic_minimum_dissipation.jl:240-351).
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        (void)nt;
        int i = b.x * 32 + (tid & 31), j = b.y * 8 + (tid >> 5), k = b.z;
        if (i >= g.N[0] || j >= g.N[1]) return;
        const int o = g.idx(i, j, k);
        const int sx = 1, sy = g.sy, sz = g.sz;
        AmdPoint<FT, STR> P(g, u, v, w, ratios, k);
        const FT fx = FT(2) * g.d[0], fy = FT(2) * g.d[1], fz = P.fz(0);
        const FT d2 = STR ? FT(3) / (FT(1) / (fx * fx) + FT(1) / (fy * fy) + FT(1) / (fz * fz)) : delta2;   // δ² (:166,190)
        auto sq = [](FT x) { return x * x; };
        const int cxy[4] = {o, o + sx, o + sy, o + sx + sy};
        const int cxz[4] = {o, o + sx, o + sz, o + sx + sz};
        const int cyz[4] = {o, o + sy, o + sz, o + sy + sz};
        FT dxv[4], dyu[4], dxw[4], dzu[4], dyw[4], dzv[4], t[4];
        for (int n = 0; n < 4; ++n) {
            dxv[n] = P.dxv(cxy[n]); dyu[n] = P.dyu(cxy[n]);
            dxw[n] = P.dxw(cxz[n], n >> 1); dzu[n] = P.dzu(cxz[n], n >> 1);     // corners 2, 3 are on the level above
            dyw[n] = P.dyw(cyz[n], n >> 1); dzv[n] = P.dzv(cyz[n], n >> 1);
        }
        const FT dxu = P.dxu(o), dyv = P.dyv(o), dzw = P.dzw(o);
        for (int n = 0; n < 4; ++n) t[n] = sq(dxv[n]);
        const FT Ixy_dxv2 = interp4<FT>(t);
        for (int n = 0; n < 4; ++n) t[n] = sq(dyu[n]);
        const FT Ixy_dyu2 = interp4<FT>(t);
        for (int n = 0; n < 4; ++n) t[n] = sq(dxw[n]);
        const FT Ixz_dxw2 = interp4<FT>(t);
        for (int n = 0; n < 4; ++n) t[n] = sq(dzu[n]);
        const FT Ixz_dzu2 = interp4<FT>(t);
        for (int n = 0; n < 4; ++n) t[n] = sq(dyw[n]);
        const FT Iyz_dyw2 = interp4<FT>(t);
        for (int n = 0; n < 4; ++n) t[n] = sq(dzv[n]);
        const FT Iyz_dzv2 = interp4<FT>(t);
        const FT Ixy_dxv = interp4<FT>(dxv), Ixy_dyu = interp4<FT>(dyu);
        const FT Ixz_dxw = interp4<FT>(dxw), Ixz_dzu = interp4<FT>(dzu);
        const FT Iyz_dyw = interp4<FT>(dyw), Iyz_dzv = interp4<FT>(dzv);
        // norm_tr_∇uᶜᶜᶜ :285-306
        FT q = sq(dxu) + sq(dyv) + sq(dzw) + Ixy_dxv2 + Ixy_dyu2 + Ixz_dxw2 + Ixz_dzu2 + Iyz_dyw2 + Iyz_dzv2;
        FT nu = FT(0);
        if (q != FT(0)) {
            // norm_uᵢₐ_uⱼₐ_Σᵢⱼᶜᶜᶜ :240-279
            FT S12[4], S13[4], S23[4];
            for (int n = 0; n < 4; ++n) {
                S12[n] = FT(0.5) * (dyu[n] + dxv[n]);
                S13[n] = FT(0.5) * (dzu[n] + dxw[n]);
                S23[n] = FT(0.5) * (dzv[n] + dyw[n]);
            }
            const FT Ixy_S12 = interp4<FT>(S12), Ixz_S13 = interp4<FT>(S13), Iyz_S23 = interp4<FT>(S23);
            for (int n = 0; n < 4; ++n) t[n] = dxv[n] * S12[n];
            const FT Ixy_dxvS12 = interp4<FT>(t);
            for (int n = 0; n < 4; ++n) t[n] = dxw[n] * S13[n];
            const FT Ixz_dxwS13 = interp4<FT>(t);
            for (int n = 0; n < 4; ++n) t[n] = dyu[n] * S12[n];
            const FT Ixy_dyuS12 = interp4<FT>(t);
            for (int n = 0; n < 4; ++n) t[n] = dyw[n] * S23[n];
            const FT Iyz_dywS23 = interp4<FT>(t);
            for (int n = 0; n < 4; ++n) t[n] = dzu[n] * S13[n];
            const FT Ixz_dzuS13 = interp4<FT>(t);
            for (int n = 0; n < 4; ++n) t[n] = dzv[n] * S23[n];
            const FT Iyz_dzvS23 = interp4<FT>(t);
            FT t1 = dxu * sq(dxu) + dyv * Ixy_dxv2 + dzw * Ixz_dxw2
                  + FT(2) * dxu * Ixy_dxvS12
                  + FT(2) * dxu * Ixz_dxwS13
                  + FT(2) * Ixy_dxv * Ixz_dxw * Iyz_S23;
            FT t2 = dxu * Ixy_dyu2 + dyv * sq(dyv) + dzw * Iyz_dyw2
                  + FT(2) * dyv * Ixy_dyuS12
                  + FT(2) * Ixy_dyu * Iyz_dyw * Ixz_S13
                  + FT(2) * dyv * Iyz_dywS23;
            FT t3 = dxu * Ixz_dzu2 + dyv * Iyz_dzv2 + dzw * sq(dzw)
                  + FT(2) * Ixz_dzu * Iyz_dzv * Ixy_S12
                  + FT(2) * dzw * Ixz_dzuS13
                  + FT(2) * dzw * Iyz_dzvS23;
            FT r = t1 + t2 + t3;
            FT Cb_zeta = FT(0) / fz;                                                       // Cb = nothing :281
            nu = -Cnu * d2 * (r - Cb_zeta) / q;                                            // :168
        }
        nu_e[o] = oc_max<FT>(FT(0), nu);
        if (ntr == 0) return;
        // ℑxzᶜᵃᶜ(norm_∂y_w) — sic, :336 — needs ∂y w at the xz corners
        for (int n = 0; n < 4; ++n) t[n] = P.dyw(cxz[n], n >> 1);
        const FT Ixz_dyw = interp4<FT>(t);
        for (int tr = 0; tr < ntr; ++tr) {
            const FT* cc = c[tr];
            auto cx = [&](int p) { return fx * ((cc[p] - cc[p - sx]) * g.rd[0]); };       // norm_∂x_c at fcc
            auto cy = [&](int p) { return fy * ((cc[p] - cc[p - sy]) * g.rd[1]); };
            auto cz = [&](int p, int up) { return P.fz(up) * ((cc[p] - cc[p - sz]) * P.rdzf(up)); };   // norm_∂z_c at ccf
            const FT cx0 = cx(o), cx1 = cx(o + sx), cy0 = cy(o), cy1 = cy(o + sy), cz0 = cz(o, 0), cz1 = cz(o + sz, 1);
            FT Ix_cx2 = FT(0.5) * (sq(cx0) + sq(cx1));
            FT Iy_cy2 = FT(0.5) * (sq(cy0) + sq(cy1));
            FT Iz_cz2 = FT(0.5) * (sq(cz0) + sq(cz1));
            FT sigma = Ix_cx2 + Iy_cy2 + Iz_cz2;                                           // norm_θᵢ²ᶜᶜᶜ :349-351
            FT kap = FT(0);
            if (sigma != FT(0)) {
                FT Ix_cx = FT(0.5) * (cx0 + cx1), Iy_cy = FT(0.5) * (cy0 + cy1), Iz_cz = FT(0.5) * (cz0 + cz1);
                // norm_uᵢⱼ_cⱼ_cᵢᶜᶜᶜ :322-347
                FT a1 = dxu * Ix_cx2 + Ixy_dxv * Ix_cx * Iy_cy + Ixz_dxw * Ix_cx * Iz_cz;
                FT a2 = Ixy_dyu * Iy_cy * Ix_cx + dyv * Iy_cy2 + Ixz_dyw * Iy_cy * Iz_cz;
                FT a3 = Ixz_dzu * Iz_cz * Ix_cx + Iyz_dzv * Iz_cz * Iy_cy + dzw * Iz_cz2;
                FT theta = a1 + a2 + a3;
                kap = -Ckappa[tr] * d2 * theta / sigma;                                    // :191
            }
            kappa_e[tr][o] = oc_max<FT>(FT(0), kap);
        }
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

}  // namespace oc
