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
        for (int k = Nz - 1; k >= 0; --k) {
            o -= g.sz;
            FT bk = b_at(o);
            FT bf = FT(0.5) * (bk + bup);              // z_dot_g_bᶜᶜᶠ(k+1) = ℑzᵃᵃᶠ b   g_dot_b.jl:4
            p = (k == Nz - 1) ? -bf * g.d[2] : p - bf * g.d[2];
            pHY[o] = p;
            bup = bk;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// AMD.  One thread per cell; gradients are re-derived from u, v, w (stencil radius 1 around the cell).
// The helper struct mirrors the reference's operator names so each term can be checked line by line.
// ---------------------------------------------------------------------------------------------------------
template <class FT>
struct AmdPoint {
    const Geom<FT>& g;
    const FT* u;
    const FT* v;
    const FT* w;
    FT rxy, ryx, rxz, rzx, ryz, rzy;   // Δᶠa/Δᶠb with Δᶠ = 2Δ   (:224-226)
    OC_HD AmdPoint(const Geom<FT>& g_, const FT* u_, const FT* v_, const FT* w_) : g(g_), u(u_), v(v_), w(w_) {
        FT fx = FT(2) * g.d[0], fy = FT(2) * g.d[1], fz = FT(2) * g.d[2];
        rxy = fx / fy; ryx = fy / fx; rxz = fx / fz; rzx = fz / fx; ryz = fy / fz; rzy = fz / fy;
    }
    // normalised gradients at their natural locations (velocity_tracer_gradients.jl:126-154); o = linear index
    OC_HD FT dxu(int o) const { return (u[o + 1] - u[o]) * g.rd[0]; }                       // ccc
    OC_HD FT dyv(int o) const { return (v[o + g.sy] - v[o]) * g.rd[1]; }                    // ccc
    OC_HD FT dzw(int o) const { return (w[o + g.sz] - w[o]) * g.rd[2]; }                    // ccc
    OC_HD FT dxv(int o) const { return rxy * ((v[o] - v[o - 1]) * g.rd[0]); }               // ffc
    OC_HD FT dyu(int o) const { return ryx * ((u[o] - u[o - g.sy]) * g.rd[1]); }            // ffc
    OC_HD FT dxw(int o) const { return rxz * ((w[o] - w[o - 1]) * g.rd[0]); }               // fcf
    OC_HD FT dzu(int o) const { return rzx * ((u[o] - u[o - g.sz]) * g.rd[2]); }            // fcf
    OC_HD FT dyw(int o) const { return ryz * ((w[o] - w[o - g.sy]) * g.rd[1]); }            // cff
    OC_HD FT dzv(int o) const { return rzy * ((v[o] - v[o - g.sz]) * g.rd[2]); }            // cff
    OC_HD FT S12(int o) const { return FT(0.5) * (dyu(o) + dxv(o)); }
    OC_HD FT S13(int o) const { return FT(0.5) * (dzu(o) + dxw(o)); }
    OC_HD FT S23(int o) const { return FT(0.5) * (dzv(o) + dyw(o)); }
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

template <class FT>
struct AmdKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
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

    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        const int o = g.idx(i, j, k);
        const int sx = 1, sy = g.sy, sz = g.sz;
        AmdPoint<FT> P(g, u, v, w);
        FT fx = FT(2) * g.d[0], fy = FT(2) * g.d[1], fz = FT(2) * g.d[2];
        FT delta2 = FT(3) / (FT(1) / (fx * fx) + FT(1) / (fy * fy) + FT(1) / (fz * fz));   // :166,190
        auto sq = [](FT x) { return x * x; };
        // interpolated products
        auto Ixy = [&](auto fn) { return interp2c<FT>(fn, o, sx, sy); };
        auto Ixz = [&](auto fn) { return interp2c<FT>(fn, o, sx, sz); };
        auto Iyz = [&](auto fn) { return interp2c<FT>(fn, o, sy, sz); };
        FT dxu = P.dxu(o), dyv = P.dyv(o), dzw = P.dzw(o);
        FT Ixy_dxv2 = Ixy([&](int q) { return sq(P.dxv(q)); });
        FT Ixy_dyu2 = Ixy([&](int q) { return sq(P.dyu(q)); });
        FT Ixz_dxw2 = Ixz([&](int q) { return sq(P.dxw(q)); });
        FT Ixz_dzu2 = Ixz([&](int q) { return sq(P.dzu(q)); });
        FT Iyz_dyw2 = Iyz([&](int q) { return sq(P.dyw(q)); });
        FT Iyz_dzv2 = Iyz([&](int q) { return sq(P.dzv(q)); });
        // norm_tr_∇uᶜᶜᶜ :285-306
        FT q = sq(dxu) + sq(dyv) + sq(dzw) + Ixy_dxv2 + Ixy_dyu2 + Ixz_dxw2 + Ixz_dzu2 + Iyz_dyw2 + Iyz_dzv2;
        FT nu = FT(0);
        if (q != FT(0)) {
            // norm_uᵢₐ_uⱼₐ_Σᵢⱼᶜᶜᶜ :240-279
            FT Ixy_dxv = Ixy([&](int p) { return P.dxv(p); });
            FT Ixy_dyu = Ixy([&](int p) { return P.dyu(p); });
            FT Ixz_dxw = Ixz([&](int p) { return P.dxw(p); });
            FT Ixz_dzu = Ixz([&](int p) { return P.dzu(p); });
            FT Iyz_dyw = Iyz([&](int p) { return P.dyw(p); });
            FT Iyz_dzv = Iyz([&](int p) { return P.dzv(p); });
            FT Ixy_S12 = Ixy([&](int p) { return P.S12(p); });
            FT Ixz_S13 = Ixz([&](int p) { return P.S13(p); });
            FT Iyz_S23 = Iyz([&](int p) { return P.S23(p); });
            FT t1 = dxu * sq(dxu) + dyv * Ixy_dxv2 + dzw * Ixz_dxw2
                  + FT(2) * dxu * Ixy([&](int p) { return P.dxv(p) * P.S12(p); })
                  + FT(2) * dxu * Ixz([&](int p) { return P.dxw(p) * P.S13(p); })
                  + FT(2) * Ixy_dxv * Ixz_dxw * Iyz_S23;
            FT t2 = dxu * Ixy_dyu2 + dyv * sq(dyv) + dzw * Iyz_dyw2
                  + FT(2) * dyv * Ixy([&](int p) { return P.dyu(p) * P.S12(p); })
                  + FT(2) * Ixy_dyu * Iyz_dyw * Ixz_S13
                  + FT(2) * dyv * Iyz([&](int p) { return P.dyw(p) * P.S23(p); });
            FT t3 = dxu * Ixz_dzu2 + dyv * Iyz_dzv2 + dzw * sq(dzw)
                  + FT(2) * Ixz_dzu * Iyz_dzv * Ixy_S12
                  + FT(2) * dzw * Ixz([&](int p) { return P.dzu(p) * P.S13(p); })
                  + FT(2) * dzw * Iyz([&](int p) { return P.dzv(p) * P.S23(p); });
            FT r = t1 + t2 + t3;
            FT Cb_zeta = FT(0) / fz;                                                       // Cb = nothing :281
            nu = -Cnu * delta2 * (r - Cb_zeta) / q;                                        // :168
        }
        nu_e[o] = oc_max<FT>(FT(0), nu);
        for (int t = 0; t < ntr; ++t) {
            const FT* cc = c[t];
            auto cx = [&](int p) { return fx * ((cc[p] - cc[p - sx]) * g.rd[0]); };       // norm_∂x_c at fcc
            auto cy = [&](int p) { return fy * ((cc[p] - cc[p - sy]) * g.rd[1]); };
            auto cz = [&](int p) { return fz * ((cc[p] - cc[p - sz]) * g.rd[2]); };
            FT Ix_cx2 = interp1c<FT>([&](int p) { return sq(cx(p)); }, o, sx);
            FT Iy_cy2 = interp1c<FT>([&](int p) { return sq(cy(p)); }, o, sy);
            FT Iz_cz2 = interp1c<FT>([&](int p) { return sq(cz(p)); }, o, sz);
            FT sigma = Ix_cx2 + Iy_cy2 + Iz_cz2;                                           // norm_θᵢ²ᶜᶜᶜ :349-351
            FT kap = FT(0);
            if (sigma != FT(0)) {
                FT Ix_cx = interp1c<FT>(cx, o, sx), Iy_cy = interp1c<FT>(cy, o, sy), Iz_cz = interp1c<FT>(cz, o, sz);
                FT Ixy_dxv = Ixy([&](int p) { return P.dxv(p); });
                FT Ixy_dyu = Ixy([&](int p) { return P.dyu(p); });
                FT Ixz_dxw = Ixz([&](int p) { return P.dxw(p); });
                FT Ixz_dzu = Ixz([&](int p) { return P.dzu(p); });
                FT Ixz_dyw = Ixz([&](int p) { return P.dyw(p); });                         // sic: ℑxzᶜᵃᶜ(norm_∂y_w) :336
                FT Iyz_dzv = Iyz([&](int p) { return P.dzv(p); });
                // norm_uᵢⱼ_cⱼ_cᵢᶜᶜᶜ :322-347
                FT a1 = dxu * Ix_cx2 + Ixy_dxv * Ix_cx * Iy_cy + Ixz_dxw * Ix_cx * Iz_cz;
                FT a2 = Ixy_dyu * Iy_cy * Ix_cx + dyv * Iy_cy2 + Ixz_dyw * Iy_cy * Iz_cz;
                FT a3 = Ixz_dzu * Iz_cz * Ix_cx + Iyz_dzv * Iz_cz * Iy_cy + dzw * Iz_cz2;
                FT theta = a1 + a2 + a3;
                kap = -Ckappa[t] * delta2 * theta / sigma;                                 // :191
            }
            kappa_e[t][o] = oc_max<FT>(FT(0), kap);
        }
    }
};

}  // namespace oc
