// oc_tendency.h — fused tendency (+ RK3/AB2 substep) kernel for one prognostic field.
//
// Replaces compute_Gu!/Gv!/Gw!/Gc! (src/Models/NonhydrostaticModels/compute_nonhydrostatic_tendencies.jl
// :138-163 -> nonhydrostatic_tendency_kernel_functions.jl:70-298), compute_flux_bc_tendencies!
// (compute_flux_bcs.jl:116-163), rk3_substep_field! (src/TimeSteppers/runge_kutta_3.jl:212-226),
// ab2_step_field! (quasi_adams_bashforth_2.jl:162-175) and _cache_field_tendencies!
// (store_tendencies.jl:6-9, by pointer swap) with ONE launch per field.
//
// Design: a CTA owns a TX×TY×TZ tile of cells.  Phase 0 evaluates every face flux of
// the tile ONCE (advective + closure flux, (TX+1)·TY·TZ + TX·(TY+1)·TZ + TX·TY·(TZ+1) faces) into
// shared memory — the reference evaluates both faces of every cell in every thread, i.e. each face
// twice.  Phase 1 forms the flux divergence per cell, adds Coriolis / hydrostatic-pressure / buoyancy /
// flux-BC terms, writes Gⁿ and the substepped state U* (state is double-buffered: read Uⁿ, write U*).
#pragma once
#include "oc_advection.h"

namespace oc {

enum FieldKind { KIND_U = 0, KIND_V = 1, KIND_W = 2, KIND_C = 3 };
enum SubstepMode { STEP_NONE = 0, STEP_RK3_FIRST = 1, STEP_RK3 = 2, STEP_AB2 = 3 };

template <class FT>
struct FluxBC {          // compute_flux_bcs.jl: G[1] += J·A/V ; G[N] -= J·A/V
    int on[6];
    FT val[6];
};

template <class FT>
struct TendencyArgs {
    Geom<FT> g;
    AdvCoef<FT> C;
    const FT* U[3];      // current state u, v, w
    const FT* c;         // tracer being stepped (KIND_C) else nullptr
    const FT* pHY;       // hydrostatic pressure anomaly or nullptr
    const FT* bT;        // buoyancy sources for the w-equation when pHY′ is absent (nullptr otherwise)
    const FT* bS;
    const FT* nu_e;      // AMD eddy viscosity (ccc) or nullptr
    const FT* kappa_e;   // AMD eddy diffusivity of this tracer or nullptr
    const FT* Gm;        // G⁻ of this field (read) or nullptr
    FT* Gn;              // Gⁿ of this field (written)
    const FT* Ucur;      // this field, current state (same as U[kind] or c)
    FT* Unew;            // this field, next state (STEP_* != NONE)
    int has_scalar;      // ScalarDiffusivity present
    FT nu, kappa;        // its ν and this tracer's κ
    int buoyancy;        // 0 none, 1 tracer b (bT), 2 seawater linear (bT = T, bS = S)
    FT grav, alpha, beta;
    int has_coriolis;    // oc_coriolis (1 FPlane; 2, 3: general tile kernel only, parameters in CoriolisExt)
    FT f;
    FluxBC<FT> fbc;      // flux boundary conditions of this field
    int add_flux_bcs;
    FT cor_beta, cor_y0; // BetaPlane in the z-marching kernel: f = f₀ + β ynode (beta_plane.jl:56-72); y0 = south face of this rank's first row
    int adv_dir[3];      // ADV_MIXED kernels: scheme code of the fluxes through the faces normal to d (adapt_advection_order.jl:18-96)
    int mode;            // SubstepMode
    FT dt, ca, cb;       // RK3_FIRST: U + (dt·γ)·G with ca = dt·γ ; RK3: U + dt(ca·G + cb·G⁻) ; AB2: ca = 1.5+χ, cb = 0.5+χ
    int ab2_euler;
};

// BetaPlane / ConstantCartesianCoriolis parameters.  Kept OUT of TendencyArgs so that the parameter block — and with it the
// generated code — of the z-marching kernel of the measured configurations (oc_march.h) stays exactly what was profiled.
template <class FT>
struct CoriolisExt {
    FT beta, y0;         // BetaPlane: f = f₀ + β ynode; y0 = y of the south face of this rank's first row
    FT cf[3];            // ConstantCartesianCoriolis fx, fy, fz; NonTraditionalBetaPlane: cf[1] = fy, cf[2] = fz
    FT gamma, R, z0;     // NonTraditionalBetaPlane γ, R; z0 = z of the bottom face (regular z)
    const FT* zc;        // stretched z: znode tables at Center / Face levels, indexable like Geom::dzc (nullptr on regular grids)
    const FT* zf;
    int tilted;          // BuoyancyForce(…; gravity_unit_vector): x_dot_g_bᶠᶜᶜ = ĝ_x ℑxᶠ b, y_dot_g_bᶜᶠᶜ = ĝ_y ℑyᶠ b   g_dot_b.jl:1-2
    FT gh[3];            // ĝ = −gravity_unit_vector   buoyancy_force.jl:52-54
    int tb_kind;         // buoyancy model of the tilted terms: 1 tracer b, 2 seawater linear (own copies: TendencyArgs::buoyancy
    const FT* tbT;       // is 0 whenever pHY′ exists)
    const FT* tbS;
};

template <class FT, int ADV, int KIND, int TX_, int TY_, int TZ_>
struct TendencyKernel {
    static constexpr int PHASES = 2;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int TX = TX_, TY = TY_, TZ = TZ_;
    static constexpr int NFX = (TX + 1) * TY * TZ, NFY = TX * (TY + 1) * TZ, NFZ = TX * TY * (TZ + 1);
    static constexpr size_t SMEM = sizeof(FT) * (size_t)(NFX + NFY + NFZ);
    static constexpr int COMP = KIND == KIND_C ? -1 : KIND;

    TendencyArgs<FT> a;
    CoriolisExt<FT> cor;

    // ---- closure fluxes -------------------------------------------------------------------------------
    // ν at the location that is Face in dims (d1,d2) (d1 < d2): ℑ_{d2}ᶠ(ℑ_{d1}ᶠ νₑ)  interpolation_operators.jl:45-56
    OC_HD FT nu_ff(int o, int d1, int d2) const {
        const Geom<FT>& g = a.g;
        int s1 = g.st(d1), s2 = g.st(d2);
        const FT* n = a.nu_e + o;
        return FT(0.5) * (FT(0.5) * (n[-s1 - s2] + n[-s2]) + FT(0.5) * (n[-s1] + n[0]));
    }

    // viscous flux A_d · τ_{comp,d} at flux index o   (closure_kernel_operators.jl:22-41;
    // abstract_scalar_diffusivity_closure.jl:189-204; velocity_tracer_gradients.jl:25-42)
    // k = level of the flux point (stretched grids: Δz⁻¹ᶜ for Σ33 at ccc, Δz⁻¹ᶠ for ∂z u, ∂z v at fcf / cff; the area of an
    // x- or y-face carries the Δz of the stepped field's own z-location)
    OC_HD FT viscous_flux(int o, int d, int k) const {
        const Geom<FT>& g = a.g;
        FT sig;
        if (d == COMP) {
            const FT* u = a.U[d] + o;
            sig = (u[g.st(d)] - u[0]) * (d == 2 ? g.rdz_at(false, k) : g.rd[d]);  // Σ_dd at ccc
        } else {
            int lo = d < COMP ? d : COMP, hi = d < COMP ? COMP : d;
            const FT* ul = a.U[lo] + o;
            const FT* uh = a.U[hi] + o;
            FT dl = (ul[0] - ul[-g.st(hi)]) * (hi == 2 ? g.rdz_at(true, k) : g.rd[hi]);   // ∂_hi u_lo
            FT dh = (uh[0] - uh[-g.st(lo)]) * g.rd[lo];                           // ∂_lo u_hi
            sig = FT(0.5) * (dl + dh);
        }
        const FT A = g.area_at(d, COMP == 2, k);
        FT flux = FT(0);
        if (a.has_scalar) flux = A * (FT(-2) * (a.nu * sig));
        if (a.nu_e) {
            FT nu;
            if (d == COMP) nu = a.nu_e[o];
            else nu = nu_ff(o, d < COMP ? d : COMP, d < COMP ? COMP : d);
            FT f2 = A * (FT(-2) * (nu * sig));
            flux = a.has_scalar ? flux + f2 : f2;
        }
        return flux;
    }

    // diffusive tracer flux A_d · q_d at face index o   (:43-48, :240-242, κ at faces :327-330)
    OC_HD FT diffusive_flux(int o, int d, int k) const {
        const Geom<FT>& g = a.g;
        int s = g.st(d);
        const FT* c = a.c + o;
        FT grad = (c[0] - c[-s]) * (d == 2 ? g.rdz_at(true, k) : g.rd[d]);
        const FT A = g.area_at(d, false, k);
        FT flux = FT(0);
        if (a.has_scalar) flux = A * (-(a.kappa * grad));
        if (a.kappa_e) {
            FT kap = FT(0.5) * (a.kappa_e[o - s] + a.kappa_e[o]);
            FT f2 = A * (-(kap * grad));
            flux = a.has_scalar ? flux + f2 : f2;
        }
        return flux;
    }

    // ---- advective fluxes -----------------------------------------------------------------------------
    // Flux of this field through the faces normal to d, at flux index (i,j,k): for momentum component
    // COMP the flux is centre-type in d when d == COMP (located at ccc) and face-type otherwise.
    // FluxFormAdvection(x, y, z): every flux direction has its own scheme — `_advective_momentum_flux_Uu(…, advection::FluxFormAdvection, …)
    // = _advective_momentum_flux_Uu(…, advection.x, …)` etc. (src/Advection/flux_form_advection.jl): the scheme of direction d does BOTH
    // interpolations of the flux through the d-faces
    OC_HD FT advective_flux(int i, int j, int k, int d) const {
        if constexpr (ADV != ADV_MIXED) return advective_flux_t<ADV>(i, j, k, d);
        else {
            switch (a.adv_dir[d]) {
                case ADV_WENO5: return advective_flux_t<ADV_WENO5>(i, j, k, d);
                case ADV_CENTERED4: return advective_flux_t<ADV_CENTERED4>(i, j, k, d);
                case ADV_UPWIND3: return advective_flux_t<ADV_UPWIND3>(i, j, k, d);
                case ADV_UPWIND5: return advective_flux_t<ADV_UPWIND5>(i, j, k, d);
                case ADV_WENO3: return advective_flux_t<ADV_WENO3>(i, j, k, d);
                case ADV_UPWIND1: return advective_flux_t<ADV_UPWIND1>(i, j, k, d);
                case ADV_WENO7: return advective_flux_t<ADV_WENO7>(i, j, k, d);
                case ADV_WENO9: return advective_flux_t<ADV_WENO9>(i, j, k, d);
                case ADV_NONE: return FT(0);
                default: return advective_flux_t<ADV_CENTERED2>(i, j, k, d);
            }
        }
    }
    template <int ADVD>
    OC_HD FT advective_flux_t(int i, int j, int k, int d) const {
        const Geom<FT>& g = a.g;
        if (g.flat[d]) return FT(0);                                               // flat_advective_fluxes.jl:13-29
        int o = g.idx(i, j, k);
        int sd = g.st(d);
        int id = d == 0 ? i : (d == 1 ? j : k);
        // Centered: the area at the flux point (centered_advective_fluxes.jl:15-33); upwind schemes: the area of the advecting
        // velocity's own point, inside the interpolation (upwind_biased_advective_fluxes.jl:23-121) — these differ only
        // for the x / y fluxes of w on a stretched grid
        FT A = g.area_at(d, COMP == 2 && adv_is_centered(ADVD), k);
        if (ADVD == ADV_NONE) return FT(0);                                        // advection = nothing
        constexpr bool CEN = adv_is_centered(ADVD);
        if (KIND == KIND_C) {
            FT u = a.U[d][o];
            const FT* c = a.c + o;
            OrderWindow w = order_window(g.wlo[d] != 0, g.whi[d] != 0, false, g.N[d]);
            if (ADVD == ADV_CENTERED2) {
                return (A * u) * (FT(0.5) * c[-sd] + FT(0.5) * c[0]);              // centered_advective_fluxes.jl:31-33
            } else if (CEN) {
                return (A * u) * symmetric_any<ADVD, FT>(a.C, c, sd, FT(1), id, w);
            } else {
                FT cr = biased_any<ADVD, FT>(a.C, c, sd, u > FT(0), id, w);         // upwind_biased_advective_fluxes.jl:99-121
                return A * u * cr;
            }
        } else {
            const FT* psi = a.U[COMP < 0 ? 0 : COMP] + o;
            const FT* adv = a.U[d] + o;
            if (d == COMP) {
                // centre-type: evaluate the face-type stencils at face id+1
                OrderWindow w = order_window(g.wlo[d] != 0, g.whi[d] != 0, true, g.N[d]);
                if (ADVD == ADV_CENTERED2) {
                    FT ut = FT(0.5) * adv[0] + FT(0.5) * adv[sd];
                    FT pt = FT(0.5) * psi[0] + FT(0.5) * psi[sd];
                    return A * ut * pt;                                            // centered_advective_fluxes.jl:15,22,27
                } else if (CEN) {
                    FT ut = symmetric_any<ADVD, FT>(a.C, adv + sd, sd, FT(1), id + 1, w);
                    FT pt = symmetric_any<ADVD, FT>(a.C, psi + sd, sd, FT(1), id + 1, w);
                    return A * ut * pt;
                } else {
                    FT ut = symmetric_any<ADVD, FT>(a.C, adv + sd, sd, A, id + 1, w);
                    FT pr = biased_any<ADVD, FT>(a.C, psi + sd, sd, ut > FT(0), id + 1, w);
                    return ut * pr;                                                // upwind_biased_advective_fluxes.jl:23-29
                }
            } else {
                int cc = COMP < 0 ? 0 : COMP;
                int sc = g.st(cc);
                int ic = cc == 0 ? i : (cc == 1 ? j : k);
                OrderWindow wc = order_window(g.wlo[cc] != 0, g.whi[cc] != 0, false, g.N[cc]);
                OrderWindow wd = order_window(g.wlo[d] != 0, g.whi[d] != 0, false, g.N[d]);
                if (ADVD == ADV_CENTERED2) {
                    FT ut = g.flat[cc] ? adv[0] : (FT(0.5) * adv[-sc] + FT(0.5) * adv[0]);
                    FT pt = FT(0.5) * psi[-sd] + FT(0.5) * psi[0];
                    return A * ut * pt;                                            // :16-26
                } else if (CEN) {
                    FT ut = g.flat[cc] ? adv[0] : symmetric_any<ADVD, FT>(a.C, adv, sc, FT(1), ic, wc);
                    FT pt = symmetric_any<ADVD, FT>(a.C, psi, sd, FT(1), id, wd);
                    return A * ut * pt;
                } else {
                    FT ut;
                    if (g.flat[cc]) ut = A * adv[0];
                    else if (cc == 2 && g.stretched()) ut = symmetric_any_z<ADVD, FT>(a.C, adv, sc, g.d[d == 0 ? 1 : 0], g.dzc + k, ic, wc);
                    else ut = symmetric_any<ADVD, FT>(a.C, adv, sc, A, ic, wc);
                    FT pr = biased_any<ADVD, FT>(a.C, psi, sd, ut > FT(0), id, wd);
                    return ut * pr;                                                // :31-93
                }
            }
        }
    }

    OC_HD FT total_flux(int i, int j, int k, int d) const {
        FT F = advective_flux(i, j, k, d);
        if (a.has_scalar || a.nu_e || a.kappa_e) {
            if (!a.g.flat[d]) {
                int o = a.g.idx(i, j, k);
                F = F + (KIND == KIND_C ? diffusive_flux(o, d, k) : viscous_flux(o, d, k));
            }
        }
        return F;
    }

    // buoyancy_perturbationᶜᶜᶜ  (linear_equation_of_state.jl:72-74, buoyancy_tracer.jl:12)
    OC_HD FT buoyancy_at(int o) const {
        if (a.buoyancy == 1) return a.bT[o];
        return a.grav * (a.alpha * a.bT[o] - a.beta * a.bS[o]);
    }

    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char* smem) const {
        const Geom<FT>& g = a.g;
        FT* fx = reinterpret_cast<FT*>(smem);
        FT* fy = fx + NFX;
        FT* fz = fy + NFY;
        const int i0 = b.x * TX, j0 = b.y * TY, k0 = b.z * TZ;
        const int shx = COMP == 0 ? -1 : 0, shy = COMP == 1 ? -1 : 0, shz = COMP == 2 ? -1 : 0;
        if (PHASE == 0) {
            for (int n = tid; n < NFX; n += nt) {
                int s = n % (TX + 1), jj = (n / (TX + 1)) % TY, kk = n / ((TX + 1) * TY);
                int j = j0 + jj, k = k0 + kk;
                FT F = FT(0);
                if (j < g.N[1] && k < g.N[2] && i0 + s <= g.N[0]) F = total_flux(i0 + s + shx, j, k, 0);
                fx[n] = F;
            }
            for (int n = tid; n < NFY; n += nt) {
                int ii = n % TX, s = (n / TX) % (TY + 1), kk = n / (TX * (TY + 1));
                int i = i0 + ii, k = k0 + kk;
                FT F = FT(0);
                if (i < g.N[0] && k < g.N[2] && j0 + s <= g.N[1]) F = total_flux(i, j0 + s + shy, k, 1);
                fy[n] = F;
            }
            for (int n = tid; n < NFZ; n += nt) {
                int ii = n % TX, jj = (n / TX) % TY, s = n / (TX * TY);
                int i = i0 + ii, j = j0 + jj;
                FT F = FT(0);
                if (i < g.N[0] && j < g.N[1] && k0 + s <= g.N[2]) F = total_flux(i, j, k0 + s + shz, 2);
                fz[n] = F;
            }
        } else {
            for (int n = tid; n < TX * TY * TZ; n += nt) {
                int ii = n % TX, jj = (n / TX) % TY, kk = n / (TX * TY);
                int i = i0 + ii, j = j0 + jj, k = k0 + kk;
                if (i >= g.N[0] || j >= g.N[1] || k >= g.N[2]) continue;
                int o = g.idx(i, j, k);
                FT u0 = a.Ucur[o];
                // exclude_periphery: wall faces of a Face-located field in a Bounded dimension are not stepped
                // (src/Utils/kernel_launching.jl:145-146)
                bool wall = false;
                if (COMP >= 0) {
                    int ic = COMP == 0 ? i : (COMP == 1 ? j : k);
                    wall = g.wlo[COMP] && ic == 0 && g.N[COMP] > 1;
                }
                if (wall) {
                    if (a.mode != STEP_NONE) a.Unew[o] = u0;
                    continue;
                }
                FT dFx = fx[(kk * TY + jj) * (TX + 1) + ii + 1] - fx[(kk * TY + jj) * (TX + 1) + ii];
                FT dFy = fy[(kk * (TY + 1) + jj + 1) * TX + ii] - fy[(kk * (TY + 1) + jj) * TX + ii];
                FT dFz = fz[((kk + 1) * TY + jj) * TX + ii] - fz[(kk * TY + jj) * TX + ii];
                FT G = -(g.rV_at(COMP == 2, k) * (dFx + dFy + dFz));
                if (KIND == KIND_W && a.buoyancy && !a.pHY && !g.flat[2]) {
                    // maybe_z_dot_g_bᶜᶜᶠ: only without the hydrostatic split (nonhydrostatic_tendency_kernel_functions.jl:168-170)
                    G = G + FT(0.5) * (buoyancy_at(o - g.sz) + buoyancy_at(o));
                }
                // (along a Flat dimension too: ℑ is the identity there, the term is ĝ·b — all planes of the Flat direction hold the same b)
                if ((KIND == KIND_U || KIND == KIND_V) && cor.tilted) {
                    const int s = KIND == KIND_U ? 1 : g.sy;
                    FT b0, b1;
                    if (cor.tb_kind == 1) { b0 = cor.tbT[o - s]; b1 = cor.tbT[o]; }
                    else {
                        b0 = a.grav * (a.alpha * cor.tbT[o - s] - a.beta * cor.tbS[o - s]);
                        b1 = a.grav * (a.alpha * cor.tbT[o] - a.beta * cor.tbS[o]);
                    }
                    G = G + cor.gh[KIND == KIND_U ? 0 : 1] * (FT(0.5) * (b0 + b1));
                }
                if (KIND != KIND_C && a.has_coriolis == 3) {
                    // ConstantCartesianCoriolis (constant_cartesian_coriolis.jl:70-81): x: ℑxᶠ(fy ℑzᶜ w − fz ℑyᶜ v), y: ℑyᶠ(fz ℑxᶜ u − fx ℑzᶜ w),
                    // z: ℑzᶠ(fx ℑyᶜ v − fy ℑxᶜ u); no active-node weighting.  Flat dimensions are stored periodic with N = 1: ℑ = identity.
                    const FT* u = a.U[0]; const FT* v = a.U[1]; const FT* w = a.U[2];
                    const FT h = FT(0.5);
                    const int sC = KIND == KIND_U ? 1 : (KIND == KIND_V ? g.sy : g.sz);
                    FT acc[2];
                    for (int n = 0; n < 2; ++n) {
                        const int p = o - (1 - n) * sC;                    // the two ccc points either side of the velocity point
                        FT hv;
                        if (KIND == KIND_U) hv = cor.cf[1] * (h * (w[p] + w[p + g.sz])) - cor.cf[2] * (h * (v[p] + v[p + g.sy]));
                        else if (KIND == KIND_V) hv = cor.cf[2] * (h * (u[p] + u[p + 1])) - cor.cf[0] * (h * (w[p] + w[p + g.sz]));
                        else hv = cor.cf[0] * (h * (v[p] + v[p + g.sy])) - cor.cf[1] * (h * (u[p] + u[p + 1]));
                        acc[n] = hv;
                    }
                    G = G - h * (acc[0] + acc[1]);
                }
                if (KIND != KIND_C && a.has_coriolis == 4) {
                    // NonTraditionalBetaPlane (non_traditional_beta_plane.jl:79-96): 2Ωʸ = fy (1 − z/R) + γ y, 2Ωᶻ = fz (1 + 2z/R) + β y
                    //   x: ℑxᶠ(2Ωʸ ℑzᶜ w − 2Ωᶻ ℑyᶜ v) with y, z at ccc;  y: 2Ωᶻ(cfc) ℑxyᶜᶠ u;  z: −2Ωʸ(ccf) ℑxzᶜᶠ u;  no active-node weighting
                    const FT* u = a.U[0]; const FT* v = a.U[1]; const FT* w = a.U[2];
                    const FT h = FT(0.5);
                    const FT y = cor.y0 + (FT(j) + (KIND == KIND_V ? FT(0) : FT(0.5))) * g.d[1];
                    const FT z = cor.zc ? (KIND == KIND_W ? cor.zf[k] : cor.zc[k])
                                        : cor.z0 + (FT(k) + (KIND == KIND_W ? FT(0) : FT(0.5))) * g.d[2];
                    const FT Oy = cor.cf[1] * (FT(1) - z / cor.R) + cor.gamma * y;
                    const FT Oz = cor.cf[2] * (FT(1) + FT(2) * z / cor.R) + cor.beta * y;
                    if (KIND == KIND_U) {
                        FT acc[2];
                        for (int n = 0; n < 2; ++n) {
                            const int p = o - (1 - n);
                            acc[n] = Oy * (h * (w[p] + w[p + g.sz])) - Oz * (h * (v[p] + v[p + g.sy]));
                        }
                        G = G - h * (acc[0] + acc[1]);
                    } else if (KIND == KIND_V) {
                        G = G - Oz * (h * (h * (u[o - g.sy] + u[o - g.sy + 1]) + h * (u[o] + u[o + 1])));
                    } else {
                        G = G + Oy * (h * (h * (u[o - g.sz] + u[o - g.sz + 1]) + h * (u[o] + u[o + 1])));
                    }
                }
                if ((KIND == KIND_U || KIND == KIND_V) && (a.has_coriolis == 1 || a.has_coriolis == 2)) {
                    // FPlane: x_f_cross_U = -f·ℑxyᶠᶜᶜ(v)/active ; y_f_cross_U = +f·ℑxyᶜᶠᶜ(u)/active   f_plane.jl:50-52
                    // BetaPlane: the same with f = f₀ + β·ynode(fcc | cfc)   beta_plane.jl:56-72
                    FT fj = a.f;
                    if (a.has_coriolis == 2) fj = a.f + cor.beta * (cor.y0 + (FT(j) + (KIND == KIND_U ? FT(0.5) : FT(0))) * g.d[1]);
                    FT num, cnt;
                    if (KIND == KIND_U) {
                        const FT* v = a.U[1] + o;
                        num = FT(0.5) * (FT(0.5) * (v[-1] + v[0]) + FT(0.5) * (v[g.sy - 1] + v[g.sy]));
                        // v-nodes (i-1,j),(i,j),(i-1,j+1),(i,j+1): active iff not on/outside a wall (inactive_node.jl:152-158)
                        int ax0 = !(g.wlo[0] && (i - 1 < 0)), ax1 = 1;
                        int ay0 = !(g.wlo[1] && (j < 1)), ay1 = !(g.whi[1] && (j + 1 > g.N[1] - 1));
                        cnt = FT(0.5) * (FT(0.5) * FT(ax0 * ay0 + ax1 * ay0) + FT(0.5) * FT(ax0 * ay1 + ax1 * ay1));
                        FT val = cnt == FT(0) ? FT(0) : num / cnt;
                        G = G - (-fj * val);
                    } else {
                        const FT* u = a.U[0] + o;
                        num = FT(0.5) * (FT(0.5) * (u[-g.sy] + u[-g.sy + 1]) + FT(0.5) * (u[0] + u[1]));
                        // u-nodes (i,j-1),(i+1,j-1),(i,j),(i+1,j)
                        int ax0 = !(g.wlo[0] && (i < 1)), ax1 = !(g.whi[0] && (i + 1 > g.N[0] - 1));
                        int ay0 = !(g.wlo[1] && (j - 1 < 0)), ay1 = 1;
                        cnt = FT(0.5) * (FT(0.5) * FT(ax0 * ay0 + ax1 * ay0) + FT(0.5) * FT(ax0 * ay1 + ax1 * ay1));
                        FT val = cnt == FT(0) ? FT(0) : num / cnt;
                        G = G - (fj * val);
                    }
                }
                if ((KIND == KIND_U || KIND == KIND_V) && a.pHY) {
                    // hydrostatic_pressure_gradient_x/y = ∂xᶠᶜᶜ / ∂yᶜᶠᶜ pHY′
                    int s = KIND == KIND_U ? 1 : g.sy;
                    if (!g.flat[KIND]) G = G - (a.pHY[o] - a.pHY[o - s]) * g.rd[KIND];
                }
                if (a.add_flux_bcs) {
                    // compute_flux_bcs.jl:126-163 — G[1] += J·A/V ; G[N] -= J·A/V
                    int ijk[3] = {i, j, k};
                    for (int d = 0; d < 3; ++d) {
                        const FT A = g.area_at(d, COMP == 2, k), V = g.vol_at(COMP == 2, k);
                        if (a.fbc.on[2 * d] && ijk[d] == 0) G = G + a.fbc.val[2 * d] * A / V;
                        if (a.fbc.on[2 * d + 1] && ijk[d] == g.N[d] - 1) G = G - a.fbc.val[2 * d + 1] * A / V;
                    }
                }
                a.Gn[o] = G;
                if (a.mode == STEP_RK3_FIRST) {
                    a.Unew[o] = u0 + a.ca * G;
                } else if (a.mode == STEP_RK3) {
                    a.Unew[o] = u0 + a.dt * (a.ca * G + a.cb * a.Gm[o]);
                } else if (a.mode == STEP_AB2) {
                    FT Gu = a.ab2_euler ? a.ca * G : a.ca * G - a.cb * a.Gm[o];
                    a.Unew[o] = u0 + a.dt * Gu;
                }
            }
        }
    }
};

// Stand-alone substep (the staged API: rk3_substep! / ab2_step! on already computed Gⁿ, G⁻), in place.
template <class FT>
struct SubstepKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* U;
    const FT* Gn;
    const FT* Gm;
    int comp;          // 0,1,2 velocity component or -1
    int mode;
    FT dt, ca, cb;
    int ab2_euler;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        if (comp >= 0) {
            int ic = comp == 0 ? i : (comp == 1 ? j : k);
            if (g.wlo[comp] && ic == 0 && g.N[comp] > 1) return;
        }
        int o = g.idx(i, j, k);
        FT G = Gn[o];
        if (mode == STEP_RK3_FIRST) U[o] = U[o] + ca * G;
        else if (mode == STEP_RK3) U[o] = U[o] + dt * (ca * G + cb * Gm[o]);
        else if (mode == STEP_AB2) {
            FT Gu = ab2_euler ? ca * G : ca * G - cb * Gm[o];
            U[o] = U[o] + dt * Gu;
        }
    }
};

// Adds flux-BC contributions to an existing Gⁿ (staged API compute_flux_bc_tendencies!)
template <class FT>
struct FluxBCKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* Gn;
    FluxBC<FT> fbc;
    int zface;         // 1: the field is Face-located in z (w)
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        int ijk[3] = {i, j, k};
        int o = g.idx(i, j, k);
        FT G = Gn[o];
        bool touched = false;
        for (int d = 0; d < 3; ++d) {
            const FT A = g.area_at(d, zface != 0, k), V = g.vol_at(zface != 0, k);
            if (fbc.on[2 * d] && ijk[d] == 0) { G = G + fbc.val[2 * d] * A / V; touched = true; }
            if (fbc.on[2 * d + 1] && ijk[d] == g.N[d] - 1) { G = G - fbc.val[2 * d + 1] * A / V; touched = true; }
        }
        if (touched) Gn[o] = G;
    }
};

// Array-valued flux boundary condition on ONE side of one field: FluxBoundaryCondition(J::AbstractArray) — getbc(bc, i, j, …) = J[i, j]
// (src/BoundaryConditions/boundary_condition.jl getbc for arrays; compute_flux_bcs.jl:116-163: G[1] += J·A/V, G[N] −= J·A/V).
// Runs after the fused tendency + substep launch of that field, on the boundary plane only: Gⁿ receives the flux term and, because the
// substep is linear in Gⁿ, U* receives coef · (flux term) (coef = Δt γ for RK3, Δt (3/2 + χ) for AB2; 0 for the staged evaluation).
// The scalar-valued sides stay fused in the tendency kernels.
template <class FT>
struct FluxArrayKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* Gn;
    FT* Unew;          // nullptr: tendencies only
    FT coef;
    const FT* J;       // n1 × n2 values, first tangential dimension fastest
    int d;             // normal dimension, side 0 (low: +) / 1 (high: −)
    int side;
    int comp;          // velocity component of the field or −1
    int zface;         // field is Face-located in z (w): area / volume metrics
    int n1, n2;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int i1 = b.x * nt + tid, i2 = b.y;
        if (i1 >= n1 || i2 >= n2) return;
        int ijk[3];
        const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        ijk[d] = side == 0 ? 0 : g.N[d] - 1;
        ijk[t1] = i1; ijk[t2] = i2;
        if (comp >= 0 && g.wlo[comp] && ijk[comp] == 0 && g.N[comp] > 1) return;          // wall faces are not stepped
        const int k = ijk[2];
        const FT A = g.area_at(d, zface != 0, k), V = g.vol_at(zface != 0, k);
        const FT dG = J[i1 + (size_t)n1 * i2] * A / V;
        const int o = g.idx(ijk[0], ijk[1], ijk[2]);
        if (side == 0) { Gn[o] = Gn[o] + dG; if (Unew) Unew[o] = Unew[o] + coef * dG; }
        else { Gn[o] = Gn[o] - dG; if (Unew) Unew[o] = Unew[o] - coef * dG; }
    }
};

// Copy interior (cache_previous_tendencies! for the staged API)
template <class FT>
struct CopyKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* dst;
    const FT* src;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        int i = b.x * nt + tid, j = b.y, k = b.z;
        if (i >= g.N[0]) return;
        int o = g.idx(i, j, k);
        dst[o] = src[o];
    }
};

}  // namespace oc
