// oc_model_impl.h — Model<FT>: the member definitions.  Included by oc_model_f64.cu and oc_model_f32.cu, which instantiate one float type each
// (two translation units, compiled in parallel: the kernels are templates, a single unit took ~4 minutes of nvcc).
#pragma once
#include "oc_model.h"

#include <algorithm>
#include <map>
#include <type_traits>

namespace oc {

// ---------------------------------------------------------------------------------------------------------
// coefficients (host).  stencil_coefficients on a uniform grid, exact rationals rounded to FT, last one
// 1 - sum(others)   (src/Advection/reconstruction_coefficients.jl:49-64)
// ---------------------------------------------------------------------------------------------------------
struct Rat {
    long long n, d;
};
static long long gcdll(long long a, long long b) { a = a < 0 ? -a : a; b = b < 0 ? -b : b; while (b) { long long t = a % b; a = b; b = t; } return a ? a : 1; }
static Rat rnorm(Rat r) { long long g = gcdll(r.n, r.d); r.n /= g; r.d /= g; if (r.d < 0) { r.n = -r.n; r.d = -r.d; } return r; }
static Rat radd(Rat a, Rat b) { return rnorm(Rat{a.n * b.d + b.n * a.d, a.d * b.d}); }

template <class FT>
static void stencil_coefficients(int r, int order, FT* out) {
    std::vector<Rat> c(order);
    for (int j = 0; j < order; ++j) {
        Rat acc{0, 1};
        for (int m = j + 1; m <= order; ++m) {
            long long num = 0;
            for (int l = 0; l <= order; ++l) {
                if (l == m) continue;
                long long p = 1;
                for (int q = 0; q <= order; ++q)
                    if (q != m && q != l) p *= (r - q + 1);
                num += p;
            }
            long long den = 1;
            for (int l = 0; l <= order; ++l)
                if (l != m) den *= (m - l);
            acc = radd(acc, rnorm(Rat{num, den}));
        }
        c[j] = acc;
    }
    FT s = FT(0);
    for (int j = 0; j < order - 1; ++j) {
        out[j] = (FT)((long double)c[j].n / (long double)c[j].d);
        s = s + out[j];
    }
    out[order - 1] = FT(1) - s;
}

template <class FT>
static AdvCoef<FT> make_coefficients() {
    AdvCoef<FT> C;
    FT c4[4];
    stencil_coefficients<FT>(1, 4, c4);                    // Centered(order=4): buffer 2 -> r = buffer-1
    for (int idx = 1; idx <= 4; ++idx) C.c4[idx - 1] = c4[4 - idx];   // calc_reconstruction_stencil: coeff[order-idx+1]
    for (int r = 0; r < 3; ++r) stencil_coefficients<FT>(r, 3, C.w5p[r]);
    for (int r = 0; r < 2; ++r) stencil_coefficients<FT>(r, 2, C.w3p[r]);
    C.w5c[0] = (FT)(3.0L / 10.0L); C.w5c[1] = (FT)(3.0L / 5.0L); C.w5c[2] = (FT)(1.0L / 10.0L);
    C.w3c[0] = (FT)(2.0L / 3.0L); C.w3c[1] = (FT)(1.0L / 3.0L);
    C.eps = (FT)1e-8f;
    // UpwindBiased: coeff_left = stencil_coefficients(r = buffer-2), coeff_right = (r = buffer-1), order = 2 buffer - 1
    // (reconstruction_coefficients.jl:88-89).  calc_reconstruction_stencil (:122-152) gives the idx-th stencil point the coefficient
    // coeff[order-idx+1]: Left idx = 1 … order is ψ[i-B] … ψ[i+B-2], Right is ψ[i-B+1] … ψ[i+B-1].  Upwind-ordered (q0 = the point
    // farthest upwind): Left q_n = ψ[i-B+n] -> coeff_left[order-1-n]; Right q_n = ψ[i+B-1-n] -> coeff_right[n].
    {
        FT l5[5], r5[5], l3[3], r3[3];
        stencil_coefficients<FT>(1, 5, l5); stencil_coefficients<FT>(2, 5, r5);
        stencil_coefficients<FT>(0, 3, l3); stencil_coefficients<FT>(1, 3, r3);
        for (int n = 0; n < 5; ++n) { C.u5l[n] = l5[4 - n]; C.u5r[n] = r5[n]; }
        for (int n = 0; n < 3; ++n) { C.u3l[n] = l3[2 - n]; C.u3r[n] = r3[n]; }
    }
    C.hi = nullptr;
    return C;
}

// the table behind AdvCoef::hi (HiOrderTab): WENO{4}, WENO{5} and their Centered(6), Centered(8) advecting-velocity schemes
// (weno_interpolants.jl:81-90 C★, :117-118 coeff_p, :175-185 smoothness coefficients; centered_reconstruction.jl / reconstruction_coefficients.jl:87)
template <class FT>
static std::vector<FT> make_hi_order_table() {
    std::vector<FT> t(HiOrderTab::SIZE, FT(0));
    static const double s7[4][10] = {{2.107, -9.402, 7.042, -1.854, 11.003, -17.246, 4.642, 7.043, -3.882, 0.547},
                                     {0.547, -2.522, 1.922, -0.494, 3.443, -5.966, 1.602, 2.843, -1.642, 0.267},
                                     {0.267, -1.642, 1.602, -0.494, 2.843, -5.966, 1.922, 3.443, -2.522, 0.547},
                                     {0.547, -3.882, 4.642, -1.854, 7.043, -17.246, 7.042, 11.003, -9.402, 2.107}};
    static const double s9[5][15] = {
        {1.07918, -6.49501, 7.58823, -4.11487, 0.86329, 10.20563, -24.62076, 13.58458, -2.88007, 15.21393, -17.04396, 3.64863, 4.82963, -2.08501, 0.22658},
        {0.22658, -1.40251, 1.65153, -0.88297, 0.18079, 2.42723, -6.11976, 3.37018, -0.70237, 4.06293, -4.64976, 0.99213, 1.38563, -0.60871, 0.06908},
        {0.06908, -0.51001, 0.67923, -0.38947, 0.08209, 1.04963, -2.99076, 1.79098, -0.38947, 2.31153, -2.99076, 0.67923, 1.04963, -0.51001, 0.06908},
        {0.06908, -0.60871, 0.99213, -0.70237, 0.18079, 1.38563, -4.64976, 3.37018, -0.88297, 4.06293, -6.11976, 1.65153, 2.42723, -1.40251, 0.22658},
        {0.22658, -2.08501, 3.64863, -2.88007, 0.86329, 4.82963, -17.04396, 13.58458, -4.11487, 15.21393, -24.62076, 7.58823, 10.20563, -6.49501, 1.07918}};
    for (int r = 0; r < 4; ++r) for (int n = 0; n < 10; ++n) t[HiOrderTab::S7 + r * 10 + n] = (FT)s7[r][n];
    for (int r = 0; r < 5; ++r) for (int n = 0; n < 15; ++n) t[HiOrderTab::S9 + r * 15 + n] = (FT)s9[r][n];
    for (int r = 0; r < 4; ++r) stencil_coefficients<FT>(r, 4, &t[HiOrderTab::P7 + r * 4]);
    for (int r = 0; r < 5; ++r) stencil_coefficients<FT>(r, 5, &t[HiOrderTab::P9 + r * 5]);
    // FT(4//35) …: the rational rounded once to Float64 (IEEE division of two exact integers), then to FT
    const double c7[4] = {4.0 / 35.0, 18.0 / 35.0, 12.0 / 35.0, 1.0 / 35.0};
    const double c9[5] = {5.0 / 126.0, 20.0 / 63.0, 10.0 / 21.0, 10.0 / 63.0, 1.0 / 126.0};
    for (int r = 0; r < 4; ++r) t[HiOrderTab::C7 + r] = (FT)c7[r];
    for (int r = 0; r < 5; ++r) t[HiOrderTab::C9 + r] = (FT)c9[r];
    // Centered(order): coefficients of the symmetric stencil ψ[i-B] … ψ[i+B-1] = stencil_coefficients(r = B-1, order) reversed
    // (calc_reconstruction_stencil gives the idx-th point the coefficient coeff[order-idx+1], like Centered(4) above)
    FT c6[6], c8[8];
    stencil_coefficients<FT>(2, 6, c6);
    stencil_coefficients<FT>(3, 8, c8);
    for (int idx = 1; idx <= 6; ++idx) t[HiOrderTab::CEN6 + idx - 1] = c6[6 - idx];
    for (int idx = 1; idx <= 8; ++idx) t[HiOrderTab::CEN8 + idx - 1] = c8[8 - idx];
    return t;
}

// ---------------------------------------------------------------------------------------------------------
template <class FT>
Model<FT>::Model(const oc_config& c) : cfg_(c) {
    if (c.abi_version != OC_ABI_VERSION) throw Error(OC_ERR_INVALID, "oc_config.abi_version mismatch");
    if (c.n_tracers < 0 || c.n_tracers > OC_MAX_TRACERS) throw Error(OC_ERR_INVALID, "n_tracers out of range");
    auto known_scheme = [](int a) { return (a >= OC_CENTERED2 && a <= OC_ADVECTION_NONE) || a == OC_WENO7 || a == OC_WENO9; };
    if (!known_scheme(c.advection))
        throw Error(OC_ERR_UNSUPPORTED, "advection scheme: Centered(order=2|4), UpwindBiased(order=1|3|5), WENO(order=3|5|7|9) or nothing");
    if (c.timestepper != OC_RK3 && c.timestepper != OC_AB2) throw Error(OC_ERR_UNSUPPORTED, "timestepper: only RungeKutta3 and QuasiAdamsBashforth2");
    F_ = 3 + c.n_tracers;
    stretched_ = c.z_stretched != 0;
    if (stretched_) {
        // only a Bounded direction can be the tridiagonal one (fourier_tridiagonal_poisson_solver.jl:86-90)
        if (c.topology[2] != OC_BOUNDED) throw Error(OC_ERR_INVALID, "a stretched z needs the Bounded topology (FourierTridiagonalPoissonSolver)");
        if (!c.z_faces) throw Error(OC_ERR_INVALID, "z_stretched without z_faces");
        if (c.N[2] < 2) throw Error(OC_ERR_UNSUPPORTED, "stretched z with fewer than 2 levels");
        for (int k = 0; k < c.N[2]; ++k)
            if (!((FT)c.z_faces[k + 1] > (FT)c.z_faces[k])) throw Error(OC_ERR_INVALID, "The elements of z must be increasing!");
    }
    // required_halo_size of the scheme (its buffer) — per direction for FluxFormAdvection (adapt_advection_order.jl:18-96)
    auto buffer_of = [](int adv) { return adv == OC_WENO9 ? 5 : adv == OC_WENO7 ? 4 : (adv == OC_WENO5 || adv == OC_UPWIND5) ? 3 : (adv == OC_CENTERED4 || adv == OC_UPWIND3 || adv == OC_WENO3) ? 2 : 1; };
    if (c.has_advection_dir)
        for (int d = 0; d < 3; ++d)
            if (!known_scheme(c.advection_dir[d])) throw Error(OC_ERR_INVALID, "advection_dir: unknown advection scheme code");
    for (int d = 0; d < 3; ++d) {
        int need = buffer_of(c.has_advection_dir ? c.advection_dir[d] : c.advection);
        if (c.has_amd || c.smagorinsky) need = std::max(need, 2);      // AbstractScalarDiffusivity{…, 2}: anisotropic_minimum_dissipation.jl, smagorinsky.jl:31
        const int t = c.topology[d];
        if (t != OC_PERIODIC && t != OC_BOUNDED && t != OC_FLAT) throw Error(OC_ERR_INVALID, "bad topology");
        if (c.N[d] < 1) throw Error(OC_ERR_INVALID, "grid size must be >= 1");
        if (t == OC_FLAT) {
            if (c.N[d] != 1 || c.H[d] != 0) throw Error(OC_ERR_INVALID, "Flat dimensions must have N = 1 and H = 0");
        } else {
            if (c.H[d] < need) throw Error(OC_ERR_INVALID, "halo too small for the advection scheme / closure (inflate_grid_halo_size)");
            if (c.H[d] > 8) throw Error(OC_ERR_UNSUPPORTED, "halo larger than 8");
            if (c.N[d] < c.H[d]) throw Error(OC_ERR_INVALID, "halo must be <= size in every non-Flat dimension (validate_halo, input_validation.jl:86-92); "
                                                                   "lower the advection scheme there with has_advection_dir (adapt_advection_order)");
            if (!(stretched_ && d == 2) && !(c.delta[d] > 0)) throw Error(OC_ERR_INVALID, "grid spacing must be positive");
        }
        Hcfg_[d] = c.H[d];
        g_.N[d] = c.N[d];
        g_.H[d] = t == OC_FLAT ? 3 : std::max(3, c.H[d]);   // internal halo >= 3 (TMA boxes); the API halo is Hcfg_
        g_.bounded[d] = t == OC_BOUNDED;
        g_.wlo[d] = g_.whi[d] = g_.bounded[d];
        if (d < 2 && c.dist_nranks > 1 && g_.bounded[d]) {       // a Bounded partitioned dimension: only the outer ranks have a wall (distributed_grids.jl:75-126)
            const int Rx = std::max(1, c.dist_ranks_x), Ry = c.dist_nranks / Rx;
            const int r = d == 0 ? c.dist_rank / std::max(1, Ry) : c.dist_rank % std::max(1, Ry), n = d == 0 ? Rx : Ry;
            if (n > 1) { g_.wlo[d] = r == 0; g_.whi[d] = r == n - 1; }
        }
        g_.flat[d] = t == OC_FLAT;
        g_.d[d] = t == OC_FLAT ? FT(1) : (FT)c.delta[d];
        g_.rd[d] = FT(1) / g_.d[d];
    }
    if (c.has_advection_dir) {
        // the scheme of flux direction d interpolates the advecting velocity along every other direction with Centered(4) (fifth-order
        // schemes, Centered(4)) or Centered(2): where the adapted halo is smaller the reference reads outside the halo — refused
        for (int d = 0; d < 3; ++d) {
            const int a = c.advection_dir[d];
            const int deep = a == OC_WENO9 ? 4 : a == OC_WENO7 ? 3 : (a == OC_WENO5 || a == OC_UPWIND5 || a == OC_CENTERED4) ? 2 : 1;
            for (int e = 0; e < 3; ++e)
                if (e != d && c.topology[e] != OC_FLAT && c.topology[d] != OC_FLAT && a != OC_ADVECTION_NONE && c.H[e] < deep)
                    throw Error(OC_ERR_UNSUPPORTED, "advection_dir: a scheme interpolates velocities two points deep along a direction whose halo is 1 (the reference reads outside the halo there)");
        }
    }
    if (c.smagorinsky) {
        if (c.smagorinsky != 1 && c.smagorinsky != 2) throw Error(OC_ERR_UNSUPPORTED, "Smagorinsky: constant coefficient (1) or LillyCoefficient (2); DynamicCoefficient is not implemented");
        if (c.has_amd) throw Error(OC_ERR_UNSUPPORTED, "AnisotropicMinimumDissipation and Smagorinsky in one closure tuple");
        for (int t = 0; t < c.n_tracers; ++t)
            if (!(c.smag_Pr[t] > 0)) throw Error(OC_ERR_INVALID, "Smagorinsky: the turbulent Prandtl number of every tracer must be positive");
    }
    if (c.has_coriolis < OC_CORIOLIS_NONE || c.has_coriolis > OC_CORIOLIS_NONTRADITIONAL_BETAPLANE) throw Error(OC_ERR_UNSUPPORTED, "Coriolis: FPlane, BetaPlane, ConstantCartesianCoriolis or NonTraditionalBetaPlane");
    if (c.tilted_gravity) {
        const double* v = c.gravity_unit_vector;
        const double nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (!(std::fabs(nrm - 1.0) < 1e-8)) throw Error(OC_ERR_INVALID, "gravity_unit_vector must be unitary (validate_unit_vector)");
    }
    if (c.has_coriolis == OC_CORIOLIS_BETAPLANE && g_.flat[1]) throw Error(OC_ERR_UNSUPPORTED, "BetaPlane on a grid with a Flat y");
    if (c.has_coriolis == OC_CORIOLIS_NONTRADITIONAL_BETAPLANE) {
        if (g_.flat[1] || g_.flat[2]) throw Error(OC_ERR_UNSUPPORTED, "NonTraditionalBetaPlane on a grid with a Flat y or z");
        if (!(c.coriolis_radius != 0.0)) throw Error(OC_ERR_INVALID, "NonTraditionalBetaPlane: radius must be non-zero");
    }
    if (c.buoyancy == OC_BUOYANCY_SEAWATER_LINEAR && (c.tracer_T < 0 || c.tracer_S < 0 || c.tracer_T >= c.n_tracers || c.tracer_S >= c.n_tracers))
        throw Error(OC_ERR_INVALID, "SeawaterBuoyancy needs tracers T and S");
    if (c.buoyancy == OC_BUOYANCY_TRACER && (c.tracer_b < 0 || c.tracer_b >= c.n_tracers)) throw Error(OC_ERR_INVALID, "BuoyancyTracer needs tracer b");
    if (stretched_) {     // the constant z metrics do not exist: poison them so that a kernel that forgot the level tables shows up
        g_.d[2] = std::numeric_limits<FT>::quiet_NaN();
        g_.rd[2] = g_.d[2];
    }
    g_.A[0] = g_.d[1] * g_.d[2];
    g_.A[1] = g_.d[0] * g_.d[2];
    g_.A[2] = g_.d[0] * g_.d[1];
    g_.V = g_.A[2] * g_.d[2];
    g_.rV = FT(1) / g_.V;
    const int pad = 128 / (int)sizeof(FT);
    g_.sy = ((pad + g_.N[0] + g_.H[0] + 1 + pad - 1) / pad) * pad;
    const int rows = g_.N[1] + 2 * g_.H[1] + 1, planes = g_.N[2] + 2 * g_.H[2] + 1;
    g_.sz = g_.sy * rows;
    field_elems_ = (size_t)g_.sz * planes;
    if (field_elems_ >= ((size_t)1 << 31)) throw Error(OC_ERR_UNSUPPORTED, "field larger than 2^31 elements");
    origin_off_ = pad + (long long)g_.H[1] * g_.sy + (long long)g_.H[2] * g_.sz;
    xpad_ = pad;
    // the z-marching TMA kernel: no Flat dimension, and the two schemes of the BASELINE configurations; the other schemes of the
    // family run in the general tile kernel (oc_tendency.h)
    // (UpwindBiased(5): the triply periodic constant-viscosity variant only — the measurement configuration of §6)
    {
        const bool any_bounded = c.topology[0] == OC_BOUNDED || c.topology[1] == OC_BOUNDED || c.topology[2] == OC_BOUNDED;
        // BetaPlane / ConstantCartesianCoriolis run in the general tile kernel: the z-marching kernel of the
        // measured BASELINE configurations keeps exactly the code (and register counts) it was profiled with
        march_ok_ = !(g_.flat[0] || g_.flat[1] || g_.flat[2]) && (c.has_coriolis <= OC_CORIOLIS_FPLANE || c.has_coriolis == OC_CORIOLIS_BETAPLANE) && !c.tilted_gravity && !c.has_advection_dir &&
                    (c.advection == OC_CENTERED2 || c.advection == OC_WENO5 || (c.advection == OC_UPWIND5 && !any_bounded && !c.has_amd && !c.smagorinsky && !c.array_diffusivity));
    }
    {
        static const char* uvw_env = getenv("OC_UVW");
        const bool any_bounded = c.topology[0] == OC_BOUNDED || c.topology[1] == OC_BOUNDED || c.topology[2] == OC_BOUNDED;
        uvw_ok_ = march_ok_ && c.advection == OC_CENTERED2 && !any_bounded && !c.has_amd && !c.smagorinsky && !c.array_diffusivity && c.dist_nranks <= 1 &&
                  (uvw_env ? atoi(uvw_env) != 0 : true);
    }
    g_.dzc = g_.dzf = g_.rdzc = g_.rdzf = g_.rVc = g_.rVf = nullptr;
    C_ = make_coefficients<FT>();
    hi_adv_ = c.advection == OC_WENO7 || c.advection == OC_WENO9;
    if (c.has_advection_dir) for (int d = 0; d < 3; ++d) hi_adv_ = hi_adv_ || c.advection_dir[d] == OC_WENO7 || c.advection_dir[d] == OC_WENO9;
    {   // the compile-time table of oc_march.h must be the very same numbers
        using K = AdvConst<FT>;
        const FT tab[] = {K::p00, K::p01, K::p02, K::p10, K::p11, K::p12, K::p20, K::p21, K::p22, K::c50, K::c51, K::c52,
                          K::q00, K::q01, K::q10, K::q11, K::c30, K::c31, K::c40, K::c41, K::c42, K::c43, K::eps};
        const FT ref[] = {C_.w5p[0][0], C_.w5p[0][1], C_.w5p[0][2], C_.w5p[1][0], C_.w5p[1][1], C_.w5p[1][2], C_.w5p[2][0], C_.w5p[2][1],
                          C_.w5p[2][2], C_.w5c[0], C_.w5c[1], C_.w5c[2], C_.w3p[0][0], C_.w3p[0][1], C_.w3p[1][0], C_.w3p[1][1],
                          C_.w3c[0], C_.w3c[1], C_.c4[0], C_.c4[1], C_.c4[2], C_.c4[3], C_.eps};
        for (size_t n = 0; n < sizeof(tab) / sizeof(tab[0]); ++n)
            if (tab[n] != ref[n]) throw Error(OC_ERR_STATE, "internal: compile-time reconstruction coefficient " + std::to_string(n) + " differs from the reference derivation");
    }
    gamma_[0] = (FT)(8.0L / 15.0L); gamma_[1] = (FT)(5.0L / 12.0L); gamma_[2] = (FT)(3.0L / 4.0L);   // runge_kutta_3.jl:69-78
    zeta_[0] = FT(0); zeta_[1] = (FT)(-17.0L / 60.0L); zeta_[2] = (FT)(-5.0L / 12.0L);
#ifndef OC_HOSTSIM
    cuda_check(cudaSetDevice(c.device), "cudaSetDevice");
    device = c.device;
    {
        int lo = 0, hi = 0;
        cuda_check(cudaDeviceGetStreamPriorityRange(&lo, &hi), "cudaDeviceGetStreamPriorityRange");
        cuda_check(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, hi), "cudaStreamCreate");     // solve + communication first
        cuda_check(cudaStreamCreateWithPriority(&stream2_, cudaStreamNonBlocking, lo), "cudaStreamCreate");
    }
    {
        cudaEvent_t a, b;
        cuda_check(cudaEventCreateWithFlags(&a, cudaEventDisableTiming), "cudaEventCreate");
        cuda_check(cudaEventCreateWithFlags(&b, cudaEventDisableTiming), "cudaEventCreate");
        ev_fork_ = a; ev_join_ = b;
    }
    launch_stream_ = stream_;
#endif
    if (hi_adv_) {
        std::vector<FT> tab = make_hi_order_table<FT>();
        hi_tab_ = (FT*)dev_alloc(sizeof(FT) * tab.size());
        dev_upload(hi_tab_, tab.data(), sizeof(FT) * tab.size(), stream_);
#ifndef OC_HOSTSIM
        cuda_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");     // `tab` is a local staging buffer
#endif
        C_.hi = hi_tab_;
        device_bytes += (int64_t)(sizeof(FT) * tab.size());
    }
    const int locs[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
    for (int f = 0; f < F_; ++f) {
        const int* loc = locs[f < 3 ? f : 3];
        FieldRec a = alloc_field(loc), b = alloc_field(loc);
        resolve_bcs(a, c.bcs[f]);
        resolve_bcs(b, c.bcs[f]);
        state_.push_back(a);
        next_.push_back(b);
        FieldRec gn = alloc_field(loc), gm = alloc_field(loc);
        resolve_bcs(gn, nullptr);
        resolve_bcs(gm, nullptr);
        Gn_.push_back(gn);
        Gm_.push_back(gm);
    }
    pNHS_ = alloc_field(locs[3]);
    resolve_bcs(pNHS_, nullptr);
    has_pHY_ = c.buoyancy != OC_BUOYANCY_NONE;          // nonhydrostatic_model.jl:147-153
    if (has_pHY_) { pHY_ = alloc_field(locs[3]); resolve_bcs(pHY_, nullptr); }
    has_amd_ = c.has_amd != 0;
    has_smag_ = c.smagorinsky != 0;
    array_diff_ = c.array_diffusivity != 0;
    if (array_diff_ && (has_amd_ || has_smag_)) throw Error(OC_ERR_UNSUPPORTED, "array-valued ScalarDiffusivity together with AnisotropicMinimumDissipation / Smagorinsky (they share the diffusivity fields)");
    has_eddy_ = has_amd_ || has_smag_ || array_diff_;
    if (has_eddy_) {
        nu_e_ = alloc_field(locs[3]);
        resolve_bcs(nu_e_, nullptr);
        for (int t = 0; t < c.n_tracers; ++t) { kappa_e_.push_back(alloc_field(locs[3])); resolve_bcs(kappa_e_.back(), nullptr); }
    }
    // slab decomposition in y (Distributed(arch; partition = Partition(1, R)))
    dist_ = c.dist_nranks > 1;
    if (dist_) {
        Rx_ = std::max(1, c.dist_ranks_x);
        if (c.dist_nranks % Rx_ != 0) throw Error(OC_ERR_INVALID, "dist_nranks is not a multiple of dist_ranks_x");
        R_ = c.dist_nranks / Rx_;
        if (c.dist_rank < 0 || c.dist_rank >= c.dist_nranks) throw Error(OC_ERR_INVALID, "dist_rank out of range");
        rx_ = c.dist_rank / R_; rank_ = c.dist_rank % R_;
        // two-dimensional models: a Flat dimension cannot be partitioned — x-y models take slabs in x, Partition(R, 1); y-z models slabs
        // in y, Partition(1, R); a Flat y would leave nothing to transpose x against (the reference's solver has the same constraint,
        // Ny % Rx = 0, distributed_fft_based_poisson_solver.jl:211-229)
        if (c.topology[1] == OC_FLAT || (c.topology[0] == OC_FLAT && Rx_ > 1) || (c.topology[2] == OC_FLAT && R_ > 1))
            throw Error(OC_ERR_UNSUPPORTED, "distributed models with a Flat dimension: (x, y) models as Partition(R, 1), (y, z) models as Partition(1, R)");
        if (g_.N[2] % R_ != 0) throw Error(OC_ERR_INVALID, "distributed FFT: Nz must be divisible by the number of ranks along y (distributed_fft_based_poisson_solver.jl:211-229)");
        if (((long long)g_.N[1] * R_) % Rx_ != 0) throw Error(OC_ERR_INVALID, "distributed FFT: Ny must be divisible by the number of ranks along x (distributed_fft_based_poisson_solver.jl:211-229)");
        if ((R_ > 1 && g_.N[1] < g_.H[1]) || (Rx_ > 1 && g_.N[0] < g_.H[0])) throw Error(OC_ERR_INVALID, "distributed models: local size smaller than the halo");
    }
    if (stretched_) build_z_tables(c.z_faces);
    cfg_.z_faces = nullptr;                              // borrowed host pointer: not kept
    // pressure solver
    // slab decomposition: y is transformed in the transposed layout, where it is whole — the local buffer carries no y permutation
    // (pencils: neither is x — a full complex local buffer, no permutation in x or y)
    const int local_bounded[3] = {Rx_ > 1 ? 0 : g_.bounded[0], dist_ ? 0 : g_.bounded[1], g_.bounded[2]};
    std::string err = fft_.init(g_.N, local_bounded, stream_, !dist_, stretched_, Rx_ > 1);
    if (!err.empty()) throw Error(OC_ERR_CUDA, err);
    fftbuf_ = (FT*)dev_alloc(fft_.buffer_bytes);
    device_bytes += (int64_t)fft_.buffer_bytes + (int64_t)fft_.work_bytes;
    if (dist_) {
#ifndef OC_HOSTSIM
        cuda_check(cudaStreamCreateWithFlags(&stream3_, cudaStreamNonBlocking), "cudaStreamCreate");
#endif
        err = dfft_.init(g_.N[0], g_.N[1], g_.N[2], R_, g_.bounded[0] != 0, stream_, stream3_, Rx_, stretched_);
#ifndef OC_HOSTSIM
        for (int c = 0; c < dfft_.C; ++c) {
            cudaEvent_t a, b;
            cuda_check(cudaEventCreateWithFlags(&a, cudaEventDisableTiming), "cudaEventCreate");
            cuda_check(cudaEventCreateWithFlags(&b, cudaEventDisableTiming), "cudaEventCreate");
            ev_a2a_.push_back(a); ev_mid_.push_back(b);
        }
#endif
        if (!err.empty()) throw Error(OC_ERR_CUDA, err);
#ifndef OC_HOSTSIM
        if (Rx_ > 1 || stretched_) dfft_.set_y_stream(stream_);      // the pencil and the tridiagonal solves are one stream: no sub-chunk pipelining on stream3_ (yet)
#endif
        distT_ = (FT*)dev_alloc(fft_.buffer_bytes);
        diststage_ = (FT*)dev_alloc(fft_.buffer_bytes);
        const int planes = (int)(field_elems_ / (size_t)g_.sz);
        halo_buf_elems_ = (size_t)2 * F_ * planes * g_.H[1] * g_.sy;
        if (Rx_ > 1) halo_buf_elems_ = std::max(halo_buf_elems_, (size_t)2 * F_ * planes * (size_t)(g_.sz / g_.sy) * g_.H[0]);      // x slabs: H columns of every row
        halo_send_ = (FT*)dev_alloc(sizeof(FT) * halo_buf_elems_);
        halo_recv_ = (FT*)dev_alloc(sizeof(FT) * halo_buf_elems_);
        device_bytes += (int64_t)(2 * fft_.buffer_bytes + dfft_.work_bytes + 2 * sizeof(FT) * halo_buf_elems_);
    }
    for (int d = 0; d < 3; ++d) {
        const int N = (d == 1 && dist_) ? g_.N[d] * R_ : ((d == 0 && dist_) ? g_.N[d] * Rx_ : g_.N[d]);       // eigenvalues are global arrays (distributed_fft_based_poisson_solver.jl:106-112)
        std::vector<double> lam(N, 0.0);
        const double L = c.topology[d] == OC_FLAT ? 1.0 : c.extent[d];
        for (int i = 0; i < N; ++i) {                    // poisson_eigenvalues.jl:8-31 (Float64)
            if (c.topology[d] == OC_PERIODIC) lam[i] = std::pow(2.0 * std::sin(i * M_PI / N) / (L / N), 2);
            else if (c.topology[d] == OC_BOUNDED) lam[i] = std::pow(2.0 * std::sin(i * M_PI / (2.0 * N)) / (L / N), 2);
        }
        lam_[d] = (double*)dev_alloc(sizeof(double) * N);
        dev_upload(lam_[d], lam.data(), sizeof(double) * N, stream_);
        if (g_.bounded[d]) {
            std::vector<Cd> tw(N);
            for (int k = 0; k < N; ++k) { double a = -M_PI * k / (2.0 * N); tw[k] = Cd{std::cos(a), std::sin(a)}; }
            tw_[d] = (Cd*)dev_alloc(sizeof(Cd) * N);
            dev_upload(tw_[d], tw.data(), sizeof(Cd) * N, stream_);
        }
    }
    boxes_dev_ = nullptr;
    for (int i = 0; i < OC_TIMER_COUNT; ++i) { timer_ms_[i] = 0; timer_n_[i] = 0; }
    if (stretched_) {
        // FourierTridiagonalPoissonSolver(grid): the Thomas factors of every horizontal wavenumber's column
        const size_t n = (size_t)fft_.L.nxc * g_.N[1] * g_.N[2];
        tri_R_ = (FT*)dev_alloc(sizeof(FT) * n);
        tri_T_ = (FT*)dev_alloc(sizeof(FT) * n);
        device_bytes += (int64_t)(2 * sizeof(FT) * n);
        TridiagSetupKernel<FT> k;
        k.L = fft_.L;
        // (slabs: the local rows hold the y-wavenumbers rank·Ny_l … of the global eigenvalue array, and only the first rank owns the
        //  singular (0, 0) column)
        k.lam[0] = lam_[0] + (Rx_ > 1 ? (size_t)rx_ * g_.N[0] : 0); k.lam[1] = lam_[1] + (dist_ ? (size_t)rank_ * g_.N[1] : 0);
        k.zero_col = (!dist_ || (rank_ == 0 && rx_ == 0)) ? 1 : 0;
        k.dzc = g_.dzc; k.rdzf = g_.rdzf;
        k.R = tri_R_; k.T = tri_T_;
        k.eps10 = 10.0 * (double)std::numeric_limits<FT>::epsilon();
        Dim3 grid;
        grid.x = (fft_.L.nxc + TridiagSetupKernel<FT>::THREADS - 1) / TridiagSetupKernel<FT>::THREADS;
        grid.y = g_.N[1];
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    }
    sync();
}

// generate_coordinate for a Bounded, variably spaced coordinate (src/Grids/grid_generation.jl:33-94), in FT arithmetic:
// halo faces continue with the first / last interior spacing, centres are face averages, Δzᶜ[k] = F[k+1] - F[k],
// Δzᶠ[k] = C[k] - C[k-1].  Tables cover the 0-based levels -(H+1) … N+H+1 (the reference's cover a subset of that range with
// the same values: every halo spacing equals the edge spacing).
template <class FT>
void Model<FT>::build_z_tables(const double* faces) {
    const int N = g_.N[2], H = g_.H[2] + 1;
    const int nt = N + 2 * H + 1;                       // levels -H … N+H
    std::vector<FT> F(nt + 1), Cc(nt + 1);
    // faces: level k <-> F[k + H]; interior faces 0 … N
    for (int k = 0; k <= N; ++k) F[k + H] = (FT)faces[k];
    const FT dlo = F[H + 1] - F[H], dhi = F[H + N] - F[H + N - 1];
    for (int m = 1; m <= H; ++m) {                      // F₋[i] = c¹ - sum(Δᶠ₋[i:H]) : m equal terms summed left to right
        FT slo = FT(0), shi = FT(0);
        for (int q = 0; q < m; ++q) { slo = slo + dlo; shi = shi + dhi; }
        F[H - m] = F[H] - slo;
        F[H + N + m] = F[H + N] + shi;
    }
    // one more face above so that the centre of the topmost level exists
    { FT shi = FT(0); for (int q = 0; q < H + 1; ++q) shi = shi + dhi; F[nt] = F[H + N] + shi; }
    for (int n = 0; n < nt; ++n) Cc[n] = (F[n + 1] + F[n]) / FT(2);
    std::vector<FT> tab(15 * (size_t)nt);
    FT* dzc = tab.data(); FT* dzf = dzc + nt; FT* rdzc = dzf + nt; FT* rdzf = rdzc + nt; FT* rVc = rdzf + nt; FT* rVf = rVc + nt;
    FT* amd = rVf + nt;      // six AMD tables (AmdKernel::lv_*)
    const FT fx = FT(2) * g_.d[0], fy = FT(2) * g_.d[1];
    for (int n = 0; n < nt; ++n) {
        dzc[n] = F[n + 1] - F[n];
        dzf[n] = n > 0 ? Cc[n] - Cc[n - 1] : Cc[1] - Cc[0];
        rdzc[n] = FT(1) / dzc[n];
        rdzf[n] = FT(1) / dzf[n];
        rVc[n] = FT(1) / (g_.A[2] * dzc[n]);             // V = Az·Δz ; V⁻¹ = 1/V
        rVf[n] = FT(1) / (g_.A[2] * dzf[n]);
        // AMD: Δᶠz = 2 Δzᶜ[k] at the index of the evaluation point (anisotropic_minimum_dissipation.jl:224-234)
        const FT fz = FT(2) * dzc[n];
        amd[0 * nt + n] = (fx / fz) * g_.rd[0];
        amd[1 * nt + n] = (fz / fx) * rdzf[n];
        amd[2 * nt + n] = (fy / fz) * g_.rd[1];
        amd[3 * nt + n] = (fz / fy) * rdzf[n];
        amd[4 * nt + n] = fz * rdzf[n];
        amd[5 * nt + n] = FT(3) / (FT(1) / (fx * fx) + FT(1) / (fy * fy) + FT(1) / (fz * fz));
        // Smagorinsky: Δᶠ² with Δᶠ = cbrt(Δxᶜᶜᶜ Δyᶜᶜᶜ Δzᶜᶜᶜ)  (smagorinsky.jl:99-100)
        const FT df = std::cbrt(g_.d[0] * g_.d[1] * dzc[n]);
        amd[6 * nt + n] = df * df;
        // NonTraditionalBetaPlane: znode at Center / Face levels (grid.z.cᵃᵃᶜ, grid.z.cᵃᵃᶠ)
        amd[7 * nt + n] = Cc[n];
        amd[8 * nt + n] = F[n];
    }
    ztab_ = (FT*)dev_alloc(sizeof(FT) * tab.size());
    dev_upload(ztab_, tab.data(), sizeof(FT) * tab.size(), stream_);
#ifndef OC_HOSTSIM
    cuda_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");     // `tab` is a stack-owned staging buffer
#endif
    device_bytes += (int64_t)(sizeof(FT) * tab.size());
    g_.dzc = ztab_ + H; g_.dzf = g_.dzc + nt; g_.rdzc = g_.dzf + nt; g_.rdzf = g_.rdzc + nt; g_.rVc = g_.rdzf + nt; g_.rVf = g_.rVc + nt;
}

template <class FT>
Model<FT>::~Model() {
    graphs_clear();
#ifndef OC_HOSTSIM
    for (Stream st : {stream_, stream2_, stream3_}) if (st) cudaStreamSynchronize(st);
#endif
    // first the imports (peer buffers mapped through CUDA IPC) and the communicator, then the buffers this rank exported: no rank ever
    // waits, inside a free, for a peer that is itself waiting
    transport_.reset();
    auto fr = [](FieldRec& f) { dev_free(f.base); };
    for (auto& f : state_) fr(f);
    for (auto& f : next_) fr(f);
    for (auto& f : Gn_) fr(f);
    for (auto& f : Gm_) fr(f);
    for (int f = 0; f < OC_MAX_FIELDS; ++f)
        for (int s = 0; s < 6; ++s) if (bc_array_[f][s]) dev_free(bc_array_[f][s]);
    fr(pNHS_); fr(pHY_); fr(nu_e_);
    for (auto& f : kappa_e_) fr(f);
    dev_free(fftbuf_);
    dev_free(ztab_); dev_free(tri_R_); dev_free(tri_T_); dev_free(hi_tab_);
    dev_free(diag_dev_);
    dev_free(distT_); dev_free(diststage_); dev_free(halo_send_); dev_free(halo_recv_);
    for (int d = 0; d < 3; ++d) { dev_free(lam_[d]); dev_free(tw_[d]); }
    for (auto& kv : halo_cache_) dev_free(kv.second.boxes);
#ifndef OC_HOSTSIM
    for (void* e : event_pool_) cudaEventDestroy((cudaEvent_t)e);
    for (auto& r : timer_recs_) { cudaEventDestroy((cudaEvent_t)r.e0); cudaEventDestroy((cudaEvent_t)r.e1); }
    if (sw0_) { cudaEventDestroy((cudaEvent_t)sw0_); cudaEventDestroy((cudaEvent_t)sw1_); }
    if (ev_fork_) { cudaEventDestroy((cudaEvent_t)ev_fork_); cudaEventDestroy((cudaEvent_t)ev_join_); }
    if (ev_phy_) cudaEventDestroy((cudaEvent_t)ev_phy_);
    if (ev_xchg_) cudaEventDestroy((cudaEvent_t)ev_xchg_);
    for (void* e : ev_a2a_) cudaEventDestroy((cudaEvent_t)e);
    for (void* e : ev_mid_) cudaEventDestroy((cudaEvent_t)e);
    for (auto& s : out_slots_) {
        if (s.ev_done) cudaEventSynchronize((cudaEvent_t)s.ev_done);
        if (s.ev_snap) { cudaEventDestroy((cudaEvent_t)s.ev_snap); cudaEventDestroy((cudaEvent_t)s.ev_done); }
        if (s.stage) cudaFree(s.stage);
    }
    if (out_stream_) cudaStreamDestroy(out_stream_);
    if (in_stream_) cudaStreamDestroy(in_stream_);
    if (stream3_) cudaStreamDestroy(stream3_);
    if (stream2_) cudaStreamDestroy(stream2_);
    if (stream_) cudaStreamDestroy(stream_);
#endif
}

template <class FT>
typename Model<FT>::FieldRec Model<FT>::alloc_field(const int face[3]) {
    FieldRec f;
    f.base = (FT*)dev_alloc(sizeof(FT) * field_elems_);
    f.p = f.base + origin_off_;
    for (int d = 0; d < 3; ++d) f.face[d] = face[d];
    device_bytes += (int64_t)(sizeof(FT) * field_elems_);
    return f;
}

// field_boundary_conditions.jl:15-60 defaults + user overrides
template <class FT>
void Model<FT>::resolve_bcs(FieldRec& f, const oc_bc* user) {
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            SideBC r;
            r.value = 0.0;
            const int t = cfg_.topology[d];
            int kind = user ? user[2 * d + s].kind : OC_BC_DEFAULT;
            if (t == OC_PERIODIC) {
                if (kind != OC_BC_DEFAULT && kind != OC_BC_PERIODIC) throw Error(OC_ERR_INVALID, "non-periodic boundary condition in a Periodic dimension");
                r.kind = OC_BC_PERIODIC;
            } else if (t == OC_FLAT) {
                r.kind = OC_BC_NONE;
            } else {
                if (!(s == 0 ? g_.wlo[d] : g_.whi[d])) {
                    // a connected side of a slab: the neighbour's rows arrive by halo exchange, whatever the user declared for the
                    // global boundary (inject_halo_communication_boundary_conditions, halo_communication_bcs.jl:14-51)
                    r.kind = OC_BC_NONE;
                    f.bc[2 * d + s] = r;
                    continue;
                }
                if (kind == OC_BC_DEFAULT) kind = f.face[d] ? OC_BC_OPEN : OC_BC_FLUX;
                if (kind == OC_BC_PERIODIC) throw Error(OC_ERR_INVALID, "periodic boundary condition in a Bounded dimension");
                if (f.face[d] && kind != OC_BC_OPEN) throw Error(OC_ERR_UNSUPPORTED, "wall-normal velocity supports only Open (impenetrable) boundary conditions");
                if (!f.face[d] && kind == OC_BC_OPEN) throw Error(OC_ERR_INVALID, "Open boundary condition on a field that is not wall-normal");
                r.kind = kind;
                if (user && user[2 * d + s].has_value) r.value = user[2 * d + s].value;
            }
            f.bc[2 * d + s] = r;
        }
}

template <class FT>
void Model<FT>::sync() {
#ifndef OC_HOSTSIM
    join_tracers();
    join_exchange();
    cuda_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
#endif
}

// The tracer tendency kernels of a stage only read the OLD velocities and their own tracer, and the pressure solve only touches
// the NEW velocities: they run concurrently on two streams (FP64-bound stencils overlap the HBM-bound FFT passes and, on several
// GPUs, the NCCL transposes).  fork: stream2 waits for everything issued so far; join: stream_ waits for the tracer kernels.
template <class FT>
void Model<FT>::fork_tracers() {
#ifndef OC_HOSTSIM
    cuda_check(cudaEventRecord((cudaEvent_t)ev_fork_, stream_), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(stream2_, (cudaEvent_t)ev_fork_, 0), "cudaStreamWaitEvent");
    launch_stream_ = stream2_;
#endif
    tracers_in_flight_ = true;
}
template <class FT>
void Model<FT>::join_tracers() {
    if (!tracers_in_flight_) return;
#ifndef OC_HOSTSIM
    cuda_check(cudaEventRecord((cudaEvent_t)ev_join_, stream2_), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(stream_, (cudaEvent_t)ev_join_, 0), "cudaStreamWaitEvent");
#endif
    tracers_in_flight_ = false;
}

// An exception that crosses an entry point may leave launch_stream_ on a side stream (fork_tracers, the sub-chunked y stage of the
// distributed solve) with work in flight there: go back to the main stream and make it wait for the side streams, so that later
// calls on the (still usable) handle are ordered instead of racing.
template <class FT>
void Model<FT>::recover() {
#ifndef OC_HOSTSIM
    launch_stream_ = stream_;
    cudaEvent_t ev = (cudaEvent_t)ev_join_;
    if (ev) {
        if (stream2_ && cudaEventRecord(ev, stream2_) == cudaSuccess) cudaStreamWaitEvent(stream_, ev, 0);
        if (stream3_ && cudaEventRecord(ev, stream3_) == cudaSuccess) cudaStreamWaitEvent(stream_, ev, 0);
    }
    cudaGetLastError();
#endif
    tracers_in_flight_ = false;
    xchg_pending_ = false;           // (stream3_ was joined above)
}

template <class FT>
typename Model<FT>::FieldRec& Model<FT>::lookup(int field) {
    if (field >= 0 && field < F_) return state_[field];
    if (field == OC_FIELD_PNHS) return pNHS_;
    if (field == OC_FIELD_PHY && has_pHY_) return pHY_;
    if (field == OC_FIELD_NU_E && has_eddy_) return nu_e_;
    if (field >= OC_FIELD_KAPPA_E0 && field < OC_FIELD_KAPPA_E0 + (int)kappa_e_.size()) return kappa_e_[field - OC_FIELD_KAPPA_E0];
    if (field >= OC_FIELD_GN0 && field < OC_FIELD_GN0 + F_) return Gn_[field - OC_FIELD_GN0];
    if (field >= OC_FIELD_GM0 && field < OC_FIELD_GM0 + F_) return Gm_[field - OC_FIELD_GM0];
    throw Error(OC_ERR_INVALID, "unknown field id " + std::to_string(field));
}

template <class FT>
void Model<FT>::field_info(int field, oc_field_info* info) {
    if ((field == OC_FIELD_PHY || field == OC_FIELD_NU_E || (field >= OC_FIELD_KAPPA_E0 && field < OC_FIELD_GN0)) && !aux_valid_) aux();
    // Gⁿ AND G⁻ are API-valid only after the evaluation the reference's update_state! performs at the end of every step: between
    // steps the library keeps the last substep's tendencies in the Gⁿ slot (they become G⁻ by a pointer swap at the next stage)
    if ((field >= OC_FIELD_GN0 && field < OC_FIELD_GN0 + F_) || (field >= OC_FIELD_GM0 && field < OC_FIELD_GM0 + F_)) compute_tendencies_if_stale();
    FieldRec& f = lookup(field);
    for (int d = 0; d < 3; ++d) {
        info->location[d] = f.face[d];
        info->interior_size[d] = g_.N[d] + ((f.face[d] && g_.whi[d]) ? 1 : 0);
        info->parent_size[d] = info->interior_size[d] + 2 * Hcfg_[d];
    }
    info->device_ptr = f.p;
    info->stride_y = g_.sy;
    info->stride_z = g_.sz;
}

template <class FT>
void Model<FT>::transfer(int field, void* host, size_t nbytes, bool parent, bool upload) {
    oc_field_info info;
    field_info(field, &info);
    FieldRec& f = lookup(field);
    int ext[3], lo[3];
    size_t n = 1;
    for (int d = 0; d < 3; ++d) {
        ext[d] = parent ? info.parent_size[d] : info.interior_size[d];
        lo[d] = parent ? -Hcfg_[d] : 0;
        n *= (size_t)ext[d];
    }
    if (n * sizeof(FT) != nbytes) throw Error(OC_ERR_INVALID, "host buffer size mismatch: expected " + std::to_string(n * sizeof(FT)) + " bytes");
    FT* origin = f.p + lo[0] + (long long)lo[1] * g_.sy + (long long)lo[2] * g_.sz;
    dev_copy_box(origin, sizeof(FT), g_.sy, g_.sz, host, ext, upload, stream_);
    if (upload) {
        // Flat dimensions are stored as periodic N=1: refresh their (internal) halo copies
        if (g_.flat[0] || g_.flat[1] || g_.flat[2]) { std::vector<FieldRec*> one{&f}; halo(one, false); }
        if (field < F_) { tend_valid_ = false; aux_valid_ = false; }
        if (array_diff_ && (field == OC_FIELD_NU_E || (field >= OC_FIELD_KAPPA_E0 && field < OC_FIELD_GN0))) tend_valid_ = false;
    }
}

// ---------------------------------------------------------------------------------------------------------
// timers
// ---------------------------------------------------------------------------------------------------------
template <class FT>
void Model<FT>::begin_timer(int cls) {
#ifndef OC_HOSTSIM
    if (!timing_) return;
    auto get = [&]() -> void* {
        if (!event_pool_.empty()) { void* e = event_pool_.back(); event_pool_.pop_back(); return e; }
        cudaEvent_t e;
        cuda_check(cudaEventCreate(&e), "cudaEventCreate");
        return (void*)e;
    };
    TimerRec r{cls, get(), get()};
    cuda_check(cudaEventRecord((cudaEvent_t)r.e0, launch_stream_), "cudaEventRecord");
    timer_recs_.push_back(r);
#else
    (void)cls;
#endif
}
template <class FT>
void Model<FT>::end_timer() {
#ifndef OC_HOSTSIM
    if (!timing_) return;
    cuda_check(cudaEventRecord((cudaEvent_t)timer_recs_.back().e1, launch_stream_), "cudaEventRecord");
#endif
}
template <class FT>
void Model<FT>::collect_timers() {
#ifndef OC_HOSTSIM
    sync();
    for (auto& r : timer_recs_) {
        float ms = 0;
        cuda_check(cudaEventElapsedTime(&ms, (cudaEvent_t)r.e0, (cudaEvent_t)r.e1), "cudaEventElapsedTime");
        timer_ms_[r.cls] += ms;
        timer_n_[r.cls] += 1;
        event_pool_.push_back(r.e0);
        event_pool_.push_back(r.e1);
    }
    timer_recs_.clear();
#endif
}
template <class FT>
void Model<FT>::timers_reset() {
    collect_timers();
    for (int i = 0; i < OC_TIMER_COUNT; ++i) { timer_ms_[i] = 0; timer_n_[i] = 0; }
}
template <class FT>
void Model<FT>::timers_get(double* ms, int64_t* n) {
    collect_timers();
    for (int i = 0; i < OC_TIMER_COUNT; ++i) { ms[i] = timer_ms_[i]; n[i] = timer_n_[i]; }
}

template <class FT>
void Model<FT>::stopwatch_start() {
#ifndef OC_HOSTSIM
    if (!sw0_) {
        cudaEvent_t a, b;
        cuda_check(cudaEventCreate(&a), "cudaEventCreate");
        cuda_check(cudaEventCreate(&b), "cudaEventCreate");
        sw0_ = a; sw1_ = b;
    }
    cuda_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
    cuda_check(cudaEventRecord((cudaEvent_t)sw0_, stream_), "cudaEventRecord");
#endif
}
template <class FT>
double Model<FT>::stopwatch_stop() {
#ifndef OC_HOSTSIM
    if (!sw0_) throw Error(OC_ERR_STATE, "stopwatch not started");
    cuda_check(cudaEventRecord((cudaEvent_t)sw1_, stream_), "cudaEventRecord");
    cuda_check(cudaEventSynchronize((cudaEvent_t)sw1_), "cudaEventSynchronize");
    float ms = 0;
    cuda_check(cudaEventElapsedTime(&ms, (cudaEvent_t)sw0_, (cudaEvent_t)sw1_), "cudaEventElapsedTime");
    return ms;
#else
    return 0.0;
#endif
}

template <class FT>
template <class K>
void Model<FT>::go(const K& k, Dim3 grid, size_t smem, int cls) {
    begin_timer(cls);
    cudaError_t e = replay_ ? cudaSuccess : launch(k, grid, smem, launch_stream_);
    end_timer();
#ifndef OC_HOSTSIM
    cuda_check(e, "kernel launch");
#else
    (void)e;
#endif
    ++launches;
}

// ---------------------------------------------------------------------------------------------------------
// halo fill: one launch for any list of fields
// ---------------------------------------------------------------------------------------------------------
template <class FT>
void Model<FT>::halo(const std::vector<FieldRec*>& fields, bool fill_open, bool defer_exchange) {
    NvtxRange nvtx_("fill_halo_regions!");
    if (fields.empty()) return;
    if ((int)fields.size() > HALO_MAX_FIELDS) throw Error(OC_ERR_INVALID, "too many fields in one halo fill");
    std::string key;
    for (FieldRec* f : fields) key += (char)('0' + f->face[0] + 2 * f->face[1] + 4 * f->face[2]);
    auto it = halo_cache_.find(key);
    if (it == halo_cache_.end()) {
        std::vector<HaloBox> boxes;
        int nb = 0;
        auto add = [&](int fi, int lo0, int n0, int lo1, int n1, int lo2, int n2) {
            if (n0 <= 0 || n1 <= 0 || n2 <= 0) return;
            HaloBox b;
            b.field = fi;
            b.lo[0] = lo0; b.lo[1] = lo1; b.lo[2] = lo2;
            b.n[0] = n0; b.n[1] = n1; b.n[2] = n2;
            b.first_block = nb;
            b.tshift = n0 <= 4 ? 2 : (n0 <= 8 ? 3 : (n1 <= 4 ? 6 : 5));     // 4 × 64, 8 × 32, 64 × 4 (slabs a few rows high) or 32 × 8 threads per tile
            const int tw = 1 << b.tshift, th = HaloKernel<FT>::THREADS >> b.tshift;
            b.nbx = (n0 + tw - 1) / tw;
            b.nby = (n1 + th - 1) / th;
            const long long blocks = (long long)b.nbx * b.nby * ((n2 + HaloKernel<FT>::ZPT - 1) / HaloKernel<FT>::ZPT);
            if ((long long)nb + blocks >= (1LL << 31)) throw Error(OC_ERR_UNSUPPORTED, "halo slabs with 2^31 or more blocks");
            nb += (int)blocks;
            boxes.push_back(b);
        };
        const int* N = g_.N;
        const int* H = g_.H;
        for (int fi = 0; fi < (int)fields.size(); ++fi) {
            const int X0 = -H[0], XN = N[0] + 2 * H[0] + 1, Y0 = -H[1], YN = N[1] + 2 * H[1] + 1;
            add(fi, X0, XN, Y0, YN, -H[2], H[2]);                  // bottom slab
            add(fi, X0, XN, Y0, YN, N[2], H[2] + 1);               // top slab (incl. the extra Face plane)
            add(fi, X0, XN, -H[1], H[1], 0, N[2]);                 // south
            add(fi, X0, XN, N[1], H[1] + 1, 0, N[2]);              // north
            add(fi, -H[0], H[0], 0, N[1], 0, N[2]);                // west
            add(fi, N[0], H[0] + 1, 0, N[1], 0, N[2]);             // east
            for (int d = 0; d < 3; ++d)                            // lower wall plane of a wall-normal velocity
                if (fields[fi]->face[d] && g_.bounded[d]) {
                    int lo[3] = {0, 0, 0}, n[3] = {N[0], N[1], N[2]};
                    n[d] = 1;
                    add(fi, lo[0], n[0], lo[1], n[1], lo[2], n[2]);
                }
        }
        HaloCache hc;
        hc.nboxes = (int)boxes.size();
        hc.nblocks = nb;
        hc.boxes = (HaloBox*)dev_alloc(sizeof(HaloBox) * boxes.size());
        dev_upload(hc.boxes, boxes.data(), sizeof(HaloBox) * boxes.size(), stream_);
        sync();
        it = halo_cache_.emplace(key, hc).first;
    }
    HaloKernel<FT> k;
    k.g = g_;
    k.nfields = (int)fields.size();
    k.nboxes = it->second.nboxes;
    k.fill_open = fill_open ? 1 : 0;
    k.skip[0] = Rx_ > 1 ? 1 : 0; k.skip[1] = (dist_ && R_ > 1) ? 1 : 0; k.skip[2] = 0;
    k.boxes = it->second.boxes;
    for (int fi = 0; fi < (int)fields.size(); ++fi) {
        k.f[fi].p = fields[fi]->p;
        for (int d = 0; d < 3; ++d) k.f[fi].face[d] = fields[fi]->face[d];
        for (int s = 0; s < 6; ++s) k.f[fi].bc[s] = fields[fi]->bc[s];
    }
    Dim3 grid;
    grid.x = it->second.nblocks;
    go(k, grid, 0, OC_TIMER_HALO);
    // array-valued Value / Gradient BCs of prognostic fields: rewrite those sides' halo plane (HaloArrayKernel)
    for (FieldRec* fr : fields) {
        if (state_.empty() || fr < &state_[0] || fr >= &state_[0] + F_) continue;
        const int f = (int)(fr - &state_[0]);
        for (int s = 0; s < 6; ++s) {
            if (!bc_array_[f][s] || (fr->bc[s].kind != OC_BC_VALUE && fr->bc[s].kind != OC_BC_GRADIENT)) continue;
            HaloArrayKernel<FT> hk;
            hk.g = g_;
            hk.p = fr->p;
            hk.A = bc_array_[f][s];
            hk.d = s / 2; hk.side = s % 2; hk.kind = fr->bc[s].kind;
            const int t1 = hk.d == 0 ? 1 : 0, t2 = hk.d == 2 ? 1 : 2;
            hk.n1 = g_.N[t1];
            // the cells HaloKernel writes for this side: interior range of Bounded tangential dimensions, whole extent of periodic ones
            // (a partitioned periodic dimension: interior only — its halo cells arrive with the exchange that follows)
            const bool in1 = g_.bounded[t1] || (t1 == 0 && Rx_ > 1) || (t1 == 1 && dist_ && R_ > 1);
            const bool in2 = g_.bounded[t2] || (t2 == 0 && Rx_ > 1) || (t2 == 1 && dist_ && R_ > 1);
            hk.lo1 = in1 ? 0 : -g_.H[t1]; hk.m1 = in1 ? g_.N[t1] : g_.N[t1] + 2 * g_.H[t1] + 1;
            hk.lo2 = in2 ? 0 : -g_.H[t2]; hk.m2 = in2 ? g_.N[t2] : g_.N[t2] + 2 * g_.H[t2] + 1;
            Dim3 hg;
            hg.x = (hk.m1 + HaloArrayKernel<FT>::THREADS - 1) / HaloArrayKernel<FT>::THREADS; hg.y = hk.m2; hg.z = 1;
            go(hk, hg, 0, OC_TIMER_HALO);
        }
    }
    if (!dist_) return;
    // pencils: x first, over every row of the parent array (wall-side boundary-condition halos included; the y-halo rows it carries are
    // overwritten next), then y over whole padded rows — the corners arrive through the two hops, where the reference sends separate
    // corner messages (halo_communication.jl:217-333)
    if (Rx_ > 1) exchange_x(fields);
    if (R_ <= 1) return;
#ifndef OC_HOSTSIM
    if (defer_exchange && stream3_) {
        // The exchange runs on the communication stream while the NEXT stage's interior tendency kernels — which read no y-halo row —
        // run on the main stream (interleave_communication_and_computation.jl:29-67); tendencies() waits for ev_xchg_ before the two
        // boundary strips.  No other NCCL call is enqueued anywhere until that wait, so the communicator sees one stream at a time.
        if (!ev_xchg_) { cudaEvent_t e; cuda_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate"); ev_xchg_ = e; }
        cuda_check(cudaEventRecord((cudaEvent_t)ev_fork_, stream_), "cudaEventRecord");
        cuda_check(cudaStreamWaitEvent(stream3_, (cudaEvent_t)ev_fork_, 0), "cudaStreamWaitEvent");
        launch_stream_ = stream3_;
        try { exchange_y(fields); } catch (...) { launch_stream_ = stream_; throw; }
        launch_stream_ = stream_;
        cuda_check(cudaEventRecord((cudaEvent_t)ev_xchg_, stream3_), "cudaEventRecord");
        xchg_pending_ = true;
        return;
    }
#else
    // host simulation: the exchange itself is synchronous, but the NEXT tendencies() still takes the split path (interior tile rows,
    // then the two boundary strips) — so that the partitioning of the launches is exercised by the CPU tests
    if (defer_exchange) { exchange_y(fields); xchg_pending_ = true; return; }
#endif
    exchange_y(fields);
}

// the main stream waits for a deferred y-halo exchange (every consumer of halo rows other than the split tendencies() calls this)
template <class FT>
void Model<FT>::join_exchange() {
    if (!xchg_pending_) return;
#ifndef OC_HOSTSIM
    cuda_check(cudaStreamWaitEvent(stream_, (cudaEvent_t)ev_xchg_, 0), "cudaStreamWaitEvent");
#endif
    xchg_pending_ = false;
}

// ---------------------------------------------------------------------------------------------------------
// distributed: y-halo exchange with the two slab neighbours (halo_communication.jl:87-333)
// ---------------------------------------------------------------------------------------------------------
template <class FT>
void Model<FT>::dist_attach(Transport* t) {
    if (!dist_) { delete t; throw Error(OC_ERR_STATE, "the model was not created with dist_nranks > 1"); }
    transport_.reset(t);
    // Peer memory for the transposed FFT (OC_DIST_P2P=0 keeps the all-to-all path: measurement / machines without P2P)
    const char* p2p_env = getenv("OC_DIST_P2P");
    p2p_ = false;
    // Measured (profiles/r02f…r02j): 2 GPUs 70.7 vs 72.7 ms per step, 4 GPUs 71.8 vs 72.5 with / without peer-memory transposes; at 8 GPUs
    // the first version (every rank writing to rank 0 first: incast) lost, 88.9 vs 77.4; the round-robin version: 77.2 vs 76.3 (r02o) — the
    // sub-chunked all-to-all still wins there.  So the default is peer memory up to 4 ranks and the all-to-all beyond; OC_DIST_P2P=1 forces it.
    if (Rx_ == 1 && R_ <= DIST_MAX_RANKS && (p2p_env ? atoi(p2p_env) != 0 : R_ <= 4)) {
        std::string e1 = transport_->map_peers(fftbuf_, peer_spec_, stream_);
        std::string e2 = transport_->map_peers(distT_, peer_T_, stream_);
        if (e2.empty()) e2 = e1;
        p2p_ = transport_->agree(!e2.empty(), stream_) == 0;          // every rank or none
#ifndef OC_HOSTSIM
        if (!p2p_ && getenv("OC_VERBOSE")) fprintf(stderr, "oceananigans_b200: peer-memory transposes unavailable (%s): NCCL all-to-all path\n", e2.c_str());
        if (p2p_) dfft_.set_y_stream(stream_);
#endif
    }
    if (getenv("OC_VERBOSE")) fprintf(stderr, "oceananigans_b200: rank %d of %d: distributed FFT transposes over %s\n", rank_, R_, p2p_ ? "peer memory" : "send/receive all-to-all");
}

// FFT(z,x) local -> transposed put into the owners' buffers -> FFT(y) -> divide -> FFT⁻¹(y) -> transposed put back -> FFT⁻¹(z,x)
// over peer memory: the all-to-all and the transpose are one kernel (TransposePutKernel).  Three barriers order the ranks: nobody
// writes a peer's T before that peer has left its previous y stage, nobody reads T before every put has landed, nobody reads the
// spectral buffer before every put back has landed.
template <class FT>
void Model<FT>::run_fft_solve_p2p() {
    NvtxRange nvtx_("distributed FFT solve (peer memory)");
    auto chk = [](const std::string& e) { if (!e.empty()) throw Error(OC_ERR_CUDA, e); };
    auto barrier = [&]() { begin_timer(OC_TIMER_COMM); std::string e = transport_->barrier(stream_); end_timer(); chk(e); };
    begin_timer(OC_TIMER_FFT); std::string e = dfft_.zx(fftbuf_, true); end_timer(); chk(e);
    local_twiddles(false);
    TransposePutKernel<FT> t;
    t.nxc = dfft_.nxc; t.nyl = g_.N[1]; t.nzl = dfft_.Nzl; t.R = R_; t.rank = rank_; t.yperm = g_.bounded[1];
    for (int r = 0; r < R_; ++r) { t.spec[r] = reinterpret_cast<Cplx<FT>*>(peer_spec_[r]); t.T[r] = reinterpret_cast<Cplx<FT>*>(peer_T_[r]); }
    Dim3 tg;
    barrier();
    t.forward = 1;
    tg.x = (dfft_.nxc + 31) / 32; tg.y = (g_.N[1] + 31) / 32; tg.z = g_.N[2];
    go(t, tg, TransposePutKernel<FT>::SMEM, OC_TIMER_COMM);
    barrier();
    const int C = dfft_.C, nz = dfft_.Nzl / C;
    for (int c = 0; c < C; ++c) { begin_timer(OC_TIMER_FFT); e = dfft_.y(distT_, true, c); end_timer(); chk(e); }
    PoissonDivideTKernel<FT> k;
    k.nxc = dfft_.nxc; k.ny = dfft_.Ny; k.nzl = dfft_.Nzl; k.kz0 = rank_ * dfft_.Nzl; k.zl0 = 0;
    k.T = reinterpret_cast<Cplx<FT>*>(distT_);
    for (int d = 0; d < 3; ++d) k.lam[d] = lam_[d];
    k.tw = g_.bounded[1] ? tw_[1] : nullptr;
    k.norm = 1.0 / ((double)g_.N[0] * dfft_.Ny * g_.N[2]);
    Dim3 grid;
    grid.x = (dfft_.Ny + 255) / 256; grid.y = dfft_.nxc; grid.z = dfft_.Nzl;
    go(k, grid, 0, OC_TIMER_POISSON_MID);
    for (int c = 0; c < C; ++c) { begin_timer(OC_TIMER_FFT); e = dfft_.y(distT_, false, c); end_timer(); chk(e); }
    (void)nz;
    t.forward = 0;
    tg.x = (dfft_.nxc + 31) / 32; tg.y = (dfft_.Ny + 31) / 32; tg.z = dfft_.Nzl;
    go(t, tg, TransposePutKernel<FT>::SMEM, OC_TIMER_COMM);
    barrier();
    local_twiddles(true);
    begin_timer(OC_TIMER_FFT); e = dfft_.zx(fftbuf_, false); end_timer(); chk(e);
}

// the DCT twiddles of the dimensions that are whole in the slab layout — x and z — on the local spectral buffer (TwiddleKernel)
template <class FT>
void Model<FT>::local_twiddles(bool inverse) {
    for (int d = 0; d < 3; d += 2) {
        if (!g_.bounded[d] || (d == 0 && Rx_ > 1) || (d == 2 && stretched_)) continue;      // (a stretched z is not transformed)
        TwiddleKernel<FT> k;
        const long long plane = (long long)dfft_.nxc * g_.N[1];
        k.N = g_.N[d]; k.spec = reinterpret_cast<Cplx<FT>*>(fftbuf_); k.tw = tw_[d]; k.inverse = inverse ? 1 : 0;
        if (d == 2) { k.sk = plane; k.inner = (int)plane; k.souter = 0; k.count = plane; k.kfast = 0; }
        else { k.sk = 1; k.inner = 1; k.souter = dfft_.nxc; k.count = (long long)g_.N[1] * g_.N[2]; k.kfast = 1; }
        Dim3 grid;
        grid.x = (int)((k.count * (k.N / 2 + 1) + 255) / 256);
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    }
}

// DistributedFourierTridiagonalPoissonSolver for a vertically stretched grid on slabs (distributed_fft_tridiagonal_solver.jl:262-293):
// transforms along x (local rows) and y (in the transposed layout), the transposed layout brought BACK so that every rank holds whole
// z-columns of its share of the (kx, ky) plane, the tridiagonal solves there (TridiagSolvePPKernel — the twiddles of Bounded x / y are
// separate passes here, so the real-coefficient kernel serves every topology), and the same way out.  Twice the transposes of the
// regular-grid solve, like the reference.  The middle pair of transposes carries no Makhoul permutation: the spectral coefficients keep
// their natural order, local row yl = wavenumber rank·Ny_l + yl.  (The reference's distributed solver leaves the mean of ϕ wherever the
// singular column's pivot puts it; this one removes it like the single-GPU solver, fourier_tridiagonal_poisson_solver.jl:222-226 — a
// constant that no gradient sees.)
template <class FT>
void Model<FT>::run_fft_solve_dist_tridiagonal() {
    NvtxRange nvtx_("distributed Fourier-tridiagonal solve");
    auto chk = [](const std::string& e) { if (!e.empty()) throw Error(OC_ERR_CUDA, e); };
    const int nxc = dfft_.nxc, nyl = g_.N[1], Ny = dfft_.Ny, nzl = dfft_.Nzl, C = dfft_.C;
    Cplx<FT>* T = reinterpret_cast<Cplx<FT>*>(distT_);
    auto transpose_y = [&](bool to_T, int yperm) {
        TransposeKernel<FT> t;
        t.nxc = nxc; t.nyl = nyl; t.nzl = nzl; t.R = R_; t.zl0 = 0; t.yperm = yperm;
        t.stage = reinterpret_cast<Cplx<FT>*>(diststage_); t.T = T; t.to_T = to_T ? 1 : 0;
        Dim3 tg;
        tg.x = (nxc + 31) / 32; tg.y = (Ny + 31) / 32; tg.z = nzl;
        go(t, tg, TransposeKernel<FT>::SMEM, OC_TIMER_POISSON_MID);
    };
    auto y_twiddles = [&](bool inverse) {
        if (!g_.bounded[1]) return;
        TwiddleKernel<FT> k;
        k.N = Ny; k.spec = T; k.tw = tw_[1]; k.inverse = inverse ? 1 : 0;
        k.sk = 1; k.inner = 1; k.souter = Ny; k.count = (long long)nzl * nxc; k.kfast = 1;
        Dim3 grid;
        grid.x = (int)((k.count * (Ny / 2 + 1) + 255) / 256);
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    };
    auto y_ffts = [&](bool fwd) {
        for (int c = 0; c < C; ++c) { begin_timer(OC_TIMER_FFT); std::string e = dfft_.y(distT_, fwd, c); end_timer(); chk(e); }
    };
    auto to_T = [&](int yperm) { for (int c = 0; c < C; ++c) all_to_all(fftbuf_, diststage_, c, C); transpose_y(true, yperm); };
    auto from_T = [&](int yperm) { transpose_y(false, yperm); for (int c = 0; c < C; ++c) all_to_all(diststage_, fftbuf_, c, C); };
    begin_timer(OC_TIMER_FFT); std::string e = dfft_.zx(fftbuf_, true); end_timer(); chk(e);        // x rows only (DistFft::xonly)
    local_twiddles(false);
    to_T(g_.bounded[1]);
    y_ffts(true);
    y_twiddles(false);
    from_T(0);
    {
        TridiagSolvePPKernel<FT> k;
        k.L = fft_.L;
        k.spec = fftbuf_;
        k.R = tri_R_; k.T = tri_T_; k.rdzf = g_.rdzf;
        k.norm = 1.0 / ((double)g_.N[0] * Ny);
        k.zero_col = rank_ == 0 ? 1 : 0;
        Dim3 grid;
        const long long n2 = 2LL * fft_.L.nxc * g_.N[1];
        grid.x = (int)((n2 + TridiagSolvePPKernel<FT>::THREADS - 1) / TridiagSolvePPKernel<FT>::THREADS);
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    }
    to_T(0);
    y_twiddles(true);
    y_ffts(false);
    from_T(g_.bounded[1]);
    local_twiddles(true);
    begin_timer(OC_TIMER_FFT); e = dfft_.zx(fftbuf_, false); end_timer(); chk(e);
}

// pencils: the x-halo exchange with the two neighbours along x — H columns of every row of the parent array
template <class FT>
void Model<FT>::exchange_x(const std::vector<FieldRec*>& fields) {
    NvtxRange nvtx_("halo exchange (x)");
    if (!transport_) throw Error(OC_ERR_STATE, "distributed model without a transport: call oc_dist_attach_nccl first");
    const int nf = (int)fields.size();
    if (nf > F_) throw Error(OC_ERR_INVALID, "halo exchange of more fields than the exchange buffers hold");
    const int planes = (int)(field_elems_ / (size_t)g_.sz), rows = g_.sz / g_.sy;
    HaloPackXKernel<FT> k;
    k.g = g_;
    k.nfields = nf; k.planes = planes; k.rows = rows; k.cols = g_.H[0];
    for (int f = 0; f < nf; ++f) { k.base[f] = fields[f]->base; k.x0[f] = (int)((fields[f]->p - fields[f]->base) % g_.sy); }
    const size_t per_side = (size_t)nf * planes * rows * g_.H[0];
    if (2 * per_side > halo_buf_elems_) throw Error(OC_ERR_STATE, "internal: exchange buffer too small for the x slabs");
    Dim3 grid;
    grid.x = (int)((2 * per_side + 255) / 256);
    k.unpack = 0; k.buf = halo_send_;
    go(k, grid, 0, OC_TIMER_COMM);
    const int prev = grank((rx_ + Rx_ - 1) % Rx_, rank_), next = grank((rx_ + 1) % Rx_, rank_);
    const size_t to_prev = g_.wlo[0] ? 0 : per_side * sizeof(FT), to_next = g_.whi[0] ? 0 : per_side * sizeof(FT);
    std::vector<Msg> msgs;
    msgs.push_back(Msg{prev, next, 4, halo_send_, to_prev, halo_recv_ + per_side, to_next});
    msgs.push_back(Msg{next, prev, 5, halo_send_ + per_side, to_next, halo_recv_, to_prev});
    begin_timer(OC_TIMER_COMM);
    std::string e = transport_->exchange(msgs, launch_stream_);
    end_timer();
    if (!e.empty()) throw Error(OC_ERR_CUDA, e);
    k.unpack = 1; k.buf = halo_recv_;
    go(k, grid, 0, OC_TIMER_COMM);
}

// all-to-all of the Rx equal chunks of a spectral buffer among the ranks of this rank's row (same y index)
template <class FT>
void Model<FT>::all_to_all_x(FT* send, FT* recv) {
    const size_t chunk = fft_.buffer_bytes / Rx_;
    std::vector<Msg> msgs;
    for (int d = 1; d < Rx_; ++d) {
        const int to = (rx_ + d) % Rx_, from = (rx_ + Rx_ - d) % Rx_;
        msgs.push_back(Msg{grank(to, rank_), grank(from, rank_), 600 + d, (char*)send + chunk * to, chunk, (char*)recv + chunk * from, chunk});
    }
    begin_timer(OC_TIMER_COMM);
#ifndef OC_HOSTSIM
    cuda_check(cudaMemcpyAsync((char*)recv + chunk * rx_, (char*)send + chunk * rx_, chunk, cudaMemcpyDeviceToDevice, stream_), "cudaMemcpyAsync D2D");
#else
    memcpy((char*)recv + chunk * rx_, (char*)send + chunk * rx_, chunk);
#endif
    std::string e = transport_ ? transport_->exchange(msgs, stream_) : std::string("distributed model without a transport");
    end_timer();
    if (!e.empty()) throw Error(OC_ERR_CUDA, e);
}

// Pencils, Partition(Rx, Ry) (distributed_fft_based_poisson_solver.jl:141-178, the two-transpose branch):
//   local (Nx_l, Ny_l, Nz)  FFT z  ->  transpose z <-> y among the Ry ranks of this column (the slab machinery, with Nx_l for the row
//   length)  ->  T1 = [zl][xl][y]  FFT y  ->  transpose y <-> x among the Rx ranks of this row  ->  T2 = [zl][yl2][x]  FFT x  ->  divide
//   -> and back.  Every stage is complex-to-complex; Bounded dimensions carry their DCT as a Makhoul permutation applied where the
//   line becomes whole (z: by the right-hand-side kernel; y, x: inside the transposes) and twiddles after / before the FFTs.
template <class FT>
void Model<FT>::run_fft_solve_pencil() {
    NvtxRange nvtx_("distributed FFT solve (pencils)");
    auto chk = [](const std::string& e) { if (!e.empty()) throw Error(OC_ERR_CUDA, e); };
    const int nxl = g_.N[0], nyl = g_.N[1], Ny = dfft_.Ny, Nx = nxl * Rx_, nzl = dfft_.Nzl, nyx = Ny / Rx_;
    const int C = dfft_.C;
    Cplx<FT>* spec = reinterpret_cast<Cplx<FT>*>(fftbuf_);
    Cplx<FT>* stage = reinterpret_cast<Cplx<FT>*>(diststage_);
    Cplx<FT>* T = reinterpret_cast<Cplx<FT>*>(distT_);
    auto twiddle_lines = [&](Cplx<FT>* base, int n, long long count, const Cd* tw, bool inverse) {
        TwiddleKernel<FT> k;
        k.N = n; k.spec = base; k.tw = tw; k.inverse = inverse ? 1 : 0;
        k.sk = 1; k.inner = 1; k.souter = n; k.count = count; k.kfast = 1;
        Dim3 grid;
        grid.x = (int)((count * (n / 2 + 1) + 255) / 256);
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    };
    auto transpose_y = [&](bool to_T, int yperm) {
        TransposeKernel<FT> t;
        t.nxc = nxl; t.nyl = nyl; t.nzl = nzl; t.R = R_; t.zl0 = 0; t.yperm = yperm;
        t.stage = stage; t.T = T; t.to_T = to_T ? 1 : 0;
        Dim3 tg;
        tg.x = (nxl + 31) / 32; tg.y = (Ny + 31) / 32; tg.z = nzl;
        go(t, tg, TransposeKernel<FT>::SMEM, OC_TIMER_POISSON_MID);
    };
    auto pencil_x = [&](int mode, Cplx<FT>* B, Cplx<FT>* Tb, int xperm) {
        PencilXKernel<FT> k;
        k.nxl = nxl; k.nyx = nyx; k.nzl = nzl; k.Rx = Rx_; k.xperm = xperm; k.mode = mode; k.B = B; k.T = Tb;
        Dim3 grid;
        grid.x = (int)(((long long)nxl * nyx * nzl * Rx_ + 255) / 256);
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    };
    // the four moves between the three layouts; `perm`: the Makhoul permutation of a Bounded dimension is applied / undone by this move
    auto z_to_y = [&](int perm) { for (int c = 0; c < C; ++c) all_to_all(fftbuf_, diststage_, c, C); transpose_y(true, perm); };
    auto y_to_z = [&](int perm) { transpose_y(false, perm); for (int c = 0; c < C; ++c) all_to_all(diststage_, fftbuf_, c, C); };
    auto pencil_yx = [&](bool pack, Cplx<FT>* B, Cplx<FT>* Tb) {        // modes 0 / 3 through a shared-memory tile
        PencilYXKernel<FT> k;
        k.nxl = nxl; k.nyx = nyx; k.nzl = nzl; k.Rx = Rx_; k.pack = pack ? 1 : 0; k.B = B; k.T = Tb;
        Dim3 grid;
        grid.x = (nxl + 31) / 32; grid.y = (nyx + 31) / 32; grid.z = nzl * Rx_;
        go(k, grid, PencilYXKernel<FT>::SMEM, OC_TIMER_POISSON_MID);
    };
    auto y_to_x = [&](int perm) { pencil_yx(true, spec, T); all_to_all_x(fftbuf_, diststage_); pencil_x(1, stage, T, perm); };
    auto x_to_y = [&](int perm) { pencil_x(2, stage, T, perm); all_to_all_x(diststage_, fftbuf_); pencil_yx(false, spec, T); };
    auto fft_y = [&](bool fwd) {
        if (!fwd && g_.bounded[1]) twiddle_lines(T, Ny, (long long)nzl * nxl, tw_[1], true);
        for (int c = 0; c < C; ++c) { begin_timer(OC_TIMER_FFT); std::string e = dfft_.y(distT_, fwd, c); end_timer(); chk(e); }
        if (fwd && g_.bounded[1]) twiddle_lines(T, Ny, (long long)nzl * nxl, tw_[1], false);
    };
    auto fft_x = [&](bool fwd) {
        if (!fwd && g_.bounded[0]) twiddle_lines(T, Nx, (long long)nzl * nyx, tw_[0], true);
        begin_timer(OC_TIMER_FFT); std::string e = dfft_.x(distT_, fwd); end_timer(); chk(e);
        if (fwd && g_.bounded[0]) twiddle_lines(T, Nx, (long long)nzl * nyx, tw_[0], false);
    };
    auto fft_z = [&](bool fwd) {
        if (!fwd) local_twiddles(true);
        begin_timer(OC_TIMER_FFT); std::string e = dfft_.z(fftbuf_, fwd); end_timer(); chk(e);
        if (fwd) local_twiddles(false);
    };
    const int py = g_.bounded[1], px = g_.bounded[0];
    if (stretched_) {
        // vertically stretched grid (distributed_fft_tridiagonal_solver.jl:262-293): y and x transforms, all the way back to whole
        // z-columns of this rank's share of the (kx, ky) plane (natural order: the moves in between carry no permutation), the tridiagonal
        // solves, and the same way out
        z_to_y(py); fft_y(true); y_to_x(px); fft_x(true);
        x_to_y(0); y_to_z(0);
        TridiagSolvePPKernel<FT> k;
        k.L = fft_.L;
        k.spec = fftbuf_;
        k.R = tri_R_; k.T = tri_T_; k.rdzf = g_.rdzf;
        k.norm = 1.0 / ((double)Nx * Ny);
        k.zero_col = (rank_ == 0 && rx_ == 0) ? 1 : 0;
        Dim3 grid;
        const long long n2 = 2LL * fft_.L.nxc * g_.N[1];
        grid.x = (int)((n2 + TridiagSolvePPKernel<FT>::THREADS - 1) / TridiagSolvePPKernel<FT>::THREADS);
        go(k, grid, 0, OC_TIMER_POISSON_MID);
        z_to_y(0); y_to_x(0);
        fft_x(false); x_to_y(px); fft_y(false); y_to_z(py);
        return;
    }
    fft_z(true);
    z_to_y(py); fft_y(true);
    y_to_x(px); fft_x(true);
    {
        PoissonDividePencilKernel<FT> k;
        k.Nx = Nx; k.nyx = nyx; k.nzl = nzl; k.y0 = rx_ * nyx; k.kz0 = rank_ * nzl;
        k.T = T;
        for (int d = 0; d < 3; ++d) k.lam[d] = lam_[d];
        k.norm = 1.0 / ((double)Nx * Ny * g_.N[2]);
        Dim3 grid;
        grid.x = (Nx + 255) / 256; grid.y = nyx; grid.z = nzl;
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    }
    fft_x(false); x_to_y(px);
    fft_y(false); y_to_z(py);
    fft_z(false);
}

template <class FT>
void Model<FT>::exchange_y(const std::vector<FieldRec*>& fields) {
    NvtxRange nvtx_("halo exchange (y)");
    if (!transport_) throw Error(OC_ERR_STATE, "distributed model without a transport: call oc_dist_attach_nccl first");
    const int nf = (int)fields.size();
    if (nf > F_) throw Error(OC_ERR_INVALID, "halo exchange of more fields than the exchange buffers hold");
    const int planes = (int)(field_elems_ / (size_t)g_.sz);
    HaloPackKernel<FT> k;
    k.g = g_;
    k.nfields = nf; k.planes = planes; k.rows = g_.H[1];
    for (int f = 0; f < nf; ++f) k.base[f] = fields[f]->base;
    const size_t per_side = (size_t)nf * planes * g_.H[1] * g_.sy;
    Dim3 grid;
    grid.x = (int)((2 * per_side + 255) / 256);
    k.unpack = 0; k.buf = halo_send_;
    go(k, grid, 0, OC_TIMER_COMM);
    const int prev = grank(rx_, (rank_ + R_ - 1) % R_), next = grank(rx_, (rank_ + 1) % R_);
    // low-edge rows go to prev (they are its high halo); high-edge rows go to next (its low halo).  Posting order: for R = 2 the
    // peer's first send (its low edge) must meet my first receive (my high halo).  A wall side (Bounded y) has no partner: the chain
    // of slabs is open, and that half of the message is empty.
    const size_t to_prev = g_.wlo[1] ? 0 : per_side * sizeof(FT), to_next = g_.whi[1] ? 0 : per_side * sizeof(FT);
    std::vector<Msg> msgs;
    msgs.push_back(Msg{prev, next, 0, halo_send_, to_prev, halo_recv_ + per_side, to_next});
    msgs.push_back(Msg{next, prev, 1, halo_send_ + per_side, to_next, halo_recv_, to_prev});
    begin_timer(OC_TIMER_COMM);
    std::string e = transport_->exchange(msgs, launch_stream_);
    end_timer();
    if (!e.empty()) throw Error(OC_ERR_CUDA, e);
    k.unpack = 1; k.buf = halo_recv_;
    go(k, grid, 0, OC_TIMER_COMM);
}

// all-to-all of sub-chunk c (of C) of the R equal chunks of a spectral buffer (distributed_transpose.jl:185-191)
template <class FT>
void Model<FT>::all_to_all(FT* send, FT* recv, int c, int C) {
    const size_t chunk = fft_.buffer_bytes / R_, sub = chunk / C, off = sub * c;
    std::vector<Msg> msgs;
    for (int d = 1; d < R_; ++d) {
        const int to = (rank_ + d) % R_, from = (rank_ + R_ - d) % R_;
        msgs.push_back(Msg{grank(rx_, to), grank(rx_, from), 2 + d + 16 * c, (char*)send + chunk * to + off, sub, (char*)recv + chunk * from + off, sub});
    }
    begin_timer(OC_TIMER_COMM);
#ifndef OC_HOSTSIM
    cuda_check(cudaMemcpyAsync((char*)recv + chunk * rank_ + off, (char*)send + chunk * rank_ + off, sub, cudaMemcpyDeviceToDevice, stream_), "cudaMemcpyAsync D2D");
#else
    memcpy((char*)recv + chunk * rank_ + off, (char*)send + chunk * rank_ + off, sub);
#endif
    std::string e = transport_ ? transport_->exchange(msgs, stream_) : std::string("distributed model without a transport");
    end_timer();
    if (!e.empty()) throw Error(OC_ERR_CUDA, e);
}

// FFT(z,x) local -> transpose -> FFT(y) -> divide -> FFT⁻¹(y) -> transpose back -> FFT⁻¹(z,x)
// (distributed_fft_based_poisson_solver.jl:141-178).  The y stage is independent per local z-level, so it runs in C sub-chunks
// on a second stream: the all-to-all of sub-chunk c+1 (NCCL, stream_) overlaps the transposes / y-FFTs / divide of sub-chunk c,
// and the way back starts as soon as a sub-chunk is finished.  (The reference does not overlap transposes with FFTs.)
template <class FT>
void Model<FT>::run_fft_solve_dist() {
    if (Rx_ > 1) { run_fft_solve_pencil(); return; }
    if (stretched_) { run_fft_solve_dist_tridiagonal(); return; }
    if (p2p_) { run_fft_solve_p2p(); return; }
    auto chk = [](const std::string& e) { if (!e.empty()) throw Error(OC_ERR_CUDA, e); };
    const int C = dfft_.C, nz = dfft_.Nzl / C;
    begin_timer(OC_TIMER_FFT); std::string e = dfft_.zx(fftbuf_, true); end_timer(); chk(e);
    local_twiddles(false);    // DCT-II post-twiddles while x and z are still local
    for (int c = 0; c < C; ++c) {
        all_to_all(fftbuf_, diststage_, c, C);
#ifndef OC_HOSTSIM
        cuda_check(cudaEventRecord((cudaEvent_t)ev_a2a_[c], stream_), "cudaEventRecord");
#endif
    }
    for (int c = 0; c < C; ++c) {
#ifndef OC_HOSTSIM
        cuda_check(cudaStreamWaitEvent(stream3_, (cudaEvent_t)ev_a2a_[c], 0), "cudaStreamWaitEvent");
        launch_stream_ = stream3_;
#endif
        TransposeKernel<FT> t;
        t.nxc = dfft_.nxc; t.nyl = g_.N[1]; t.nzl = dfft_.Nzl; t.R = R_; t.zl0 = c * nz; t.yperm = g_.bounded[1];
        t.stage = reinterpret_cast<Cplx<FT>*>(diststage_); t.T = reinterpret_cast<Cplx<FT>*>(distT_);
        Dim3 tg;
        tg.x = (dfft_.nxc + 31) / 32; tg.y = (dfft_.Ny + 31) / 32; tg.z = nz;
        t.to_T = 1;
        go(t, tg, TransposeKernel<FT>::SMEM, OC_TIMER_POISSON_MID);
        begin_timer(OC_TIMER_FFT); e = dfft_.y(distT_, true, c); end_timer(); chk(e);
        PoissonDivideTKernel<FT> k;
        k.nxc = dfft_.nxc; k.ny = dfft_.Ny; k.nzl = dfft_.Nzl; k.kz0 = rank_ * dfft_.Nzl; k.zl0 = c * nz;
        k.T = reinterpret_cast<Cplx<FT>*>(distT_);
        for (int d = 0; d < 3; ++d) k.lam[d] = lam_[d];
        k.tw = g_.bounded[1] ? tw_[1] : nullptr;
        k.norm = 1.0 / ((double)g_.N[0] * dfft_.Ny * g_.N[2]);
        Dim3 grid;
        grid.x = (dfft_.Ny + 255) / 256; grid.y = dfft_.nxc; grid.z = nz;
        go(k, grid, 0, OC_TIMER_POISSON_MID);
        begin_timer(OC_TIMER_FFT); e = dfft_.y(distT_, false, c); end_timer(); chk(e);
        t.to_T = 0;
        go(t, tg, TransposeKernel<FT>::SMEM, OC_TIMER_POISSON_MID);
#ifndef OC_HOSTSIM
        cuda_check(cudaEventRecord((cudaEvent_t)ev_mid_[c], stream3_), "cudaEventRecord");
        launch_stream_ = stream_;
        cuda_check(cudaStreamWaitEvent(stream_, (cudaEvent_t)ev_mid_[c], 0), "cudaStreamWaitEvent");
#endif
        all_to_all(diststage_, fftbuf_, c, C);
    }
    local_twiddles(true);     // DCT-III pre-twiddles, x and z local again
    begin_timer(OC_TIMER_FFT); e = dfft_.zx(fftbuf_, false); end_timer(); chk(e);
}

template <class FT>
void Model<FT>::fill_halo_regions(const int* fields, int n, int fill_open) {
    std::vector<FieldRec*> list;
    for (int i = 0; i < n; ++i) list.push_back(&lookup(fields[i]));
    halo(list, fill_open != 0);
}

// ---------------------------------------------------------------------------------------------------------
// auxiliary fields: AMD diffusivities (+ their halos), hydrostatic pressure   (compute_auxiliaries! :58-69)
// ---------------------------------------------------------------------------------------------------------
template <class FT>
void Model<FT>::aux() {
    NvtxRange nvtx_("compute_auxiliaries!");
    if (has_amd_) {
        auto run_amd = [&](auto k) {
            k.g = g_;
            k.u = state_[0].p; k.v = state_[1].p; k.w = state_[2].p;
            k.nu_e = nu_e_.p;
            k.Cnu = (FT)cfg_.amd_Cnu;
            k.ntr = cfg_.n_tracers;
            for (int t = 0; t < cfg_.n_tracers; ++t) { k.c[t] = state_[3 + t].p; k.kappa_e[t] = kappa_e_[t].p; k.Ckappa[t] = (FT)cfg_.amd_Ckappa[t]; }
            k.set_consts();
            const size_t nt = (size_t)g_.N[2] + 2 * (g_.H[2] + 1) + 1;
            const FT* at = stretched_ ? g_.rVf + nt : nullptr;         // the AMD tables follow the six metric tables (build_z_tables)
            k.lv_kxw = at; k.lv_kzu = at ? at + nt : nullptr; k.lv_kyw = at ? at + 2 * nt : nullptr; k.lv_kzv = at ? at + 3 * nt : nullptr;
            k.lv_kcz = at ? at + 4 * nt : nullptr; k.lv_d2 = at ? at + 5 * nt : nullptr;
            k.Cb = (FT)cfg_.amd_Cb;
            k.buoyancy = cfg_.buoyancy;
            k.bT = k.bS = nullptr;
            if (cfg_.buoyancy == OC_BUOYANCY_TRACER) k.bT = state_[3 + cfg_.tracer_b].p;
            else if (cfg_.buoyancy == OC_BUOYANCY_SEAWATER_LINEAR) { k.bT = state_[3 + cfg_.tracer_T].p; k.bS = state_[3 + cfg_.tracer_S].p; }
            k.grav = (FT)cfg_.gravity; k.alpha = (FT)cfg_.thermal_expansion; k.beta = (FT)cfg_.haline_contraction;
            Dim3 ag;
            ag.x = (g_.N[0] + 31) / 32; ag.y = (g_.N[1] + 7) / 8; ag.z = g_.N[2];
            go(k, ag, 0, OC_TIMER_AUX);
        };
        // the buoyancy modification (amd_has_Cb) is a separate instantiation: without buoyancy ∂b = 0 and the plain kernel is exact
        const bool cb = cfg_.amd_has_Cb && cfg_.buoyancy != OC_BUOYANCY_NONE;
        if (cb) { if (stretched_) run_amd(AmdKernel<FT, true, true>{}); else run_amd(AmdKernel<FT, false, true>{}); }
        else if (stretched_) run_amd(AmdKernel<FT, true>{});
        else run_amd(AmdKernel<FT, false>{});
        std::vector<FieldRec*> list{&nu_e_};
        for (auto& f : kappa_e_) list.push_back(&f);
        halo(list, true);
    }
    if (has_smag_) {
        auto run_smag = [&](auto k) {
            k.g = g_;
            k.u = state_[0].p; k.v = state_[1].p; k.w = state_[2].p;
            k.nu_e = nu_e_.p;
            k.ntr = cfg_.n_tracers;
            for (int t = 0; t < cfg_.n_tracers; ++t) { k.kappa_e[t] = kappa_e_[t].p; k.rPr[t] = FT(1) / (FT)cfg_.smag_Pr[t]; }
            const FT C = (FT)cfg_.smag_C;
            k.cs2 = C * C;
            k.lilly = cfg_.smagorinsky == 2;
            k.Cb = (FT)cfg_.smag_Cb;
            k.buoyancy = cfg_.buoyancy;
            k.bT = k.bS = nullptr;
            if (cfg_.buoyancy == OC_BUOYANCY_TRACER) k.bT = state_[3 + cfg_.tracer_b].p;
            else if (cfg_.buoyancy == OC_BUOYANCY_SEAWATER_LINEAR) { k.bT = state_[3 + cfg_.tracer_T].p; k.bS = state_[3 + cfg_.tracer_S].p; }
            k.grav = (FT)cfg_.gravity; k.alpha = (FT)cfg_.thermal_expansion; k.beta = (FT)cfg_.haline_contraction;
            if (stretched_) {
                const size_t nt = (size_t)g_.N[2] + 2 * (g_.H[2] + 1) + 1;
                k.lv_df2 = g_.rVf + nt + 6 * nt;          // after the six metric tables and the six AMD tables (build_z_tables)
                k.df2 = FT(0);
            } else {
                const FT df = std::cbrt(g_.d[0] * g_.d[1] * g_.d[2]);
                k.df2 = df * df;
                k.lv_df2 = nullptr;
            }
            Dim3 ag;
            ag.x = (g_.N[0] + 31) / 32; ag.y = (g_.N[1] + 7) / 8; ag.z = g_.N[2];
            go(k, ag, 0, OC_TIMER_AUX);
        };
        if (stretched_) run_smag(SmagorinskyKernel<FT, true>{});
        else run_smag(SmagorinskyKernel<FT, false>{});
        std::vector<FieldRec*> list{&nu_e_};
        for (auto& f : kappa_e_) list.push_back(&f);
        halo(list, true);
    }
    hydrostatic_pressure();
    aux_valid_ = true;
}

// update_hydrostatic_pressure!  (update_hydrostatic_pressure.jl:12-49): one column scan, launched on launch_stream_
template <class FT>
void Model<FT>::hydrostatic_pressure() {
    hydrostatic_pressure_rows(g_.flat[1] ? 0 : -1, g_.flat[1] ? g_.N[1] : g_.N[1] + 2);
}

// … for the rows jlo … jlo + nj - 1 (the reference's range is -1 … Ny, p_kernel_parameters :41-49; the distributed split computes the
// two halo rows after the exchange that fills T and S there)
template <class FT>
void Model<FT>::hydrostatic_pressure_rows(int jlo, int nj) {
    if (has_pHY_ && !g_.flat[2]) {
        HydrostaticPressureKernel<FT> k;
        k.g = g_;
        k.pHY = pHY_.p;
        k.buoyancy = cfg_.buoyancy;
        if (cfg_.buoyancy == OC_BUOYANCY_TRACER) { k.bT = state_[3 + cfg_.tracer_b].p; k.bS = nullptr; }
        else { k.bT = state_[3 + cfg_.tracer_T].p; k.bS = state_[3 + cfg_.tracer_S].p; }
        k.grav = (FT)cfg_.gravity; k.alpha = (FT)cfg_.thermal_expansion; k.beta = (FT)cfg_.haline_contraction;
        k.tilted = cfg_.tilted_gravity ? 1 : 0;
        k.gz = -(FT)cfg_.gravity_unit_vector[2];
        k.ilo = g_.flat[0] ? 0 : -1; k.ni = g_.flat[0] ? g_.N[0] : g_.N[0] + 2;
        k.jlo = jlo; k.nj = nj;
        Dim3 grid;
        grid.x = (k.ni + HydrostaticPressureKernel<FT>::THREADS - 1) / HydrostaticPressureKernel<FT>::THREADS;
        grid.y = k.nj;
        go(k, grid, 0, OC_TIMER_AUX);
    }
}

// ---------------------------------------------------------------------------------------------------------
// tendencies (+ fused substep)
// ---------------------------------------------------------------------------------------------------------
template <class FT>
template <int KIND>
void Model<FT>::launch_tendency(int fidx, TendencyArgs<FT>& a) {
    (void)fidx;
    Dim3 grid;
    constexpr int TX = 32, TY = 8, TZ = 8;
    grid.x = (g_.N[0] + TX - 1) / TX;
    grid.y = (g_.N[1] + TY - 1) / TY;
    grid.z = (g_.N[2] + TZ - 1) / TZ;
    auto run = [&](auto k) {
        k.a = a;
        k.cor.beta = (FT)cfg_.coriolis_beta;
        // south face of this rank's first row (slab decomposition in y: rank r owns rows r·Ny … (r+1)·Ny − 1)
        k.cor.y0 = (FT)cfg_.origin_y + (FT)((dist_ ? rank_ : 0) * g_.N[1]) * g_.d[1];
        for (int d = 0; d < 3; ++d) k.cor.cf[d] = (FT)cfg_.coriolis_fxyz[d];
        k.cor.gamma = (FT)cfg_.coriolis_gamma; k.cor.R = (FT)cfg_.coriolis_radius; k.cor.z0 = (FT)cfg_.origin_z;
        {
            const size_t nt = (size_t)g_.N[2] + 2 * (g_.H[2] + 1) + 1;
            k.cor.zc = stretched_ ? g_.rVf + nt + 7 * nt : nullptr;       // after the metric, AMD and Smagorinsky tables (build_z_tables)
            k.cor.zf = stretched_ ? g_.rVf + nt + 8 * nt : nullptr;
        }
        k.cor.tilted = (cfg_.tilted_gravity && cfg_.buoyancy != OC_BUOYANCY_NONE) ? 1 : 0;
        for (int d = 0; d < 3; ++d) k.cor.gh[d] = -(FT)cfg_.gravity_unit_vector[d];
        k.cor.tb_kind = cfg_.buoyancy;
        k.cor.tbT = k.cor.tbS = nullptr;
        if (cfg_.buoyancy == OC_BUOYANCY_TRACER) k.cor.tbT = state_[3 + cfg_.tracer_b].p;
        else if (cfg_.buoyancy == OC_BUOYANCY_SEAWATER_LINEAR) { k.cor.tbT = state_[3 + cfg_.tracer_T].p; k.cor.tbS = state_[3 + cfg_.tracer_S].p; }
        go(k, grid, k.SMEM, OC_TIMER_TENDENCY);
    };
    if (cfg_.has_advection_dir) { run(TendencyKernel<FT, ADV_MIXED, KIND, TX, TY, TZ>{}); return; }
    switch (cfg_.advection) {
        case OC_WENO5: run(TendencyKernel<FT, ADV_WENO5, KIND, TX, TY, TZ>{}); break;
        case OC_CENTERED4: run(TendencyKernel<FT, ADV_CENTERED4, KIND, TX, TY, TZ>{}); break;
        case OC_UPWIND3: run(TendencyKernel<FT, ADV_UPWIND3, KIND, TX, TY, TZ>{}); break;
        case OC_UPWIND5: run(TendencyKernel<FT, ADV_UPWIND5, KIND, TX, TY, TZ>{}); break;
        case OC_WENO3: run(TendencyKernel<FT, ADV_WENO3, KIND, TX, TY, TZ>{}); break;
        case OC_UPWIND1: run(TendencyKernel<FT, ADV_UPWIND1, KIND, TX, TY, TZ>{}); break;
        case OC_WENO7: run(TendencyKernel<FT, ADV_WENO7, KIND, TX, TY, TZ>{}); break;
        case OC_WENO9: run(TendencyKernel<FT, ADV_WENO9, KIND, TX, TY, TZ>{}); break;
        case OC_ADVECTION_NONE: run(TendencyKernel<FT, ADV_NONE, KIND, TX, TY, TZ>{}); break;
        default: run(TendencyKernel<FT, ADV_CENTERED2, KIND, TX, TY, TZ>{}); break;
    }
}

// ---------------------------------------------------------------------------------------------------------
// z-marching TMA kernel (oc_march.h)
// ---------------------------------------------------------------------------------------------------------
#ifndef OC_HOSTSIM
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cuda_check(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q), "cudaGetDriverEntryPoint");
        if (q != cudaDriverEntryPointSuccess || !p) throw Error(OC_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}
#endif

// TMA descriptor of a whole field allocation (x fastest; rows sy, planes sz) with a bx × by × 1 box
template <class FT>
TileSrc<FT> Model<FT>::tile_src(const FT* base, int bx, int by) {
    const int rows = g_.sz / g_.sy, planes = (int)(field_elems_ / (size_t)g_.sz);
#ifndef OC_HOSTSIM
    auto key = std::make_tuple((const void*)base, bx, by);
    auto it = tmap_cache_.find(key);
    if (it != tmap_cache_.end()) return it->second;
    TileSrc<FT> t;
    cuuint64_t dims[3] = {(cuuint64_t)g_.sy, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)g_.sy * sizeof(FT), (cuuint64_t)g_.sz * sizeof(FT)};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_tiled_fn()(&t.map, sizeof(FT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                   (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(OC_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    tmap_cache_.emplace(key, t);
    return t;
#else
    (void)bx; (void)by;
    TileSrc<FT> t;
    t.base = base;
    t.dim[0] = g_.sy; t.dim[1] = rows; t.dim[2] = planes;
    t.stride[0] = 1; t.stride[1] = g_.sy; t.stride[2] = g_.sz;
    return t;
#endif
}

template <class FT>
template <int KIND>
void Model<FT>::launch_march_tendency(int fidx, TendencyArgs<FT>& a, int part) {
    auto run = [&](auto k) {
        using K = decltype(k);
        using SP = typename K::SP;
        constexpr int TX = K::TX, TY = K::TY;
        k.a = a;
        k.xpad = xpad_;
        const FT* self = state_[fidx].base;
        auto vel_base = [&](int c) -> const FT* { return state_[c].base; };
        k.src[0] = tile_src(self, SP::R0::BX, SP::R0::BY);
        k.src[1] = tile_src(vel_base(SP::F1), SP::R1::BX, SP::R1::BY);
        k.src[2] = tile_src(vel_base(SP::F2), SP::R2::BX, SP::R2::BY);
        if (SP::NR > 3) k.src[3] = tile_src(vel_base(SP::F3 < 0 ? 0 : SP::F3), SP::R3::BX, SP::R3::BY);
        else k.src[3] = k.src[0];
        Dim3 grid;
        grid.x = (g_.N[0] + TX - 1) / TX;
        grid.y = (g_.N[1] + TY - 1) / TY;
        // z chunks: enough CTAs to fill the machine several times over, but chunks of at least 16 levels
        const int tiles = grid.x * grid.y;
        int zch = (148 * 8 + tiles - 1) / tiles;
        zch = std::max(1, std::min(zch, (g_.N[2] + 15) / 16));
        k.KC = (g_.N[2] + zch - 1) / zch;
        grid.z = (g_.N[2] + k.KC - 1) / k.KC;
        // Kernels that overlap the pressure solve (stream2) leave a third of every SM free — two CTAs per SM instead of three, by
        // asking for more shared memory than they use — so that the FFT passes and the NCCL kernels can be co-resident.
        size_t smem = K::SMEM;
        if (launch_stream_ != stream_ && smem < 80 * 1024) smem = 80 * 1024;
        // tile rows of this launch: everything, the interior (no stencil of tile rows 1 … NBY-2 reaches a y-halo row: TY >= 8 > 3),
        // or the two boundary strips.  With fewer than three tile rows there is no interior: the strips are everything.
        const int NBY = grid.y;
        int rows[2][2] = {{0, NBY}, {0, 0}};
        if (part == PART_INTERIOR) { rows[0][0] = 1; rows[0][1] = NBY >= 3 ? NBY - 2 : 0; }
        else if (part == PART_STRIPS && NBY >= 3) { rows[0][0] = 0; rows[0][1] = 1; rows[1][0] = NBY - 1; rows[1][1] = 1; }
        for (int r = 0; r < 2; ++r) {
            if (rows[r][1] <= 0) continue;
            k.by0 = rows[r][0];
            grid.y = rows[r][1];
            begin_timer(OC_TIMER_TENDENCY);
            cudaError_t e = replay_ ? cudaSuccess : launch_march(k, grid, smem, launch_stream_);
            end_timer();
#ifndef OC_HOSTSIM
            cuda_check(e, "march kernel launch");
#else
            (void)e;
#endif
            ++launches;
        }
    };
    const bool bnd = g_.bounded[0] || g_.bounded[1] || g_.bounded[2];
    const bool gen = has_eddy_;
    // Two cells per thread (32×16 tiles) for the u, v and tracer kernels of the triply periodic WENO(5) configurations (C3 / C5).
    // Measured (profiles/r01g_two_cells_per_thread.txt): tendency time per step 42.1 -> 40.9 ms (C3 Float64), 31.6 -> 28.5 ms (C3
    // Float32); but Centered(2) 1.89 -> 2.02 ms (C2) and the Bounded-z AMD kernels 31.7 -> 36.8 ms (C4, 83-92 registers, 18 instead of
    // 27 warps per SM to hide their global loads), UpwindBiased(5) 29.3 -> 32.5 ms — so those stay on one cell per thread: with little
    // FP64 work per cell the kernels need the third CTA's warps more than the saved instructions.  OC_MARCH_CPT=1 switches back (measurement).
    static const char* cpt_env = getenv("OC_MARCH_CPT");
    const bool two_cells = cpt_env ? atoi(cpt_env) == 2 : true;
    // (A 32×16-tile variant — MarchKernel<…, 16>: two 544-thread CTAs per SM, 34 warps, 50 registers — was measured slower,
    //  48.3 vs 45.4 ms per step at 512³: the kernel is bound by the FP64 pipe and dependent-issue latency, not by warp count.)
    auto pick = [&](auto adv) {
        constexpr int ADV = decltype(adv)::value;
        const bool zonly = !g_.bounded[0] && !g_.bounded[1] && g_.bounded[2];     // the LES topology (Periodic, Periodic, Bounded)
        if constexpr (ADV == ADV_UPWIND5) {
            run(MarchKernel<FT, ADV, KIND, 0, 0>{});       // march_ok_ admits only the triply periodic constant-viscosity case
        } else {
        if (stretched_) {          // z Bounded and variably spaced: level tables instead of the constant z metrics
            if (zonly && !gen) run(MarchKernel<FT, ADV, KIND, 4, 0, 8, 1>{});
            else if (zonly) run(MarchKernel<FT, ADV, KIND, 4, 1, 8, 1>{});
            else if (!gen) run(MarchKernel<FT, ADV, KIND, 7, 0, 8, 1>{});
            else run(MarchKernel<FT, ADV, KIND, 7, 1, 8, 1>{});
        }
        else if (two_cells && !bnd && !gen) {
            if constexpr (KIND != KIND_W && ADV == ADV_WENO5) run(MarchKernel<FT, ADV, KIND, 0, 0, 16>{});
            else run(MarchKernel<FT, ADV, KIND, 0, 0>{});
        }
        else if (!bnd && !gen) run(MarchKernel<FT, ADV, KIND, 0, 0>{});
        else if (!bnd) run(MarchKernel<FT, ADV, KIND, 0, 1>{});
        else if (zonly && !gen) run(MarchKernel<FT, ADV, KIND, 4, 0>{});
        else if (zonly) run(MarchKernel<FT, ADV, KIND, 4, 1>{});
        else if (!gen) run(MarchKernel<FT, ADV, KIND, 7, 0>{});
        else run(MarchKernel<FT, ADV, KIND, 7, 1>{});
        }
    };
    if (cfg_.advection == OC_WENO5) pick(std::integral_constant<int, ADV_WENO5>{});
    else if (cfg_.advection == OC_UPWIND5) pick(std::integral_constant<int, ADV_UPWIND5>{});
    else pick(std::integral_constant<int, ADV_CENTERED2>{});
}

// one launch for u, v and w (oc_uvw.h): Centered(2), no Bounded dimension, constant ν
template <class FT>
void Model<FT>::launch_uvw(const TendencyArgs<FT>& au, const TendencyArgs<FT>& av, const TendencyArgs<FT>& aw) {
    typedef UvwCenteredKernel<FT> K;
    K k;
    k.a.g = g_;
    k.a.pHY = au.pHY;
    const TendencyArgs<FT>* as[3] = {&au, &av, &aw};
    for (int c = 0; c < 3; ++c) { k.a.Gm[c] = as[c]->Gm; k.a.Gn[c] = as[c]->Gn; k.a.Unew[c] = as[c]->Unew; }
    k.a.nu = au.has_scalar ? au.nu : FT(0);
    k.a.has_coriolis = au.has_coriolis; k.a.f = au.f; k.a.cor_beta = au.cor_beta; k.a.cor_y0 = au.cor_y0;
    k.a.mode = au.mode; k.a.dt = au.dt; k.a.ca = au.ca; k.a.cb = au.cb; k.a.ab2_euler = au.ab2_euler;
    for (int c = 0; c < 3; ++c) k.src[c] = tile_src(state_[c].base, K::RS::BX, K::RS::BY);
    k.xpad = xpad_;
    k.by0 = 0;
    Dim3 grid;
    grid.x = (g_.N[0] + K::TX - 1) / K::TX;
    grid.y = (g_.N[1] + K::TY - 1) / K::TY;
    const int tiles = grid.x * grid.y;
    int zch = (148 * 8 + tiles - 1) / tiles;
    zch = std::max(1, std::min(zch, (g_.N[2] + 15) / 16));
    k.KC = (g_.N[2] + zch - 1) / zch;
    grid.z = (g_.N[2] + k.KC - 1) / k.KC;
    begin_timer(OC_TIMER_TENDENCY);
    cudaError_t e = replay_ ? cudaSuccess : launch_march(k, grid, K::SMEM, launch_stream_);
    end_timer();
#ifndef OC_HOSTSIM
    cuda_check(e, "uvw kernel launch");
#else
    (void)e;
#endif
    ++launches;
}

template <class FT>
void Model<FT>::tendencies(int mode, double dt, int stage, double chi, bool euler, bool add_flux_bcs, bool swap_state, bool defer_tracer_join) {
    NvtxRange nvtx_("compute_tendencies! + substep");
    join_tracers();
    // The hydrostatic-pressure scan (HBM-bound, needed by the u and v kernels only) runs on the second stream beside the w and tracer
    // tendency kernels (bound by the FP64 pipe / instruction issue, two ~100 KB CTAs per SM: its small register-only CTAs fit into what
    // they leave free); u and v are launched last and wait for it.  OC_PHY_ASYNC=0 switches back (measurement).
    static const char* phy_env = getenv("OC_PHY_ASYNC");
    bool phy_async = false;
    // Distributed models: the previous stage's y-halo exchange may still be in flight on the communication stream (halo(): defer_exchange).
    // Then every field's INTERIOR tile rows are launched first — none of their stencils reaches a y-halo row —, the main stream (and the
    // tracer stream) wait for the exchange, and the two boundary strips follow (interleave_communication_and_computation.jl:29-67,
    // compute_nonhydrostatic_buffer_tendencies.jl:10-84).  pHY′: rows 0 … Ny-1 now, the two halo rows after the exchange.  Closures with
    // eddy-viscosity fields read halo rows in aux(): no split for them.
    const bool split = xchg_pending_ && march_ok_ && !has_eddy_ && Rx_ == 1;
    if (xchg_pending_ && !split) join_exchange();
#ifndef OC_HOSTSIM
    phy_async = !aux_valid_ && has_pHY_ && !g_.flat[2] && !has_eddy_ && march_ok_ && F_ > 3 && (phy_env ? atoi(phy_env) != 0 : true);
    if (phy_async) {
        if (!ev_phy_) { cudaEvent_t e; cuda_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate"); ev_phy_ = e; }
        if (!replay_) {
            cuda_check(cudaEventRecord((cudaEvent_t)ev_fork_, stream_), "cudaEventRecord");
            cuda_check(cudaStreamWaitEvent(stream2_, (cudaEvent_t)ev_fork_, 0), "cudaStreamWaitEvent");
        }
        launch_stream_ = stream2_;
        if (split) hydrostatic_pressure_rows(0, g_.N[1]); else hydrostatic_pressure();
        launch_stream_ = stream_;
        if (!replay_) cuda_check(cudaEventRecord((cudaEvent_t)ev_phy_, stream2_), "cudaEventRecord");
        aux_valid_ = true;
    }
#else
    (void)phy_env;
#endif
    if (split && has_pHY_ && !g_.flat[2] && !aux_valid_) { hydrostatic_pressure_rows(0, g_.N[1]); aux_valid_ = true; }
    const bool do_split = split && xchg_pending_;
    if (!aux_valid_) aux();
    // Measured (profiles/): on one GPU the co-residency costs the tracer kernels more than the overlap wins (69.4 vs 67.5 ms);
    // across GPUs it hides part of the NCCL transposes (77.1 vs 79.4 ms at 2 GPUs) — so it is on for distributed models only.
    static const char* ov_env = getenv("OC_OVERLAP");
    const bool want = ov_env ? atoi(ov_env) != 0 : dist_;
    const bool overlap = defer_tracer_join && F_ > 3 && march_ok_ && want;
    auto make_args = [&](int f) {
        TendencyArgs<FT> a;
        memset(&a, 0, sizeof(a));
        a.g = g_;
        a.C = C_;
        for (int d = 0; d < 3; ++d) a.U[d] = state_[d].p;
        a.c = f >= 3 ? state_[f].p : nullptr;
        a.pHY = (has_pHY_ && !g_.flat[2]) ? pHY_.p : nullptr;
        a.bT = a.bS = nullptr;
        a.buoyancy = 0;     // the w-equation gets no buoyancy term when pHY′ exists (always, when buoyancy != nothing)
        a.nu_e = (has_eddy_ && f < 3) ? nu_e_.p : nullptr;
        a.kappa_e = (has_eddy_ && f >= 3) ? kappa_e_[f - 3].p : nullptr;
        a.Gm = Gm_[f].p;
        a.Gn = Gn_[f].p;
        a.Ucur = state_[f].p;
        a.Unew = next_[f].p;
        a.has_scalar = cfg_.has_scalar_diffusivity;
        a.nu = (FT)cfg_.nu;
        a.kappa = f >= 3 ? (FT)cfg_.kappa[f - 3] : FT(0);
        a.grav = (FT)cfg_.gravity; a.alpha = (FT)cfg_.thermal_expansion; a.beta = (FT)cfg_.haline_contraction;
        a.has_coriolis = cfg_.has_coriolis;
        a.f = (FT)cfg_.coriolis_f;
        a.cor_beta = (FT)cfg_.coriolis_beta;
        a.cor_y0 = (FT)cfg_.origin_y + (FT)((dist_ ? rank_ : 0) * g_.N[1]) * g_.d[1];
        for (int s = 0; s < 6; ++s) {
            const SideBC& bc = state_[f].bc[s];
            const oc_bc& ub = cfg_.bcs[f][s];
            a.fbc.on[s] = (bc.kind == OC_BC_FLUX && ub.kind == OC_BC_FLUX && ub.has_value && !bc_array_[f][s]) ? 1 : 0;
            a.fbc.val[s] = (FT)bc.value;
        }
        a.add_flux_bcs = add_flux_bcs ? 1 : 0;
        for (int d = 0; d < 3; ++d) a.adv_dir[d] = cfg_.has_advection_dir ? cfg_.advection_dir[d] : cfg_.advection;
        a.mode = mode;
        a.dt = (FT)dt;
        a.ab2_euler = euler ? 1 : 0;
        if (mode == STEP_RK3_FIRST) { a.ca = (FT)dt * gamma_[0]; a.cb = FT(0); }
        else if (mode == STEP_RK3) { a.ca = gamma_[stage - 1]; a.cb = zeta_[stage - 1]; }
        else if (mode == STEP_AB2) { a.ca = FT(1.5) + (FT)chi; a.cb = FT(0.5) + (FT)chi; }
        return a;
    };
    // launch order when the pHY′ scan is in flight: the kernels that do not read it first — w and the tracers, then u and v; with the
    // distributed overlap (tracers beside the pressure solve, after the velocities) only w: w, u, v, tracers
    auto field_of = [&](int n) { return !phy_async ? n : (overlap ? (n == 0 ? 2 : (n < 3 ? n - 1 : n)) : (n < F_ - 2 ? n + 2 : n - (F_ - 2))); };
    bool phy_pending = phy_async;
    auto launch_pass = [&](int part) {
        int momentum_seen = 0;
        for (int n = 0; n < F_; ++n) {
            const int f = field_of(n);
            if (uvw_ok_ && f < 3) {            // u, v, w in ONE launch, issued where the last of the three would have been
                if (++momentum_seen < 3) continue;
                if (phy_pending) {
#ifndef OC_HOSTSIM
                    if (!replay_) cuda_check(cudaStreamWaitEvent(stream_, (cudaEvent_t)ev_phy_, 0), "cudaStreamWaitEvent");
#endif
                    phy_pending = false;
                }
                TendencyArgs<FT> au = make_args(0), av = make_args(1), aw = make_args(2);
                launch_uvw(au, av, aw);
                continue;
            }
            if (f >= 3 && overlap) { if (!tracers_in_flight_) fork_tracers(); }
#ifndef OC_HOSTSIM
            if (overlap) launch_stream_ = f >= 3 ? stream2_ : stream_;
#endif
            if (f < 2 && phy_pending) {
#ifndef OC_HOSTSIM
                if (!replay_) cuda_check(cudaStreamWaitEvent(stream_, (cudaEvent_t)ev_phy_, 0), "cudaStreamWaitEvent");
#endif
                phy_pending = false;
            }
            TendencyArgs<FT> a = make_args(f);
            if (march_ok_) {
                if (f == 0) launch_march_tendency<KIND_U>(f, a, part);
                else if (f == 1) launch_march_tendency<KIND_V>(f, a, part);
                else if (f == 2) launch_march_tendency<KIND_W>(f, a, part);
                else launch_march_tendency<KIND_C>(f, a, part);
            } else if (f == 0) launch_tendency<KIND_U>(f, a);
            else if (f == 1) launch_tendency<KIND_V>(f, a);
            else if (f == 2) launch_tendency<KIND_W>(f, a);
            else launch_tendency<KIND_C>(f, a);
            if (add_flux_bcs && part != PART_INTERIOR) {          // array-valued Flux BCs: boundary-plane pass on the same stream (FluxArrayKernel)
                const FT coef = mode == STEP_RK3_FIRST ? a.ca : ((mode == STEP_RK3 || mode == STEP_AB2) ? a.dt * a.ca : FT(0));
                apply_flux_arrays(f, a.Gn, mode == STEP_NONE ? nullptr : a.Unew, coef);
            }
        }
    };
    if (!do_split) {
        launch_pass(PART_ALL);
    } else {
        launch_pass(PART_INTERIOR);
#ifndef OC_HOSTSIM
        launch_stream_ = stream_;
        cuda_check(cudaStreamWaitEvent(stream_, (cudaEvent_t)ev_xchg_, 0), "cudaStreamWaitEvent");
        if (tracers_in_flight_) cuda_check(cudaStreamWaitEvent(stream2_, (cudaEvent_t)ev_xchg_, 0), "cudaStreamWaitEvent");
#endif
        xchg_pending_ = false;
        if (has_pHY_ && !g_.flat[2]) {           // the two halo rows of pHY′ (T, S are there now); the v strip at j = 0 reads row -1
            hydrostatic_pressure_rows(-1, 1);
            hydrostatic_pressure_rows(g_.N[1], 1);
        }
        launch_pass(PART_STRIPS);
    }
    launch_stream_ = stream_;
    if (swap_state && mode != STEP_NONE)
        for (int f = 0; f < F_; ++f) std::swap(state_[f].p, next_[f].p), std::swap(state_[f].base, next_[f].base);
}

// cache_previous_tendencies! by pointer swap (quasi_adams_bashforth_2.jl:116-120, runge_kutta_3.jl:131,147): the fused stage leaves its
// evaluation in the Gⁿ slot (gn_pending_); whoever evaluates next — a fused stage, update_state!(compute_tendencies = true), a read of
// Gⁿ / G⁻ — first rotates it into the G⁻ slot.  Exactly one rotation per consumed evaluation, whatever the caller interleaves.
template <class FT>
void Model<FT>::rotate_pending_tendencies() {
    if (!gn_pending_) return;
    for (int f = 0; f < F_; ++f) std::swap(Gn_[f].p, Gm_[f].p), std::swap(Gn_[f].base, Gm_[f].base);
    gn_pending_ = false;
}

template <class FT>
void Model<FT>::compute_tendencies_if_stale() {
    if (tend_valid_ && !gn_pending_) return;
    compute_tendencies();
}

template <class FT>
void Model<FT>::compute_tendencies() {
    rotate_pending_tendencies();
    tendencies(STEP_NONE, 0.0, 1, 0.0, false, false, false);
    tend_valid_ = true;
}

template <class FT>
void Model<FT>::update_state(int compute_tend) {
    std::vector<FieldRec*> list;
    for (auto& f : state_) list.push_back(&f);
    halo(list, false);                                   // fill_open_bcs = false   update_nonhydrostatic_model_state.jl:34
    aux();
    if (compute_tend) compute_tendencies();
}

template <class FT>
void Model<FT>::compute_flux_bc_tendencies() {
    for (int f = 0; f < F_; ++f) {
        FluxBCKernel<FT> k;
        k.g = g_;
        k.Gn = Gn_[f].p;
        k.zface = f == 2 ? 1 : 0;
        bool any = false;
        for (int s = 0; s < 6; ++s) {
            const oc_bc& ub = cfg_.bcs[f][s];
            k.fbc.on[s] = (state_[f].bc[s].kind == OC_BC_FLUX && ub.kind == OC_BC_FLUX && ub.has_value && !bc_array_[f][s]) ? 1 : 0;
            k.fbc.val[s] = (FT)state_[f].bc[s].value;
            any = any || k.fbc.on[s];
        }
        if (any) go(k, grid_xyz(256), 0, OC_TIMER_SUBSTEP);
        apply_flux_arrays(f, Gn_[f].p, nullptr, FT(0));
    }
}

// FluxBoundaryCondition(array): one boundary-plane launch per array-valued side of field f
template <class FT>
void Model<FT>::apply_flux_arrays(int f, FT* Gn, FT* Unew, FT coef) {
    for (int s = 0; s < 6; ++s) {
        if (!bc_array_[f][s] || state_[f].bc[s].kind != OC_BC_FLUX) continue;
        FluxArrayKernel<FT> k;
        k.g = g_;
        k.Gn = Gn; k.Unew = Unew; k.coef = coef;
        k.J = bc_array_[f][s];
        k.d = s / 2; k.side = s % 2;
        k.comp = f < 3 ? f : -1;
        k.zface = f == 2 ? 1 : 0;
        const int t1 = k.d == 0 ? 1 : 0, t2 = k.d == 2 ? 1 : 2;
        k.n1 = g_.N[t1]; k.n2 = g_.N[t2];
        Dim3 grid;
        grid.x = (k.n1 + FluxArrayKernel<FT>::THREADS - 1) / FluxArrayKernel<FT>::THREADS; grid.y = k.n2; grid.z = 1;
        go(k, grid, 0, OC_TIMER_SUBSTEP);
    }
}

template <class FT>
void Model<FT>::set_bc_array(int field, int side, const void* host, size_t nbytes) {
    if (field < 0 || field >= F_) throw Error(OC_ERR_INVALID, "set_bc_array: not a prognostic field index");
    if (side < 0 || side > 5) throw Error(OC_ERR_INVALID, "set_bc_array: side must be 0 … 5 (west, east, south, north, bottom, top)");
    const oc_bc& ub = cfg_.bcs[field][side];
    {
        // distributed models: the array is this rank's share of the side (N₁ × N₂ LOCAL values).  On a rank where the side is connected to
        // a neighbour instead of a wall the condition does not apply — but the call is collective (the refreshed halo plane travels
        // to the neighbours), so such a rank still takes part in the exchange
        const int d = side / 2;
        if (dist_ && g_.bounded[d] && !(side % 2 == 0 ? g_.wlo[d] : g_.whi[d])) {
            if (ub.kind == OC_BC_VALUE || ub.kind == OC_BC_GRADIENT) { join_tracers(); std::vector<FieldRec*> one{&state_[field]}; halo(one, false); }
            return;
        }
    }
    const int kind = state_[field].bc[side].kind;
    if (!(kind == ub.kind && (kind == OC_BC_FLUX || kind == OC_BC_VALUE || kind == OC_BC_GRADIENT)))
        throw Error(OC_ERR_INVALID, "set_bc_array: this side of the field must have been created with a Flux, Value or Gradient boundary condition");
    if (kind != OC_BC_FLUX && state_[field].face[side / 2])
        throw Error(OC_ERR_INVALID, "set_bc_array: Value / Gradient conditions apply to fields located at Center in the wall-normal direction");
    const int d = side / 2, t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
    const size_t n = (size_t)g_.N[t1] * g_.N[t2];
    if (n * sizeof(FT) != nbytes) throw Error(OC_ERR_INVALID, "host buffer size mismatch: expected " + std::to_string(n * sizeof(FT)) + " bytes");
    join_tracers();
    if (!bc_array_[field][side]) {
        bc_array_[field][side] = (FT*)dev_alloc(n * sizeof(FT));
        device_bytes += (int64_t)(n * sizeof(FT));
    }
    dev_upload(bc_array_[field][side], host, n * sizeof(FT), stream_);
#ifndef OC_HOSTSIM
    cuda_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");     // the caller's buffer may go away
#endif
    cfg_.bcs[field][side].has_value = 1;
    graphs_clear();                                       // (a first upload allocates the array: new kernel arguments)
    tend_valid_ = false;
    aux_valid_ = false;
    if (kind != OC_BC_FLUX) { std::vector<FieldRec*> one{&state_[field]}; halo(one, false); }      // the halo plane follows the new values
}

// boundary conditions of νₑ / κₑ (build_diffusivity_fields: anisotropic_minimum_dissipation.jl:358-372, smagorinsky.jl:163-176)
template <class FT>
void Model<FT>::set_diffusivity_bc(int field, int side, int kind, double value) {
    const bool is_nu = field == OC_FIELD_NU_E, is_kappa = field >= OC_FIELD_KAPPA_E0 && field < OC_FIELD_KAPPA_E0 + (int)kappa_e_.size();
    if (!has_eddy_ || !(is_nu || is_kappa)) throw Error(OC_ERR_INVALID, "set_diffusivity_bc: OC_FIELD_NU_E or OC_FIELD_KAPPA_E0 + tracer of a model with an eddy-viscosity closure");
    if (side < 0 || side > 5) throw Error(OC_ERR_INVALID, "set_diffusivity_bc: side must be 0 … 5 (west, east, south, north, bottom, top)");
    if (cfg_.topology[side / 2] != OC_BOUNDED) throw Error(OC_ERR_INVALID, "set_diffusivity_bc: non-periodic boundary condition in a dimension that is not Bounded");
    if (kind != OC_BC_FLUX && kind != OC_BC_VALUE && kind != OC_BC_GRADIENT) throw Error(OC_ERR_INVALID, "set_diffusivity_bc: Flux (no flux), Value or Gradient");
    join_tracers();
    FieldRec& f = lookup(field);
    f.bc[side].kind = kind;
    f.bc[side].value = kind == OC_BC_FLUX ? 0.0 : value;
    graphs_clear();                                       // kernel arguments are baked into captured graphs
    aux_valid_ = false;
    tend_valid_ = false;
}

template <class FT>
void Model<FT>::rk3_substep(double dt, int stage) {
    if (stage < 1 || stage > 3) throw Error(OC_ERR_INVALID, "RK3 stage must be 1, 2 or 3");
    for (int f = 0; f < F_; ++f) {
        SubstepKernel<FT> k;
        k.g = g_;
        k.U = state_[f].p; k.Gn = Gn_[f].p; k.Gm = Gm_[f].p;
        k.comp = f < 3 ? f : -1;
        k.dt = (FT)dt;
        k.ab2_euler = 0;
        if (stage == 1) { k.mode = STEP_RK3_FIRST; k.ca = (FT)dt * gamma_[0]; k.cb = FT(0); }
        else { k.mode = STEP_RK3; k.ca = gamma_[stage - 1]; k.cb = zeta_[stage - 1]; }
        go(k, grid_xyz(256), 0, OC_TIMER_SUBSTEP);
    }
    tend_valid_ = false;
    aux_valid_ = false;
}

template <class FT>
void Model<FT>::ab2_step(double dt, double chi) {
    for (int f = 0; f < F_; ++f) {
        SubstepKernel<FT> k;
        k.g = g_;
        k.U = state_[f].p; k.Gn = Gn_[f].p; k.Gm = Gm_[f].p;
        k.comp = f < 3 ? f : -1;
        k.mode = STEP_AB2;
        k.dt = (FT)dt;
        k.ca = FT(1.5) + (FT)chi; k.cb = FT(0.5) + (FT)chi;
        k.ab2_euler = ((FT)chi == FT(-0.5)) ? 1 : 0;
        go(k, grid_xyz(256), 0, OC_TIMER_SUBSTEP);
    }
    tend_valid_ = false;
    aux_valid_ = false;
}

template <class FT>
void Model<FT>::cache_previous_tendencies() {
    for (int f = 0; f < F_; ++f) {
        CopyKernel<FT> k;
        k.g = g_;
        k.dst = Gm_[f].p; k.src = Gn_[f].p;
        go(k, grid_xyz(256), 0, OC_TIMER_SUBSTEP);
    }
}

// ---------------------------------------------------------------------------------------------------------
// pressure
// ---------------------------------------------------------------------------------------------------------
template <class FT>
void Model<FT>::run_fft_solve() {
    NvtxRange nvtx_("solve! (FFTBasedPoissonSolver)");
    if (dist_) { run_fft_solve_dist(); return; }
    begin_timer(OC_TIMER_FFT);
    std::string e = replay_ ? std::string() : fft_.forward(fftbuf_);
    end_timer();
    if (!e.empty()) throw Error(OC_ERR_CUDA, e);
    if (stretched_) {
        // solve!(ϕ, ::BatchedTridiagonalSolver, rhs) + ϕ .-= mean(ϕ)   fourier_tridiagonal_poisson_solver.jl:213-226
        if (!g_.bounded[0] && !g_.bounded[1]) {
            TridiagSolvePPKernel<FT> k;
            k.L = fft_.L;
            k.spec = fftbuf_;
            k.R = tri_R_; k.T = tri_T_; k.rdzf = g_.rdzf;
            k.norm = 1.0 / ((double)g_.N[0] * g_.N[1]);
            k.zero_col = 1;
            Dim3 grid;
            const long long n2 = 2LL * fft_.L.nxc * g_.N[1];
            grid.x = (int)((n2 + TridiagSolvePPKernel<FT>::THREADS - 1) / TridiagSolvePPKernel<FT>::THREADS);
            go(k, grid, 0, OC_TIMER_POISSON_MID);
        } else {
        TridiagSolveKernel<FT> k;
        k.L = fft_.L;
        k.spec = reinterpret_cast<Cplx<FT>*>(fftbuf_);
        k.R = tri_R_; k.T = tri_T_; k.rdzf = g_.rdzf;
        k.tw[0] = tw_[0]; k.tw[1] = tw_[1];
        k.nrep[0] = g_.bounded[0] ? g_.N[0] / 2 + 1 : fft_.L.nxc;
        k.nrep[1] = g_.bounded[1] ? g_.N[1] / 2 + 1 : g_.N[1];
        k.norm = 1.0 / ((double)g_.N[0] * g_.N[1]);
        Dim3 grid;
        grid.x = (k.nrep[0] + TridiagSolveKernel<FT>::THREADS - 1) / TridiagSolveKernel<FT>::THREADS;
        grid.y = k.nrep[1];
        go(k, grid, 0, OC_TIMER_POISSON_MID);
        }
    } else if (!g_.bounded[0] && !g_.bounded[1] && !g_.bounded[2]) {
        PoissonDivideKernel<FT> k;
        k.L = fft_.L;
        k.spec = reinterpret_cast<Cplx<FT>*>(fftbuf_);
        for (int d = 0; d < 3; ++d) k.lam[d] = lam_[d];
        k.norm = 1.0 / ((double)g_.N[0] * g_.N[1] * g_.N[2]);
        k.chunk = 2048;
        Dim3 grid;
        grid.x = (fft_.L.nxc * g_.N[1] + k.chunk - 1) / k.chunk;
        grid.y = g_.N[2];
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    } else if (!g_.bounded[0] && !g_.bounded[1] && g_.bounded[2]) {
        PoissonMidZKernel<FT> k;
        k.L = fft_.L;
        k.spec = reinterpret_cast<Cplx<FT>*>(fftbuf_);
        for (int d = 0; d < 3; ++d) k.lam[d] = lam_[d];
        k.twz = tw_[2];
        k.norm = 1.0 / ((double)g_.N[0] * g_.N[1] * g_.N[2]);
        Dim3 grid;
        grid.x = (fft_.L.nxc * g_.N[1] + 255) / 256;
        grid.y = g_.N[2] / 2 + 1;
        go(k, grid, 0, OC_TIMER_POISSON_MID);
    } else {
    PoissonMidKernel<FT> k;
    k.L = fft_.L;
    k.spec = reinterpret_cast<Cplx<FT>*>(fftbuf_);
    for (int d = 0; d < 3; ++d) { k.lam[d] = lam_[d]; k.tw[d] = tw_[d]; }
    k.nrep[0] = g_.bounded[0] ? g_.N[0] / 2 + 1 : fft_.L.nxc;
    k.nrep[1] = g_.bounded[1] ? g_.N[1] / 2 + 1 : g_.N[1];
    k.nrep[2] = g_.bounded[2] ? g_.N[2] / 2 + 1 : g_.N[2];
    k.norm = 1.0 / ((double)g_.N[0] * g_.N[1] * g_.N[2]);
    Dim3 grid;
    grid.x = (k.nrep[0] + PoissonMidKernel<FT>::THREADS - 1) / PoissonMidKernel<FT>::THREADS;
    grid.y = k.nrep[1];
    grid.z = k.nrep[2];
    go(k, grid, 0, OC_TIMER_POISSON_MID);
    }
    begin_timer(OC_TIMER_FFT);
    e = replay_ ? std::string() : fft_.inverse(fftbuf_);
    end_timer();
    if (!e.empty()) throw Error(OC_ERR_CUDA, e);
}

template <class FT>
void Model<FT>::pressure_solve_from_state() {
    NvtxRange nvtx_("compute_pressure_correction!");
    PoissonRhsKernel<FT> k;
    k.g = g_;
    k.L = fft_.L;
    k.u = state_[0].p; k.v = state_[1].p; k.w = state_[2].p;
    k.buf = fftbuf_;
    go(k, grid_xyz(256), 0, OC_TIMER_POISSON_RHS);
    run_fft_solve();
}

template <class FT>
void Model<FT>::projection(double dt) {
    NvtxRange nvtx_("make_pressure_correction!");
    const FT* prev_row = nullptr;
    const FT* prev_col = nullptr;
    if (dist_) {
        // the pressure gradient at the first local row needs the y-neighbour's last row of ϕ: one dense (Nx, Nz) message
        // (the reference fills all of pNHS's halos, halo_communication.jl:87-187; only this row is ever read)
        const size_t n = (size_t)g_.N[0] * g_.N[2];
        if (2 * n > halo_buf_elems_) throw Error(OC_ERR_STATE, "internal: exchange buffer too small for the ϕ row");
        PhiRowKernel<FT> r;
        r.L = fft_.L; r.buf = fftbuf_; r.row = halo_send_;
        Dim3 rg;
        rg.x = (g_.N[0] + 255) / 256; rg.y = g_.N[2];
        go(r, rg, 0, OC_TIMER_COMM);
        std::vector<Msg> msgs;
        if (R_ > 1) {
            const int prev = grank(rx_, (rank_ + R_ - 1) % R_), next = grank(rx_, (rank_ + 1) % R_);
            msgs.push_back(Msg{next, prev, 1, halo_send_, g_.whi[1] ? 0 : n * sizeof(FT), halo_recv_, g_.wlo[1] ? 0 : n * sizeof(FT)});
        }
        const size_t m = (size_t)g_.N[1] * g_.N[2];
        if (Rx_ > 1) {
            // pencils: likewise the x-neighbour's last column, dense (Ny, Nz)
            if (2 * (n + m) > halo_buf_elems_) throw Error(OC_ERR_STATE, "internal: exchange buffer too small for the ϕ column");
            PhiColKernel<FT> c;
            c.L = fft_.L; c.buf = fftbuf_; c.col = halo_send_ + n;
            Dim3 cg;
            cg.x = (g_.N[1] + 255) / 256; cg.y = g_.N[2];
            go(c, cg, 0, OC_TIMER_COMM);
            const int prev = grank((rx_ + Rx_ - 1) % Rx_, rank_), next = grank((rx_ + 1) % Rx_, rank_);
            msgs.push_back(Msg{next, prev, 2, halo_send_ + n, g_.whi[0] ? 0 : m * sizeof(FT), halo_recv_ + n, g_.wlo[0] ? 0 : m * sizeof(FT)});
        }
        begin_timer(OC_TIMER_COMM);
        std::string e = transport_ ? transport_->exchange(msgs, stream_) : std::string("distributed model without a transport");
        end_timer();
        if (!e.empty()) throw Error(OC_ERR_CUDA, e);
        prev_row = (g_.wlo[1] || R_ <= 1) ? nullptr : halo_recv_;          // first slab of a Bounded y: the wall face is not corrected
        prev_col = (g_.wlo[0] || Rx_ <= 1) ? nullptr : halo_recv_ + n;
    }
    ProjectionKernel<FT> k;
    k.g = g_;
    k.L = fft_.L;
    k.buf = fftbuf_;
    k.u = state_[0].p; k.v = state_[1].p; k.w = state_[2].p;
    k.pNHS = pNHS_.p;
    k.prev_row = prev_row;
    k.prev_col = prev_col;
    k.dt_plus = std::max((double)std::numeric_limits<FT>::epsilon(), dt);
    go(k, grid_xyz(256), 0, OC_TIMER_PROJECTION);
}

template <class FT>
void Model<FT>::compute_pressure_correction(double dt) {
    (void)dt;
    std::vector<FieldRec*> vel{&state_[0], &state_[1], &state_[2]};
    halo(vel, true);
    pressure_solve_from_state();
    PoissonUnpackKernel<FT> k;
    k.g = g_;
    k.L = fft_.L;
    k.buf = fftbuf_;
    k.field = pNHS_.p;
    k.dense = nullptr;
    go(k, grid_xyz(256), 0, OC_TIMER_PROJECTION);
    std::vector<FieldRec*> p{&pNHS_};
    halo(p, true);
}

template <class FT>
void Model<FT>::make_pressure_correction(double dt) {
    GradSubKernel<FT> k;
    k.g = g_;
    k.u = state_[0].p; k.v = state_[1].p; k.w = state_[2].p;
    k.p = pNHS_.p;
    k.dt_plus = 1.0;
    go(k, grid_xyz(256), 0, OC_TIMER_PROJECTION);
    ScaleKernel<FT> s;
    s.g = g_;
    s.p = pNHS_.p;
    s.dt_plus = std::max((double)std::numeric_limits<FT>::epsilon(), dt);
    go(s, grid_xyz(256), 0, OC_TIMER_PROJECTION);
    tend_valid_ = false;
    aux_valid_ = false;
}

template <class FT>
void Model<FT>::diagnostics(oc_diagnostics* out) {
    join_tracers();
    if (!diag_dev_) diag_dev_ = (unsigned long long*)dev_alloc(sizeof(unsigned long long) * 5);
    const double big = 1.0e300;
    unsigned long long init[5] = {0, 0, 0, 0, 0};
    memcpy(&init[0], &big, 8);
    dev_upload(diag_dev_, init, sizeof(init), stream_);
    DiagnosticsKernel<FT> k;
    k.g = g_;
    k.u = state_[0].p; k.v = state_[1].p; k.w = state_[2].p;
    k.out = diag_dev_;
    Dim3 grid;
    grid.x = 1; grid.y = g_.N[1]; grid.z = g_.N[2];
    go(k, grid, DiagnosticsKernel<FT>::SMEM, OC_TIMER_AUX);
    unsigned long long res[5];
    dev_download(res, diag_dev_, sizeof(res), stream_);
    double d[4];
    memcpy(d, res, sizeof(d));
    out->cell_advection_timescale = d[0] >= big ? INFINITY : d[0];
    out->max_abs_u = d[1]; out->max_abs_v = d[2]; out->max_abs_w = d[3];
    out->has_nan = res[4] ? 1 : 0;
    out->pad = 0;
}

// ---------------------------------------------------------------------------------------------------------
// asynchronous output: snapshot a box of a field in stream order (D2D into a staging buffer), then copy it to the host on a separate
// stream while the time stepping continues
// ---------------------------------------------------------------------------------------------------------
// a free ticket slot with its events and a device staging buffer of at least nbytes
template <class FT>
int Model<FT>::acquire_slot(size_t nbytes) {
    int t = -1;
    // prefer a free slot whose staging buffer is already large enough (steady-state loops then never reallocate)
    for (size_t i = 0; i < out_slots_.size(); ++i) if (!out_slots_[i].busy && out_slots_[i].cap >= nbytes) { t = (int)i; break; }
    if (t < 0) for (size_t i = 0; i < out_slots_.size(); ++i) if (!out_slots_[i].busy) { t = (int)i; break; }
    if (t < 0) {
        if (out_slots_.size() >= 64) throw Error(OC_ERR_STATE, "more than 64 transfers in flight: call oc_output_wait");
        out_slots_.emplace_back();
        t = (int)out_slots_.size() - 1;
    }
    OutputSlot& s = out_slots_[t];
#ifndef OC_HOSTSIM
    if (!s.ev_snap) {
        cudaEvent_t a, b;
        cuda_check(cudaEventCreateWithFlags(&a, cudaEventDisableTiming), "cudaEventCreate");
        cuda_check(cudaEventCreateWithFlags(&b, cudaEventDisableTiming), "cudaEventCreate");
        s.ev_snap = a; s.ev_done = b;
    }
    if (s.cap < nbytes) {
        if (s.stage) { cuda_check(cudaFree(s.stage), "cudaFree"); device_bytes -= (int64_t)s.cap; }
        s.stage = nullptr; s.cap = 0;
        void* p = nullptr;
        cuda_check(cudaMalloc(&p, nbytes), "cudaMalloc(transfer staging)");
        s.stage = (FT*)p; s.cap = nbytes;
        device_bytes += (int64_t)nbytes;
    }
#else
    (void)nbytes;
#endif
    return t;
}

// set!(field, host_array) in stream order without a host-side wait: H2D into staging on in_stream_, D2D into the field on stream_
template <class FT>
int Model<FT>::upload_begin(int field, const void* host, size_t nbytes) {
    NvtxRange nvtx_("upload_begin");
    if (field < 0 || field >= F_) throw Error(OC_ERR_INVALID, "oc_upload_begin: not a prognostic field index");
    join_tracers();
    oc_field_info info;
    field_info(field, &info);
    FieldRec& f = lookup(field);
    size_t cnt = 1;
    int n[3];
    for (int d = 0; d < 3; ++d) { n[d] = info.interior_size[d]; cnt *= (size_t)n[d]; }
    if (cnt * sizeof(FT) != nbytes) throw Error(OC_ERR_INVALID, "host buffer size mismatch: expected " + std::to_string(cnt * sizeof(FT)) + " bytes");
    const int t = acquire_slot(nbytes);
    OutputSlot& s = out_slots_[t];
#ifndef OC_HOSTSIM
    if (!in_stream_) cuda_check(cudaStreamCreateWithFlags(&in_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    cuda_check(cudaMemcpyAsync(s.stage, host, nbytes, cudaMemcpyHostToDevice, in_stream_), "cudaMemcpyAsync(upload)");
    cuda_check(cudaEventRecord((cudaEvent_t)s.ev_snap, in_stream_), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(stream_, (cudaEvent_t)s.ev_snap, 0), "cudaStreamWaitEvent");
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr(s.stage, (size_t)n[0] * sizeof(FT), (size_t)n[0] * sizeof(FT), (size_t)n[1]);
    p.dstPtr = make_cudaPitchedPtr(f.p, (size_t)g_.sy * sizeof(FT), (size_t)g_.sy * sizeof(FT), (size_t)(g_.sz / g_.sy));
    p.extent = make_cudaExtent((size_t)n[0] * sizeof(FT), (size_t)n[1], (size_t)n[2]);
    p.kind = cudaMemcpyDeviceToDevice;
    cuda_check(cudaMemcpy3DAsync(&p, stream_), "cudaMemcpy3DAsync(upload)");
    cuda_check(cudaEventRecord((cudaEvent_t)s.ev_done, stream_), "cudaEventRecord");
#else
    dev_copy_box(f.p, sizeof(FT), g_.sy, g_.sz, const_cast<void*>(host), n, true, stream_);
#endif
    if (g_.flat[0] || g_.flat[1] || g_.flat[2]) { std::vector<FieldRec*> one{&f}; halo(one, false); }     // like transfer(): Flat dimensions are stored as periodic N = 1
    tend_valid_ = false;
    aux_valid_ = false;
    s.busy = true;
    return t;
}

template <class FT>
int Model<FT>::output_begin(int field, const int lo[3], const int n[3], void* host, size_t nbytes) {
    NvtxRange nvtx_("output_begin");
    join_tracers();
    oc_field_info info;
    field_info(field, &info);                // brings auxiliary fields / tendencies up to date like a download
    FieldRec& f = lookup(field);
    size_t cnt = 1;
    for (int d = 0; d < 3; ++d) {
        if (n[d] < 1 || lo[d] < -Hcfg_[d] || lo[d] + n[d] > info.interior_size[d] + Hcfg_[d])
            throw Error(OC_ERR_INVALID, "output box outside the parent array of the field");
        cnt *= (size_t)n[d];
    }
    if (cnt * sizeof(FT) != nbytes) throw Error(OC_ERR_INVALID, "host buffer size mismatch: expected " + std::to_string(cnt * sizeof(FT)) + " bytes");
    const int t = acquire_slot(nbytes);
    OutputSlot& s = out_slots_[t];
    FT* origin = f.p + lo[0] + (long long)lo[1] * g_.sy + (long long)lo[2] * g_.sz;
#ifndef OC_HOSTSIM
    if (!out_stream_) cuda_check(cudaStreamCreateWithFlags(&out_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr(origin, (size_t)g_.sy * sizeof(FT), (size_t)g_.sy * sizeof(FT), (size_t)(g_.sz / g_.sy));
    p.dstPtr = make_cudaPitchedPtr(s.stage, (size_t)n[0] * sizeof(FT), (size_t)n[0] * sizeof(FT), (size_t)n[1]);
    p.extent = make_cudaExtent((size_t)n[0] * sizeof(FT), (size_t)n[1], (size_t)n[2]);
    p.kind = cudaMemcpyDeviceToDevice;
    cuda_check(cudaMemcpy3DAsync(&p, stream_), "cudaMemcpy3DAsync(snapshot)");
    cuda_check(cudaEventRecord((cudaEvent_t)s.ev_snap, stream_), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(out_stream_, (cudaEvent_t)s.ev_snap, 0), "cudaStreamWaitEvent");
    cuda_check(cudaMemcpyAsync(host, s.stage, nbytes, cudaMemcpyDeviceToHost, out_stream_), "cudaMemcpyAsync(output)");
    cuda_check(cudaEventRecord((cudaEvent_t)s.ev_done, out_stream_), "cudaEventRecord");
#else
    dev_copy_box(origin, sizeof(FT), g_.sy, g_.sz, host, n, false, stream_);      // the host simulation snapshots synchronously
#endif
    s.busy = true;
    return t;
}

template <class FT>
void Model<FT>::output_wait(int ticket) {
    if (ticket < 0 || ticket >= (int)out_slots_.size() || !out_slots_[ticket].busy) throw Error(OC_ERR_INVALID, "oc_output_wait: unknown or finished ticket");
#ifndef OC_HOSTSIM
    cuda_check(cudaEventSynchronize((cudaEvent_t)out_slots_[ticket].ev_done), "cudaEventSynchronize(output)");
#endif
    out_slots_[ticket].busy = false;
}

template <class FT>
bool Model<FT>::output_test(int ticket) {
    if (ticket < 0 || ticket >= (int)out_slots_.size() || !out_slots_[ticket].busy) throw Error(OC_ERR_INVALID, "oc_output_test: unknown or finished ticket");
#ifndef OC_HOSTSIM
    cudaError_t e = cudaEventQuery((cudaEvent_t)out_slots_[ticket].ev_done);
    if (e == cudaErrorNotReady) return false;
    cuda_check(e, "cudaEventQuery(output)");
#endif
    return true;
}

// maximum(abs, interior(field)); NaN if the field holds a NaN (like Julia's maximum)
template <class FT>
double Model<FT>::field_maximum_abs(int field) {
    join_tracers();
    oc_field_info info;
    field_info(field, &info);          // brings auxiliary fields / tendencies up to date exactly like a download would
    FieldRec& f = lookup(field);
    if (!diag_dev_) diag_dev_ = (unsigned long long*)dev_alloc(sizeof(unsigned long long) * 5);
    unsigned long long init[2] = {0, 0};
    dev_upload(diag_dev_, init, sizeof(init), stream_);
    FieldMaxAbsKernel<FT> k;
    k.g = g_;
    k.f = f.p;
    k.nx = info.interior_size[0];
    k.out = diag_dev_;
    Dim3 grid;
    grid.x = 1; grid.y = info.interior_size[1]; grid.z = info.interior_size[2];
    go(k, grid, FieldMaxAbsKernel<FT>::SMEM, OC_TIMER_AUX);
    unsigned long long res[2];
    dev_download(res, diag_dev_, sizeof(res), stream_);
    double d;
    memcpy(&d, &res[0], 8);
    return res[1] ? (double)NAN : d;
}

// Checkpointer pickup of G⁻ (checkpointer.jl:230-262).  Between steps the library keeps the tendencies of the last substep in the
// Gⁿ slot and marks them stale; the next stage turns them into G⁻ by a pointer swap (stage()).  A restored G⁻ therefore goes into
// that slot, and the stale mark is (re)asserted so that no fresh evaluation is swapped over it.
template <class FT>
void Model<FT>::restore_previous_tendency(int field, const void* host, size_t nbytes) {
    if (field < 0 || field >= F_) throw Error(OC_ERR_INVALID, "restore_previous_tendency: not a prognostic field index");
    join_tracers();
    FieldRec& f = Gn_[field];
    int ext[3];
    size_t n = 1;
    for (int d = 0; d < 3; ++d) {
        ext[d] = g_.N[d] + ((f.face[d] && g_.whi[d]) ? 1 : 0) + 2 * Hcfg_[d];
        n *= (size_t)ext[d];
    }
    if (n * sizeof(FT) != nbytes) throw Error(OC_ERR_INVALID, "host buffer size mismatch: expected " + std::to_string(n * sizeof(FT)) + " bytes");
    FT* origin = f.p - Hcfg_[0] - (long long)Hcfg_[1] * g_.sy - (long long)Hcfg_[2] * g_.sz;
    dev_copy_box(origin, sizeof(FT), g_.sy, g_.sz, const_cast<void*>(host), ext, true, stream_);
    gn_pending_ = true;
    tend_valid_ = false;
}

template <class FT>
void Model<FT>::poisson_solve(const void* rhs, void* phi, size_t nbytes) {
    size_t n = (size_t)g_.N[0] * g_.N[1] * g_.N[2];
    if (nbytes != n * sizeof(FT)) throw Error(OC_ERR_INVALID, "poisson_solve: buffer size mismatch");
    FT* dense = (FT*)dev_alloc(nbytes);
    dev_upload(dense, rhs, nbytes, stream_);
    PoissonLoadKernel<FT> l;
    l.L = fft_.L;
    l.rhs = dense;
    l.buf = fftbuf_;
    l.dzc = stretched_ ? g_.dzc : nullptr;
    go(l, grid_xyz(256), 0, OC_TIMER_POISSON_RHS);
    run_fft_solve();
    PoissonUnpackKernel<FT> k;
    k.g = g_;
    k.L = fft_.L;
    k.buf = fftbuf_;
    k.field = nullptr;
    k.dense = dense;
    go(k, grid_xyz(256), 0, OC_TIMER_PROJECTION);
    dev_download(phi, dense, nbytes, stream_);
    dev_free(dense);
}

// ---------------------------------------------------------------------------------------------------------
// CUDA Graphs for launch-bound grids
// ---------------------------------------------------------------------------------------------------------
template <class FT>
bool Model<FT>::graph_eligible() const {
#ifndef OC_HOSTSIM
    static const char* env = getenv("OC_GRAPHS");
    if (env && atoi(env) == 0) return false;
    const long long cells = (long long)g_.N[0] * g_.N[1] * g_.N[2];
    return !timing_ && !dist_ && !tracers_in_flight_ && cells <= (env && atoi(env) > 1 ? (1LL << 40) : (1LL << 20));     // OC_GRAPHS=2: any size (measurement)
#else
    return false;
#endif
}

template <class FT>
void Model<FT>::graphs_clear() {
#ifndef OC_HOSTSIM
    for (auto& kv : graphs_) if (kv.second.exec) cudaGraphExecDestroy((cudaGraphExec_t)kv.second.exec);
#endif
    graphs_.clear();
}

// body() enqueues one time step AND does its host-side bookkeeping.  First use of a key: eager (lazy initialisations — halo box tables,
// function attributes — happen outside any capture).  Second use: captured into a graph, instantiated, launched.  From then on: body()
// runs with replay_ set (bookkeeping only) and the graph is launched.
template <class FT>
template <class Body>
void Model<FT>::run_graphed(const std::string& key, Body&& body) {
#ifndef OC_HOSTSIM
    if (!graph_eligible()) { body(); return; }
    if (graphs_.size() > 16) graphs_clear();
    GraphEntry& e = graphs_[key];
    if (e.exec) {
        replay_ = true;
        try { body(); } catch (...) { replay_ = false; throw; }
        replay_ = false;
        cuda_check(cudaGraphLaunch((cudaGraphExec_t)e.exec, stream_), "cudaGraphLaunch");
        return;
    }
    if (e.seen++ == 0) { body(); return; }
    cuda_check(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
    cudaGraph_t graph = nullptr;
    try {
        body();
    } catch (...) {
        cudaStreamEndCapture(stream_, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        throw;
    }
    cuda_check(cudaStreamEndCapture(stream_, &graph), "cudaStreamEndCapture");
    cudaGraphExec_t exec = nullptr;
    cudaError_t err = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    cuda_check(err, "cudaGraphInstantiate");
    e.exec = exec;
    cuda_check(cudaGraphLaunch(exec, stream_), "cudaGraphLaunch");
#else
    (void)key;
    body();
#endif
}

// ---------------------------------------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------------------------------------
template <class FT>
void Model<FT>::set_finalize(int enforce) {
    NvtxRange nvtx_("set!(model)");
    std::vector<FieldRec*> all;
    for (auto& f : state_) all.push_back(&f);
    halo(all, true);                                     // set!: fill_halo_regions!(ϕ) per field (open BCs included)
    aux_valid_ = false;
    tend_valid_ = false;
    if (enforce) {
        pressure_solve_from_state();                     // Δt = 1   set_nonhydrostatic_model.jl:52-56
        projection(1.0);
        halo(all, false);
    }
}

// One fused stage: [aux] -> tendency+substep per field -> halo(U, open) -> Poisson -> projection -> halo(all)
template <class FT>
void Model<FT>::stage(int mode, double dt, int stage_no, double stage_dt, double chi, bool euler) {
    NvtxRange nvtx_("time-stepper stage");
    rotate_pending_tendencies();                         // the previous stage's evaluation becomes G⁻ (cache by swap)
    tendencies(mode, dt, stage_no, chi, euler, true, true, /*defer_tracer_join=*/true);
    gn_pending_ = true;
    tend_valid_ = false;
    aux_valid_ = false;
    std::vector<FieldRec*> vel{&state_[0], &state_[1], &state_[2]};
    halo(vel, true);
    pressure_solve_from_state();
    projection(stage_dt);
    join_tracers();                                      // the tracer substeps (stream2) before their halos are filled
    std::vector<FieldRec*> all;
    for (auto& f : state_) all.push_back(&f);
    // stages 1 and 2 of an RK3 step are followed by another stage of the same call: their y-halo exchange is overlapped with that stage's
    // interior tendency kernels (tendencies()); the last stage's exchange completes before time_step! returns.  OC_XCHG_OVERLAP=0: off.
    const char* xo_env = getenv("OC_XCHG_OVERLAP");
    const bool defer = dist_ && Rx_ == 1 && R_ > 1 && march_ok_ && !has_eddy_ && mode != STEP_AB2 && stage_no < 3 && (xo_env ? atoi(xo_env) != 0 : true);
    halo(all, false, defer);
}

template <class FT>
void Model<FT>::time_step_rk3(double dt) {
    NvtxRange nvtx_("time_step! (RungeKutta3)");
    if (cfg_.timestepper != OC_RK3) throw Error(OC_ERR_STATE, "model was created with another time stepper");
    // stage_Δt(Δt, γ, ζ) = Δt (γ + ζ) with γ, ζ in FT   runge_kutta_3.jl:107-109,176-177
    const double dt1 = dt * (double)gamma_[0];
    const double dt2 = dt * (double)(FT)(gamma_[1] + zeta_[1]);
    const double dt3 = dt * (double)(FT)(gamma_[2] + zeta_[2]);
    const double tn1 = clock.time + dt;
    join_tracers();
    char key[160];
    snprintf(key, sizeof(key), "rk3 %p %p %d %d %a", (void*)state_[0].p, (void*)Gn_[0].p, (int)gn_pending_, (int)aux_valid_, dt);
    run_graphed(key, [&] {
        stage(STEP_RK3_FIRST, dt, 1, dt1, 0.0, false);
        stage(STEP_RK3, dt, 2, dt2, 0.0, false);
        stage(STEP_RK3, dt, 3, dt3, 0.0, false);
    });
    clock.time += dt1; clock.stage += 1; clock.last_stage_dt = dt1;
    clock.time += dt2; clock.stage += 1; clock.last_stage_dt = dt2;
    const double corrected = tn1 - clock.time;           // :148-161
    clock.time += dt3;
    clock.iteration += 1;
    clock.stage = 1;
    clock.last_dt = dt;
    clock.last_stage_dt = corrected;
}

template <class FT>
void Model<FT>::time_step_ab2(double dt, int euler_in) {
    NvtxRange nvtx_("time_step! (QuasiAdamsBashforth2)");
    if (cfg_.timestepper != OC_AB2) throw Error(OC_ERR_STATE, "model was created with another time stepper");
    const bool euler = euler_in || (dt != clock.last_dt);            // quasi_adams_bashforth_2.jl:88
    const double chi = euler ? -0.5 : cfg_.ab2_chi;
    join_tracers();
    char key[160];
    snprintf(key, sizeof(key), "ab2 %p %p %d %d %a %d", (void*)state_[0].p, (void*)Gn_[0].p, (int)gn_pending_, (int)aux_valid_, dt, (int)euler);
    run_graphed(key, [&] { stage(STEP_AB2, dt, 1, dt, chi, euler); });
    clock.time += dt;
    clock.iteration += 1;
    clock.stage = 1;
    clock.last_dt = dt;
    clock.last_stage_dt = dt;
}

}  // namespace oc
