// oc_uvw.h — ONE launch for the three momentum tendencies (+ RK3 / AB2 substep) of Centered(order = 2) models on grids without a
// Bounded dimension: the fused schedule of SURVEY §8d (K1) for the bandwidth-bound scheme.
//
// Replaces compute_Gu! + compute_Gv! + compute_Gw! (compute_nonhydrostatic_tendencies.jl:83-111, nonhydrostatic_tendency_kernel_functions.jl
// :70-230 with centered_advective_fluxes.jl:15-27 and abstract_scalar_diffusivity_closure.jl:189-204) and their substeps.  Per-field
// launches (oc_march.h) read u, v and w three times per stage — 21 field passes where 12 are needed; with ~35 FP64 instructions per
// cell and field Centered(2) is bound by exactly that traffic (BASELINE config C2: three launches of 0.21 ms moving 0.94 GB each).
//
// Design.  A CTA owns a 32 × 8 column of cells and marches in z.  u, v and w are staged once, by TMA, into three rings of x–y planes
// with a one-cell halo (levels k-1 … k+1 live, PF more in flight).  Every thread then forms all 18 face fluxes of its cell's three
// control volumes directly from shared memory — each flux is evaluated by the two cells that share the face (Centered(2) fluxes cost
// ~10 FP64 instructions; recomputing them is cheaper than the shared-memory exchange and the per-level barrier it needs) with exactly
// the expressions of MarchKernel / the reference, so both cells see the same bits and the flux form stays conservative.  No flux
// buffers, no barrier on the critical path: the only synchronisation is the split arrive / wait that protects ring slots from the
// next TMA load (a warp waits only if another warp is a whole level behind).
#pragma once
#include "oc_march.h"

namespace oc {

template <class FT>
struct UvwArgs {
    Geom<FT> g;
    const FT* pHY;       // hydrostatic pressure anomaly or nullptr
    const FT* Gm[3];     // G⁻ (read)
    FT* Gn[3];           // Gⁿ (written)
    FT* Unew[3];         // next state (mode != STEP_NONE)
    FT nu;               // ScalarDiffusivity ν (0 without a closure)
    int has_coriolis;    // 0 none, 1 FPlane, 2 BetaPlane
    FT f, cor_beta, cor_y0;
    int mode;            // SubstepMode
    FT dt, ca, cb;
    int ab2_euler;
};

struct UvwState {
    int sk;       // ring slot of level k (all three rings share one geometry)
    int o;        // global offset of this thread's cell at level k
    int nit;
    int own;      // the cell lies inside the domain
};

template <class FT>
struct UvwCenteredKernel {
    static constexpr int TX = 32, TY = 8;
    static constexpr int THREADS = TX * TY;
    static constexpr int MIN_BLOCKS = 4;
    static constexpr int PF = 2, NSYNC = 2;
    static constexpr int E = 16 / (int)sizeof(FT);                              // box rows start 16-byte aligned
    using RS = RingSpec<-E, TX + 2 * E, -1, TY + 2, -1, 1, 3 + PF>;             // i-1 … i+1 (padded), j-1 … j+1, k-1 … k+1
    using G = Ring<FT, RS>;
    using SP = UvwCenteredKernel;                                               // (interface of launch_march_tendency: SP::PF …)
    static constexpr size_t OFF_BAR = 0, OFF_R = 128;
    static constexpr size_t SMEM = OFF_R + 3 * (size_t)G::BYTES;
    static constexpr int LEVEL_BYTES = 3 * G::BOX_BYTES;
    static_assert(SMEM <= (233472 - MIN_BLOCKS * 1024) / MIN_BLOCKS, "shared memory must allow MIN_BLOCKS CTAs per SM");
    typedef UvwState State;

    UvwArgs<FT> a;
    TileSrc<FT> src[3];   // u, v, w
    int xpad, KC, by0;

    OC_HD int k_begin(const Block& b) const { return b.z * KC; }
    OC_HD int k_end(const Block& b) const { int e = (b.z + 1) * KC; return e < a.g.N[2] ? e : a.g.N[2]; }
    OC_HD int iterations(const Block& b) const { return k_end(b) - k_begin(b); }

    OC_HD uint64_t* bar_d(char* smem, int it) const { return reinterpret_cast<uint64_t*>(smem + OFF_BAR) + (it % MARCH_NBAR); }
    OC_HD uint64_t* bar_s(char* smem, int it) const { return reinterpret_cast<uint64_t*>(smem + OFF_BAR) + MARCH_NBAR + (it % NSYNC); }
    OC_HD G ring(char* smem, int c, int k, int sk) const { return G{reinterpret_cast<FT*>(smem + OFF_R + (size_t)c * G::BYTES), k, sk}; }

    OC_DEV void issue_level(char* smem, int i0, int j0, int lev, uint64_t* bar) const {
        for (int c = 0; c < 3; ++c) {
            const G r = ring(smem, c, 0, 0);
            tile_issue<FT>(r.slot(lev), &src[c], i0 + RS::XO + xpad, j0 + RS::YO + a.g.H[1], lev + a.g.H[2], bar, RS::BX, RS::BY);
        }
    }

    OC_DEV void begin0(const Block&, int tid, char* smem) const {
        if (tid == 0) {
            uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
            for (int n = 0; n < MARCH_NBAR; ++n) mbar_init(bar + n, 1);
            for (int n = 0; n < NSYNC; ++n) mbar_init(bar + MARCH_NBAR + n, THREADS / 32);
            mbar_fence_init();
        }
    }
    OC_DEV void begin1(const Block& b, int tid, char* smem, State& st) const {
        const int lane = tid & (TX - 1), row = tid / TX;
        const int i = b.x * TX + lane, j = (b.y + by0) * TY + row, kb = k_begin(b);
        st.own = i < a.g.N[0] && j < a.g.N[1];
        st.o = a.g.idx(i, j, kb);
        st.nit = iterations(b);
        st.sk = G::slot_of(kb);
        if (tid != 0) return;
        const int i0 = b.x * TX, j0 = (b.y + by0) * TY;
        // iteration 0 needs levels kb-1, kb, kb+1; iteration it needs level kb+it+1 on top
        mbar_expect(bar_d(smem, 0), 3 * LEVEL_BYTES);
        for (int l = -1; l <= 1; ++l) issue_level(smem, i0, j0, kb + l, bar_d(smem, 0));
        for (int it = 1; it < PF && it < st.nit; ++it) {
            mbar_expect(bar_d(smem, it), LEVEL_BYTES);
            issue_level(smem, i0, j0, kb + it + 1, bar_d(smem, it));
        }
    }

    OC_DEV void sync_wait(char* smem, int it) const { mbar_wait(bar_s(smem, it), (it / NSYNC) & 1); }
    OC_DEV void sync_arrive(char* smem, int it) const {
#if defined(__CUDA_ARCH__)
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(bar_s(smem, it));
#else
        mbar_arrive(bar_s(smem, it));
#endif
    }

    // velocity component c at tile-local (ii, jj), level l
    struct View {
        G u, v, w;
        OC_HD FT operator()(int c, int ii, int jj, int l) const { return c == 0 ? u(ii, jj, l) : (c == 1 ? v(ii, jj, l) : w(ii, jj, l)); }
    };

    // Total flux (advective + viscous) of momentum component COMP through the faces normal to D at flux index (ii, jj, l): the cell
    // centre for D == COMP, the D-face otherwise.  The expressions — and their order — are MarchKernel's for ADV_CENTERED2, CLO = 0
    // (oc_march.h: advective_flux, viscous_flux, total_flux), i.e. centered_advective_fluxes.jl:15-27 and
    // abstract_scalar_diffusivity_closure.jl:189-204 with velocity_tracer_gradients.jl:25-42.
    template <int COMP, int D>
    OC_HD FT flux(const View& V, int ii, int jj, int l) const {
        const Geom<FT>& g = a.g;
        const FT A = g.A[D];
        constexpr int dx = D == 0, dy = D == 1, dz = D == 2;
        if (D == COMP) {
            const FT p0 = V(COMP, ii, jj, l), p1 = V(COMP, ii + dx, jj + dy, l + dz);
            const FT ut = FT(0.5) * p0 + FT(0.5) * p1;
            const FT F = A * ut * ut;
            const FT sig = (p1 - p0) * g.rd[D];
            return F + (FT(-2) * a.nu * A) * sig;
        } else {
            constexpr int cx = COMP == 0, cy = COMP == 1, cz = COMP == 2;
            constexpr int lo = D < COMP ? D : COMP, hi = D < COMP ? COMP : D;
            const FT ut = FT(0.5) * V(D, ii - cx, jj - cy, l - cz) + FT(0.5) * V(D, ii, jj, l);
            const FT pt = FT(0.5) * V(COMP, ii - dx, jj - dy, l - dz) + FT(0.5) * V(COMP, ii, jj, l);
            const FT F = A * ut * pt;
            const FT dl = (V(lo, ii, jj, l) - V(lo, ii - (hi == 0), jj - (hi == 1), l - (hi == 2))) * g.rd[hi];
            const FT dh = (V(hi, ii, jj, l) - V(hi, ii - (lo == 0), jj - (lo == 1), l - (lo == 2))) * g.rd[lo];
            const FT sig = FT(0.5) * (dl + dh);
            return F + (FT(-2) * a.nu * A) * sig;
        }
    }

    // -V⁻¹ (δx Fx + δy Fy + δz Fz) of component COMP at cell (ii, jj, k): centre-type fluxes live at index c and c-1, face-type at f and f+1
    template <int COMP>
    OC_HD FT divergence(const View& V, int ii, int jj, int k) const {
        const FT dFx = COMP == 0 ? flux<COMP, 0>(V, ii, jj, k) - flux<COMP, 0>(V, ii - 1, jj, k) : flux<COMP, 0>(V, ii + 1, jj, k) - flux<COMP, 0>(V, ii, jj, k);
        const FT dFy = COMP == 1 ? flux<COMP, 1>(V, ii, jj, k) - flux<COMP, 1>(V, ii, jj - 1, k) : flux<COMP, 1>(V, ii, jj + 1, k) - flux<COMP, 1>(V, ii, jj, k);
        const FT dFz = COMP == 2 ? flux<COMP, 2>(V, ii, jj, k) - flux<COMP, 2>(V, ii, jj, k - 1) : flux<COMP, 2>(V, ii, jj, k + 1) - flux<COMP, 2>(V, ii, jj, k);
        return -(a.g.rV * (dFx + dFy + dFz));
    }

    OC_HD void finish(int comp, int o, FT u0, FT G) const {
        a.Gn[comp][o] = G;
        if (a.mode == STEP_RK3_FIRST) {
            a.Unew[comp][o] = u0 + a.ca * G;
        } else if (a.mode == STEP_RK3) {
            a.Unew[comp][o] = u0 + a.dt * (a.ca * G + a.cb * a.Gm[comp][o]);
        } else if (a.mode == STEP_AB2) {
            const FT Gu = a.ab2_euler ? a.ca * G : a.ca * G - a.cb * a.Gm[comp][o];
            a.Unew[comp][o] = u0 + a.dt * Gu;
        }
    }

    template <int PHASE>
    OC_DEV void step(const Block& b, int tid, char* smem, int it, State& st) const {
        const Geom<FT>& g = a.g;
        const int n = st.nit;
        const int k = k_begin(b) + it;
        if (PHASE == 0) {
            if (it >= n) return;
            mbar_wait(bar_d(smem, it), (it / MARCH_NBAR) & 1);
            if (!st.own) return;
            const int lane = tid & (TX - 1), row = tid / TX;
            const int ii = lane, jj = row;
            const int j = (b.y + by0) * TY + row;
            const View V{ring(smem, 0, k, st.sk), ring(smem, 1, k, st.sk), ring(smem, 2, k, st.sk)};
            const int o = st.o;
            FT Gu = divergence<0>(V, ii, jj, k);
            FT Gv = divergence<1>(V, ii, jj, k);
            const FT Gw = divergence<2>(V, ii, jj, k);
            if (a.has_coriolis) {
                // FPlane / BetaPlane (f_plane.jl:50-52, beta_plane.jl:56-72): ∓ f ℑxy of the other horizontal component (no walls: every node active)
                const FT fu = a.has_coriolis == 2 ? a.f + a.cor_beta * (a.cor_y0 + (FT(j) + FT(0.5)) * g.d[1]) : a.f;
                const FT fv = a.has_coriolis == 2 ? a.f + a.cor_beta * (a.cor_y0 + FT(j) * g.d[1]) : a.f;
                const FT vu = FT(0.25) * ((V(1, ii - 1, jj, k) + V(1, ii, jj, k)) + (V(1, ii - 1, jj + 1, k) + V(1, ii, jj + 1, k)));
                const FT uv = FT(0.25) * ((V(0, ii, jj - 1, k) + V(0, ii + 1, jj - 1, k)) + (V(0, ii, jj, k) + V(0, ii + 1, jj, k)));
                Gu = Gu - (-fu * (vu / FT(1)));
                Gv = Gv - (fv * (uv / FT(1)));
            }
            if (a.pHY) {
                const FT p = a.pHY[o];
                Gu = Gu - (p - a.pHY[o - 1]) * g.rd[0];
                Gv = Gv - (p - a.pHY[o - g.sy]) * g.rd[1];
            }
            finish(0, o, V(0, ii, jj, k), Gu);
            finish(1, o, V(1, ii, jj, k), Gv);
            finish(2, o, V(2, ii, jj, k), Gw);
        } else if (PHASE == 1) {
            // every thread has finished iteration it-1 (sync_wait): the slot of level k-2, last read there, may be overwritten
            if (tid == 0) {
                const int lit = it + PF;
                if (lit < n) {
                    proxy_fence_async();
                    mbar_expect(bar_d(smem, lit), LEVEL_BYTES);
                    issue_level(smem, b.x * TX, (b.y + by0) * TY, k + PF + 1, bar_d(smem, lit));
                }
            }
        } else {
            st.sk = G::next_slot(st.sk);
            st.o += g.sz;
        }
    }
};

}  // namespace oc
