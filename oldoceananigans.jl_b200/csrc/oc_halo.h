// oc_halo.h — ONE fused halo-fill launch for any set of fields and all six sides.
//
// Replaces fill_halo_regions! and its ~3 launches per field (src/BoundaryConditions/fill_halo_regions.jl
// :25-108, fill_halo_regions_periodic.jl:5-32, _flux.jl:9-27, _value_gradient.jl:7-119, _open.jl:2-14) and
// the ordering rule of boundary_condition_ordering.jl:113-128 (non-periodic sides first, periodic sides
// afterwards over the whole parent extent so that corners receive periodic images of BC halos).
//
// Because every halo value is ultimately a function of ONE interior cell, the ordered multi-pass fill
// collapses to a single pass: for a halo cell P, wrap its periodic coordinates into the interior; if the
// result Q lies in the first halo plane of exactly one Bounded dimension (and inside 1:N in the other
// Bounded dimensions) apply that side's BC formula to the adjacent interior cell; halo planes 2..H of
// Bounded sides, and Bounded×Bounded corners, are never written — exactly like the reference.  Open
// (impenetrable) BCs set the wall-normal velocity ON the boundary face (index 1 and N+1).
#pragma once
#include "oc_common.h"

namespace oc {

enum { HALO_MAX_FIELDS = 24, HALO_MAX_BOXES = HALO_MAX_FIELDS * 9 };

// one side of one field, after default resolution
struct SideBC {
    int kind;      // oc_bc_kind: 1 periodic, 2 flux, 3 value, 4 gradient, 5 open, 6 none
    double value;
};

template <class FT>
struct HaloField {
    FT* p;
    int face[3];       // location: 1 = Face
    SideBC bc[6];
};

// A halo slab of one field.  Blocks tile it in (x, y) with 2^tshift × (256 >> tshift) threads — 32 × 8 for slabs that are long in x
// (coalesced rows), 4 × 64 for the west / east slabs that are only H cells wide — and one level of z per block, so that a thread's
// cell follows from shifts and ONE block-uniform division (the first version decoded a linear cell index with two divisions and
// three modulo operations per thread: ~60 % of the kernel's instructions, profiles/r02b_ncu_other_c3_summary.txt).
struct HaloBox {
    int field;
    int lo[3];
    int n[3];
    int first_block;   // prefix sum of blocks
    int tshift;        // log2 of the tile width in x
    int nbx, nby;      // tiles in x and y (and ceil(n[2] / ZPT) groups of levels)
};

template <class FT>
struct HaloKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    int nfields, nboxes;
    int fill_open;
    int skip[3];               // 1: halos of this (periodic) dimension are filled by neighbour exchange, not here (oc_dist.h)
    HaloField<FT> f[HALO_MAX_FIELDS];
    const HaloBox* boxes;      // device array [nboxes]

    // What one halo cell P of field `fld` needs: op 0 nothing, 1 store the wall value, 2 copy the interior cell `src` (through the BC
    // formula of side (bc_dim, bc_side) when bc_dim >= 0)
    struct Act {
        int op, src, dst, bc_dim, bc_side;
        FT wall;
    };
    OC_HD Act resolve(const HaloField<FT>& fld, const int P[3]) const {
        Act r{0, 0, 0, -1, 0, FT(0)};
        int Q[3];
        int bc_dim = -1, bc_side = 0, nbc = 0;
        bool wall = false;
        FT wall_value = FT(0);
        bool moved = false;
        for (int d = 0; d < 3; ++d) {
            int idx = P[d], N = g.N[d];
            if (!g.bounded[d]) {                       // Periodic (and Flat stored as periodic)
                if (skip[d] && (idx < 0 || idx >= N)) return r;
                int q = idx < 0 ? idx + N : (idx >= N ? idx - N : idx);          // one wrap is enough when N >= H …
                if ((unsigned)q >= (unsigned)N) { q %= N; if (q < 0) q += N; }    // … tiny grids (N < internal halo) wrap repeatedly
                if (q != idx) moved = true;
                Q[d] = q;
            } else if (fld.face[d]) {                  // wall-normal velocity: interior points 0..N
                if (idx < 0 || idx > N) return r;      // never filled (field_boundary_conditions.jl:15-25: Open only)
                if (idx == N && !g.whi[d]) return r;   // connected side of a slab: the neighbour's first face, by exchange
                Q[d] = idx;
                if ((idx == 0 && g.wlo[d]) || (idx == N && g.whi[d])) {
                    const SideBC& s = fld.bc[2 * d + (idx == 0 ? 0 : 1)];
                    if (s.kind == 5 && fill_open) { wall = true; wall_value = FT(s.value); }
                }
            } else {                                   // Center-located in a Bounded dimension
                if (idx >= 0 && idx < N) Q[d] = idx;
                else if (idx == -1 && g.wlo[d]) { Q[d] = 0; bc_dim = d; bc_side = 0; ++nbc; }
                else if (idx == N && g.whi[d]) { Q[d] = N - 1; bc_dim = d; bc_side = 1; ++nbc; }
                else return r;                         // halo planes 2..H are never written; connected sides of a slab: by exchange
            }
        }
        if (nbc > 1) return r;                         // Bounded×Bounded corners are never written
        if (nbc == 1 || wall) {
            // non-periodic fills cover the interior tangential range 1:N of the other dimensions only
            // (fill_halo_regions.jl:119-128); periodic coordinates were wrapped above.
            for (int d = 0; d < 3; ++d)
                if (g.bounded[d] && d != bc_dim && !(wall && fld.face[d] && (P[d] == 0 || P[d] == g.N[d]) && nbc == 0)) {
                    if (P[d] < 0 || P[d] >= g.N[d]) return r;
                }
        }
        if (nbc == 0 && !wall && !moved) return r;     // a plain interior cell: nothing to do
        r.dst = g.idx(P[0], P[1], P[2]);
        if (wall && nbc == 0) { r.op = 1; r.wall = wall_value; return r; }
        if (nbc == 1) {
            const int kind = fld.bc[2 * bc_dim + bc_side].kind;
            if (kind != 2 && kind != 3 && kind != 4) return r;                   // none
            r.bc_dim = bc_dim; r.bc_side = bc_side;
        }
        r.op = 2;
        r.src = g.idx(Q[0], Q[1], Q[2]);
        return r;
    }

    // ZPT consecutive z-levels per thread: their loads are independent and issued together (one load per thread left the kernel bound
    // by memory latency: 0.38 ms for 0.74 GB of sector traffic at 512³)
    static constexpr int ZPT = 4;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        // locate the box of this block (boxes are few: linear scan from the back)
        int bi = nboxes - 1;
        while (bi > 0 && boxes[bi].first_block > b.x) --bi;
        const HaloBox bx = boxes[bi];
        (void)nt;
        const unsigned lb = (unsigned)(b.x - bx.first_block);
        const unsigned per_plane = (unsigned)bx.nbx * (unsigned)bx.nby;
        const unsigned pl = lb / per_plane, rem = lb - pl * per_plane;
        const unsigned by = rem / (unsigned)bx.nbx, bxi = rem - by * (unsigned)bx.nbx;
        const int tx = tid & ((1 << bx.tshift) - 1), ty = tid >> bx.tshift;
        const int c0 = (int)(bxi << bx.tshift) + tx, c1 = (int)by * (THREADS >> bx.tshift) + ty;
        if (c0 >= bx.n[0] || c1 >= bx.n[1]) return;
        const HaloField<FT>& fld = f[bx.field];
        FT* p = fld.p;
        Act act[ZPT];
        FT v[ZPT];
#ifndef OC_HOSTSIM
#pragma unroll
#endif
        for (int z = 0; z < ZPT; ++z) {
            const int c2 = (int)pl * ZPT + z;
            int P[3] = {bx.lo[0] + c0, bx.lo[1] + c1, bx.lo[2] + c2};
            act[z] = c2 < bx.n[2] ? resolve(fld, P) : Act{0, 0, 0, -1, 0, FT(0)};
        }
#ifndef OC_HOSTSIM
#pragma unroll
#endif
        for (int z = 0; z < ZPT; ++z) v[z] = act[z].op == 2 ? p[act[z].src] : FT(0);
#ifndef OC_HOSTSIM
#pragma unroll
#endif
        for (int z = 0; z < ZPT; ++z) {
            const Act& a = act[z];
            if (a.op == 1) { p[a.dst] = a.wall; continue; }
            if (a.op != 2) continue;
            FT cI = v[z];
            if (a.bc_dim >= 0) {
                const SideBC& s = fld.bc[2 * a.bc_dim + a.bc_side];
                // Δ between the interior and the halo point, at the boundary face (fill_halo_regions_value_gradient.jl:44,60)
                FT delta = a.bc_dim == 2 ? g.dz_at(true, a.bc_side == 0 ? 0 : g.N[2]) : g.d[a.bc_dim];
                if (s.kind == 4) {
                    FT grad = FT(s.value);                                       // _value_gradient.jl:9-10
                    cI = cI + grad * (a.bc_side == 0 ? -delta : delta);
                } else if (s.kind == 3) {
                    FT val = FT(s.value);                                        // :12-13
                    if (a.bc_side == 0) { FT grad = (cI - val) / (delta / FT(2)); cI = cI + grad * (-delta); }
                    else { FT grad = (val - cI) / (delta / FT(2)); cI = cI + grad * delta; }
                }                                                                // kind 2, Flux: c[0] = c[1], c[N+1] = c[N]   fill_halo_regions_flux.jl:9-27
            }
            p[a.dst] = cI;
        }
    }
};

// Array-valued Value / Gradient boundary condition on ONE side of one Center-located field: ValueBoundaryCondition(A::AbstractArray),
// GradientBoundaryCondition(A) — getbc(bc, i, j, …) = A[i, j] in _fill_*_halo! (fill_halo_regions_value_gradient.jl:7-119).  Runs right
// after HaloKernel (which filled these cells with the side's scalar) and rewrites exactly the cells HaloKernel writes for that side: the
// first halo plane, over the interior range of Bounded tangential dimensions and the whole (wrapped) extent of periodic ones, so that
// corners receive the periodic images of the BC halos (boundary_condition_ordering.jl:113-128).
template <class FT>
struct HaloArrayKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    FT* p;
    const FT* A;       // n1 × n2 values over the tangential interior, first tangential dimension fastest
    int d, side, kind; // normal dimension; 0 low / 1 high; 3 value, 4 gradient
    int n1;
    int lo1, m1, lo2, m2;   // index ranges of the two tangential dimensions covered by this launch
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int a1 = b.x * nt + tid, a2 = b.y;
        if (a1 >= m1 || a2 >= m2) return;
        const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
        int P[3], Q[3];
        P[t1] = lo1 + a1; P[t2] = lo2 + a2;
        P[d] = side == 0 ? -1 : g.N[d];
        Q[d] = side == 0 ? 0 : g.N[d] - 1;
        for (int n = 0; n < 2; ++n) {
            const int t = n == 0 ? t1 : t2;
            int q = P[t];
            if (!g.bounded[t]) { q %= g.N[t]; if (q < 0) q += g.N[t]; }      // periodic (and Flat, stored as periodic with N = 1)
            Q[t] = q;
        }
        const FT val = A[Q[t1] + (size_t)n1 * Q[t2]];
        const FT delta = d == 2 ? g.dz_at(true, side == 0 ? 0 : g.N[2]) : g.d[d];
        FT cI = p[g.idx(Q[0], Q[1], Q[2])];
        if (kind == 4) cI = cI + val * (side == 0 ? -delta : delta);                          // _value_gradient.jl:9-10
        else if (side == 0) { const FT grad = (cI - val) / (delta / FT(2)); cI = cI + grad * (-delta); }   // :12-13
        else { const FT grad = (val - cI) / (delta / FT(2)); cI = cI + grad * delta; }
        p[g.idx(P[0], P[1], P[2])] = cI;
    }
};

}  // namespace oc
