"""Oracle: flux-form advection, Centered(order=2) and WENO(order=5).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates src/Advection/:
  reconstruction_coefficients.jl:49-64,87-89,122-152   stencil_coefficients / calc_reconstruction_stencil
  centered_reconstruction.jl:5-55, centered_advective_fluxes.jl:15-33
  upwind_biased_reconstruction.jl, upwind_biased_advective_fluxes.jl:21-121
  weno_reconstruction.jl:77-93, weno_interpolants.jl:71-83,117-137,169-174,204-216,261-266,290-338,409-437,500
  topologically_conditional_interpolation.jl:46-52,75-120
  momentum_advection_operators.jl:46-83, tracer_advection_operators.jl:30-34
  src/Utils/newton_div.jl:8-23
"""
from fractions import Fraction

import numpy as np

from .operators import O, dC, dF, sh

_err = dict(over="ignore", invalid="ignore", divide="ignore")


# ---------------------------------------------------------------------------------
# Coefficients
# ---------------------------------------------------------------------------------
def _round(FT, frac):
    return FT(float(frac)) if FT is np.float64 else np.float32(float(frac))


def stencil_coefficients(FT, r, order):
    """reconstruction_coefficients.jl:49-64 on a uniform grid (xr = xi = 1:100, i = 50):
    xr[i] - xi[i-(r-q+1)] = r-q+1 and xi[i-(r-m+1)] - xi[i-(r-l+1)] = m-l.
    The first order-1 coefficients are rounded to FT, the last is 1 - sum(others) in FT arithmetic."""
    FT = np.dtype(FT).type
    coeffs = []
    for j in range(order):
        c = Fraction(0)
        for m in range(j + 1, order + 1):
            num = Fraction(0)
            for l in range(order + 1):
                if l == m:
                    continue
                p = Fraction(1)
                for q in range(order + 1):
                    if q != m and q != l:
                        p *= (r - q + 1)
                num += p
            den = Fraction(1)
            for l in range(order + 1):
                if l != m:
                    den *= (m - l)
            c += num / den
        coeffs.append(c)
    rounded = [_round(FT, c) for c in coeffs[:-1]]
    s = FT(0)
    for c in rounded:          # Base.sum of a short Vector: left-to-right
        s = FT(s + c)
    return tuple(rounded) + (FT(FT(1) - s),)


def centered_coefficients(FT, buffer):
    """uniform_reconstruction_coefficients(FT, Val(:symmetric), buffer) :87
    returned in *stencil order* ψ[i-buffer] ... ψ[i+buffer-1] as used by calc_reconstruction_stencil:122-152
    (coefficient of the idx-th stencil point is coeff[order - idx + 1])."""
    order = 2 * buffer
    c = stencil_coefficients(FT, buffer - 1, order)
    return tuple(c[order - idx] for idx in range(1, order + 1))


class Centered:
    def __init__(self, FT=np.float64, order=2):
        assert order in (2, 4, 6, 8)
        self.FT = np.dtype(FT).type
        self.buffer = order // 2
        self.coeffs = centered_coefficients(self.FT, self.buffer)
        self.buffer_scheme = Centered(FT, order - 2) if order > 2 else None
        self.kind = "centered"


class UpwindBiased1:
    """UpwindBiased(order=1): buffer 1, symmetric part = Centered(2)"""

    def __init__(self, FT=np.float64):
        self.FT = np.dtype(FT).type
        self.buffer = 1
        self.advecting_velocity_scheme = Centered(FT, 2)
        self.buffer_scheme = None
        self.kind = "upwind1"


class UpwindBiased:
    """UpwindBiased(order = 1, 3, 5): buffer N = (order+1)/2, advecting_velocity_scheme = Centered(order-1) (Centered(2) for
    order 1), buffer_scheme = UpwindBiased(order-2)   upwind_biased_reconstruction.jl:57-86.
    coeff_left / coeff_right = uniform_reconstruction_coefficients(FT, Val(:left / :right), buffer)
    (reconstruction_coefficients.jl:87-89)."""

    def __init__(self, FT=np.float64, order=3):
        assert order in (1, 3, 5)
        self.FT = np.dtype(FT).type
        self.order = order
        self.buffer = (order + 1) // 2
        self.kind = "upwind"
        B = self.buffer
        if B > 1:
            self.advecting_velocity_scheme = Centered(FT, order - 1)
            self.buffer_scheme = UpwindBiased(FT, order - 2)
            self.coeff_left = stencil_coefficients(self.FT, B - 2, order)
            self.coeff_right = stencil_coefficients(self.FT, B - 1, order)
        else:
            self.advecting_velocity_scheme = Centered(FT, 2)
            self.buffer_scheme = None
            self.coeff_left = self.coeff_right = (self.FT(1),)


class NoAdvection:
    """advection = nothing: every advective flux is zero(grid) (momentum_advection_operators.jl:86-95, tracer_advection_operators.jl)"""

    def __init__(self, FT=np.float64):
        self.FT = np.dtype(FT).type
        self.buffer = 1
        self.kind = "none"


class FluxFormAdvection:
    """FluxFormAdvection(x, y, z) (src/Advection/flux_form_advection.jl): the scheme of direction d computes the whole flux through
    the d-faces — `_advective_momentum_flux_Uu(i, j, k, grid, advection::FluxFormAdvection, U, u) = _advective_momentum_flux_Uu(i, j, k,
    grid, advection.x, U, u)` and likewise for every component and direction."""

    def __init__(self, x, y, z):
        self.dirs = (x, y, z)
        self.FT = x.FT
        self.buffer = max(s.buffer for s in self.dirs)
        self.kind = "fluxform"


def scheme_of(scheme, d):
    return scheme.dirs[d] if getattr(scheme, "kind", None) == "fluxform" else scheme


def adapt_advection_order(advection, grid):
    """adapt_advection_order(advection, grid)  src/Advection/adapt_advection_order.jl:18-96 — Centered(2N), UpwindBiased(2N-1),
    WENO(2N-1) in the directions with N < buffer (WENO(1) = UpwindBiased(1): weno_reconstruction.jl:83-85); Flat directions untouched."""
    if advection.kind == "none":
        return advection
    FT = advection.FT
    dirs = [scheme_of(advection, d) for d in range(3)]
    changed = False
    for d in range(3):
        sch, N = dirs[d], grid.N[d]
        if grid.flat(d) or N >= sch.buffer:
            continue
        if sch.kind == "centered":
            new = Centered(FT, 2 * N)
        elif sch.kind in ("upwind", "upwind1"):
            new = UpwindBiased(FT, 2 * N - 1)
        else:
            new = WENO(FT, 2 * N - 1) if 2 * N - 1 >= 3 else UpwindBiased(FT, 1)
        dirs[d], changed = new, True
    return FluxFormAdvection(*dirs) if changed else advection


def _upwind_value(sch, S, left):
    """biased_interpolate for UpwindBiased{B}: S = (ψ[i-B], …, ψ[i+B-1]); calc_reconstruction_stencil
    (reconstruction_coefficients.jl:122-152): the idx-th point of the left stencil is ψ[i+idx-B-1], of the right stencil
    ψ[i+idx-B], each with coefficient coeff[order-idx+1]; summed left to right."""
    B, order = sch.buffer, sch.order
    L = R = None
    for idx in range(1, order + 1):
        tl = sch.coeff_left[order - idx] * S[idx - 1]
        tr = sch.coeff_right[order - idx] * S[idx]
        L = tl if L is None else L + tl
        R = tr if R is None else R + tr
    return np.where(left, L, R)


class WENO:
    """WENO{N,FT,Float32}: order 5 (buffer 3) or 3 (buffer 2).  weno_reconstruction.jl:77-93"""

    # smoothness_coefficients weno_interpolants.jl:169-174
    SMOOTH = {2: ((1, -2, 1), (1, -2, 1)),
              3: ((10, -31, 11, 25, -19, 4), (4, -13, 5, 13, -13, 4), (4, -19, 11, 25, -31, 10)),
              # weno_interpolants.jl:175-185 (Float64 decimal literals converted to FT)
              4: ((2.107, -9.402, 7.042, -1.854, 11.003, -17.246, 4.642, 7.043, -3.882, 0.547),
                  (0.547, -2.522, 1.922, -0.494, 3.443, -5.966, 1.602, 2.843, -1.642, 0.267),
                  (0.267, -1.642, 1.602, -0.494, 2.843, -5.966, 1.922, 3.443, -2.522, 0.547),
                  (0.547, -3.882, 4.642, -1.854, 7.043, -17.246, 7.042, 11.003, -9.402, 2.107)),
              5: ((1.07918, -6.49501, 7.58823, -4.11487, 0.86329, 10.20563, -24.62076, 13.58458, -2.88007, 15.21393, -17.04396, 3.64863, 4.82963, -2.08501, 0.22658),
                  (0.22658, -1.40251, 1.65153, -0.88297, 0.18079, 2.42723, -6.11976, 3.37018, -0.70237, 4.06293, -4.64976, 0.99213, 1.38563, -0.60871, 0.06908),
                  (0.06908, -0.51001, 0.67923, -0.38947, 0.08209, 1.04963, -2.99076, 1.79098, -0.38947, 2.31153, -2.99076, 0.67923, 1.04963, -0.51001, 0.06908),
                  (0.06908, -0.60871, 0.99213, -0.70237, 0.18079, 1.38563, -4.64976, 3.37018, -0.88297, 4.06293, -6.11976, 1.65153, 2.42723, -1.40251, 0.22658),
                  (0.22658, -2.08501, 3.64863, -2.88007, 0.86329, 4.82963, -17.04396, 13.58458, -4.11487, 15.21393, -24.62076, 7.58823, 10.20563, -6.49501, 1.07918))}
    CSTAR = {2: (Fraction(2, 3), Fraction(1, 3)),
             3: (Fraction(3, 10), Fraction(3, 5), Fraction(1, 10)),
             4: (Fraction(4, 35), Fraction(18, 35), Fraction(12, 35), Fraction(1, 35)),
             5: (Fraction(5, 126), Fraction(20, 63), Fraction(10, 21), Fraction(10, 63), Fraction(1, 126))}
    # global_smoothness_indicator  weno_interpolants.jl:303-307: τ = |Σ TAU[r] β_r|, summed left to right
    TAU = {2: (1, -1), 3: (1, 0, -1), 4: (1, 3, -3, -1), 5: (1, 2, -6, 2, 1)}

    def __init__(self, FT=np.float64, order=5):
        assert order in (3, 5, 7, 9)
        self.FT = np.dtype(FT).type
        self.buffer = (order + 1) // 2
        self.kind = "weno"
        self.advecting_velocity_scheme = Centered(FT, order - 1)
        self.buffer_scheme = WENO(FT, order - 2) if order > 3 else UpwindBiased1(FT)
        B = self.buffer
        self.coeff_p = tuple(stencil_coefficients(self.FT, r, B) for r in range(B))   # :117-118
        self.cstar = tuple(_round(self.FT, c) for c in self.CSTAR[B])                # :78-83
        self.smooth = tuple(tuple(self.FT(c) for c in row) for row in self.SMOOTH[B])
        self.eps = self.FT(np.float32(1e-8))                                          # ϵ = 1f-8 :71


def newton_div(FT, a, b):
    """newton_div(Float32, a, b::FT)   src/Utils/newton_div.jl:8-23"""
    if FT is np.float32:
        return a * (np.float32(1) / b)                    # :23  a * inv_fast(b)
    inv_b = (np.float32(1) / b.astype(np.float32)).astype(np.float64)
    x = a * inv_b
    # x = fma(fma(x, -b, a), inv_b, x): emulate the two fused ops in extended precision
    r = (a.astype(np.longdouble) - x.astype(np.longdouble) * b.astype(np.longdouble)).astype(np.float64)
    x = (r.astype(np.longdouble) * inv_b.astype(np.longdouble) + x.astype(np.longdouble)).astype(np.float64)
    return x


def _weno_value(sch, S, left):
    """biased_interpolate for WENO{B}: S = (ψ[i-B], …, ψ[i+B-1]); ``left`` boolean array (bias(ũ) isa LeftBias).
    weno_interpolants.jl:432-437 (stencils), :204-216,261 (β), :308-309 (τ), :290-297 (α), :332-338 (ω), :500."""
    B, FT = sch.buffer, sch.FT
    # sub-stencils r = 0..B-1 : Left (S[B-r .. 2B-1-r]) ; Right (S[B+r], S[B+r-1], …)  (1-based in the source)
    stencils = []
    for r in range(B):
        L = [S[B - 1 - r + j] for j in range(B)]
        R = [S[B + r - j] for j in range(B)]
        stencils.append([np.where(left, L[j], R[j]) for j in range(B)])
    betas = []
    for r in range(B):
        psi, C = stencils[r], sch.smooth[r]
        beta, c = None, 0
        for s in range(B - 1):
            inner = None
            for i in range(s, B):
                term = C[c + i - s] * psi[i]
                inner = term if inner is None else inner + term
            c += B - s
            term = psi[s] * inner
            beta = term if beta is None else beta + term
        beta = beta + psi[B - 1] * psi[B - 1] * C[c]
        betas.append(beta)
    acc = None                                           # global_smoothness_indicator :303-307 (β[1] + 3β[2] - 3β[3] - β[4] …)
    for r, t in enumerate(sch.TAU[B]):
        if t == 0:
            continue
        term = betas[r] if abs(t) == 1 else FT(abs(t)) * betas[r]
        acc = term if acc is None else (acc + term if t > 0 else acc - term)
    tau = np.abs(acc)
    alphas = []
    for r in range(B):
        q = newton_div(FT, tau, betas[r] + sch.eps)
        alphas.append(sch.cstar[r] * (FT(1) + q * q))
    ssum = alphas[0]
    for r in range(1, B):
        ssum = ssum + alphas[r]
    rs = FT(1) / ssum
    out = None
    for r in range(B):
        w = alphas[r] * rs
        p = None
        for j in range(B):                               # sum(coeff .* ψ): left to right
            t = sch.coeff_p[r][j] * stencils[r][j]
            p = t if p is None else p + t
        t = w * p
        out = t if out is None else out + t
    return out


def _centered_value(sch, S):
    """symmetric_interpolate for Centered{B}: S = (ψ[i-B], …, ψ[i+B-1])"""
    out = None
    for c, s in zip(sch.coeffs, S):
        t = c * s
        out = t if out is None else out + t
    return out


def _inside(ctx, d, o, lo, hi):
    idx = ctx.index(d, o)
    return (idx >= lo) & (idx <= hi)


def symmetric_face(ctx, scheme, q, d):
    """_symmetric_interpolate_ξᶠ(i,j,k, grid, scheme, q): face-type symmetric interpolation with the
    boundary fallback chain of topologically_conditional_interpolation.jl:46-52,99-120."""
    g = ctx.g

    def at(o, sch):
        if sch.kind != "centered":
            inner = sch.advecting_velocity_scheme
        else:
            inner = sch
        Bc = inner.buffer
        S = [q(sh(o, d, n)) for n in range(-Bc, Bc)]
        val = _centered_value(inner, S)
        if g.bounded(d) and sch.buffer > 1:
            R, N = sch.buffer, g.N[d]
            ok = _inside(ctx, d, o, R + 1, N + 1 - R)       # outside_symmetric_haloᶠ
            return np.where(ok, val, at(o, sch.buffer_scheme))
        return val

    if g.flat(d):
        return q                                            # flat_advective_fluxes / identity
    def run(o):
        with np.errstate(**_err):
            return at(o, scheme)
    return run


def symmetric_center(ctx, scheme, q, d):
    """_symmetric_interpolate_ξᶜ(i) = face-type at i+1 with the centre-type bounds check"""
    g = ctx.g

    def at(o, sch):
        inner = sch.advecting_velocity_scheme if sch.kind != "centered" else sch
        Bc = inner.buffer
        o1 = sh(o, d, 1)
        S = [q(sh(o1, d, n)) for n in range(-Bc, Bc)]
        val = _centered_value(inner, S)
        if g.bounded(d) and sch.buffer > 1:
            R, N = sch.buffer, g.N[d]
            ok = _inside(ctx, d, o, R, N + 1 - R)           # outside_symmetric_haloᶜ
            return np.where(ok, val, at(o, sch.buffer_scheme))
        return val

    if g.flat(d):
        return q
    def run(o):
        with np.errstate(**_err):
            return at(o, scheme)
    return run


def _biased(ctx, scheme, q, d, left_q, center_type):
    """_biased_interpolate_ξᶠ / ξᶜ with boundary fallback WENO5 -> WENO3 -> Upwind1."""
    g = ctx.g

    def at(o, sch, left):
        oo = sh(o, d, 1) if center_type else o
        if sch.kind == "upwind1":
            # calc_reconstruction_stencil(FT, 1, :left/:right): 1*ψ[i-1] (left), 1*ψ[i] (right)
            return np.where(left, q(sh(oo, d, -1)), q(oo))
        if sch.kind == "centered":
            B = sch.buffer
            return _centered_value(sch, [q(sh(oo, d, n)) for n in range(-B, B)])
        B = sch.buffer
        S = [q(sh(oo, d, n)) for n in range(-B, B)]
        val = _upwind_value(sch, S, left) if sch.kind == "upwind" else _weno_value(sch, S, left)
        if g.bounded(d) and sch.buffer_scheme is not None:
            R, N = B, g.N[d]
            if center_type:   # outside_biased_haloᶜ
                lo, hi = max(R, R - 1), min(N + 1 - (R - 1), N + 1 - R)
            else:             # outside_biased_haloᶠ
                lo, hi = max(R + 1, R), min(N + 1 - (R - 1), N + 1 - R)
            ok = _inside(ctx, d, o, lo, hi)
            return np.where(ok, val, at(o, sch.buffer_scheme, left))
        return val

    def run(o):
        with np.errstate(**_err):
            return at(o, scheme, left_q(o))
    return run


def biased_face(ctx, scheme, q, d, left_q):
    return _biased(ctx, scheme, q, d, left_q, False)


def biased_center(ctx, scheme, q, d, left_q):
    return _biased(ctx, scheme, q, d, left_q, True)


# ---------------------------------------------------------------------------------
# Advective fluxes and flux divergences
# ---------------------------------------------------------------------------------
def _zero_q(ctx):
    return lambda o: ctx.zeros()


def momentum_flux(ctx, scheme, U, comp, d, psi_f):
    """advective_momentum_flux_{U,V,W}{u,v,w}: flux of velocity component ``comp`` (0,1,2) carried
    in direction ``d`` by U[d].  Returns a quantity.  Flux location: centre-type in d if d == comp
    (ccc), otherwise face-type in both d and comp."""
    g, FT = ctx.g, ctx.FT
    scheme = scheme_of(scheme, d)
    if g.flat(d) or scheme.kind == "none":
        return _zero_q(ctx)                                  # flat_advective_fluxes.jl:13-29 ; advection = nothing
    adv = ctx.field(U[d])
    psi = ctx.field(psi_f)
    if scheme.kind == "centered":
        # centered_advective_fluxes.jl:15-27 :  A * sym(U) * sym(ψ), A at the flux point (z-location of the stepped field)
        A = ctx.area(d, "f" if comp == 2 else "c")
        if d == comp:
            ut = symmetric_center(ctx, scheme, adv, d)
            pt = symmetric_center(ctx, scheme, psi, d)
        else:
            ut = symmetric_face(ctx, scheme, adv, comp)      # interpolate advecting velocity along comp
            pt = symmetric_face(ctx, scheme, psi, d)         # interpolate advected along d
        return lambda o: A(o) * ut(o) * pt(o)
    # upwind_biased_advective_fluxes.jl:23-93 :  ũ = sym(A*U) ; ψᴿ = biased(ψ; bias(ũ)) ; ũ*ψᴿ
    # (Ax_qᶠᶜᶜ, Ay_qᶜᶠᶜ, Az_qᶜᶜᶠ: the area at the advecting velocity's own point, inside the interpolation)
    A = ctx.area(d, "f" if d == 2 else "c")
    aq = lambda o: A(o) * adv(o)
    if d == comp:
        ut = symmetric_center(ctx, scheme, aq, d)
        pt = biased_center(ctx, scheme, psi, d, lambda o: ut(o) > 0)
    else:
        ut = symmetric_face(ctx, scheme, aq, comp)
        pt = biased_face(ctx, scheme, psi, d, lambda o: ut(o) > 0)
    return lambda o: ut(o) * pt(o)


def div_momentum(ctx, scheme, U, comp):
    """div_𝐯u / div_𝐯v / div_𝐯w   momentum_advection_operators.jl:46-83"""
    g = ctx.g
    total = None
    for d in range(3):
        F = momentum_flux(ctx, scheme, U, comp, d, U[comp])
        term = (dF(ctx, F, d) if d == comp else dC(ctx, F, d))(O)
        total = term if total is None else total + term
    return ctx.rvol("f" if comp == 2 else "c")(O) * total


def tracer_flux(ctx, scheme, U, c_f, d):
    """advective_tracer_flux_{x,y,z}"""
    g = ctx.g
    scheme = scheme_of(scheme, d)
    if g.flat(d) or scheme.kind == "none":
        return _zero_q(ctx)
    u = ctx.field(U[d])
    c = ctx.field(c_f)
    A = ctx.area(d, "f" if d == 2 else "c")                  # Axᶠᶜᶜ, Ayᶜᶠᶜ, Azᶜᶜᶠ
    if scheme.kind == "centered":
        ct = symmetric_face(ctx, scheme, c, d)
        return lambda o: (A(o) * u(o)) * ct(o)               # Ax_q(U) * sym(c)  :31-33
    ct = biased_face(ctx, scheme, c, d, lambda o: u(o) > 0)
    return lambda o: A(o) * u(o) * ct(o)                     # Ax * ũ * cᴿ  :99-121


def div_tracer(ctx, scheme, U, c_f):
    """div_Uc   tracer_advection_operators.jl:30-34"""
    g = ctx.g
    total = None
    for d in range(3):
        term = dC(ctx, tracer_flux(ctx, scheme, U, c_f, d), d)(O)
        total = term if total is None else total + term
    return ctx.rvol("c")(O) * total
