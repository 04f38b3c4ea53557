"""Oracle: turbulence closures (ScalarDiffusivity, AnisotropicMinimumDissipation, Smagorinsky / SmagorinskyLilly), buoyancy, Coriolis.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  src/TurbulenceClosures/closure_kernel_operators.jl:22-48
  src/TurbulenceClosures/abstract_scalar_diffusivity_closure.jl:189-204,240-242,310-332
  src/TurbulenceClosures/velocity_tracer_gradients.jl:6-46,126-250
  src/TurbulenceClosures/turbulence_closure_implementations/scalar_diffusivity.jl:195-198
  src/TurbulenceClosures/turbulence_closure_implementations/anisotropic_minimum_dissipation.jl:154-351
  src/TurbulenceClosures/turbulence_closure_implementations/Smagorinskys/smagorinsky.jl:31-158,
      Smagorinskys/lilly_coefficient.jl:47-135, Smagorinskys/scale_invariant_operators.jl:10-13
  src/BuoyancyFormulations/linear_equation_of_state.jl:72-80, buoyancy_tracer.jl:12-16, seawater_buoyancy.jl:219-224, g_dot_b.jl:2-4
  src/Coriolis/f_plane.jl:50-52, src/Operators/interpolation_operators.jl:119-131, src/Grids/inactive_node.jl
"""
import numpy as np

from .operators import O, dC, dF, ddC, ddF, iC, iF, sh


class ScalarDiffusivity:
    """ScalarDiffusivity(ν, κ): numbers, or arrays at (Center, Center, Center) — `νᶜᶜᶜ(…, ν::AbstractArray, …) = ν[i, j, k]`, `νᶠᶠᶜ = ℑxy ν`,
    `κᶠᶜᶜ = ℑx κ` … (abstract_scalar_diffusivity_closure.jl:323-332): the model turns array coefficients into Center fields whose halos
    follow the default boundary conditions, and the flux operators interpolate them like eddy viscosities."""

    def __init__(self, nu=0.0, kappa=0.0):
        self.nu, self.kappa = nu, kappa       # kappa: number | array, or dict tracer -> number | array
        self.kind = "scalar"
        vals = [nu] + (list(kappa.values()) if isinstance(kappa, dict) else [kappa])
        self.is_array = any(isinstance(v, np.ndarray) for v in vals)
        self.nu_field, self.kappa_fields = None, {}

    def kappa_for(self, name):
        return self.kappa[name] if isinstance(self.kappa, dict) else self.kappa


class AnisotropicMinimumDissipation:
    def __init__(self, C=1.0 / 3.0, Cnu=None, Ckappa=None, Cb=None):
        self.Cnu = C if Cnu is None else Cnu
        self.Ckappa = C if Ckappa is None else Ckappa
        self.Cb = Cb          # buoyancy modification multiplier; None turns the term off  (anisotropic_minimum_dissipation.jl:62-68)
        self.kind = "amd"

    def Ckappa_for(self, name):
        return self.Ckappa[name] if isinstance(self.Ckappa, dict) else self.Ckappa


class Smagorinsky:
    """Smagorinsky(coefficient=0.16, Pr=1.0)  smagorinsky.jl:31-84; with Cb given: SmagorinskyLilly(C, Cb, Pr), i.e.
    Smagorinsky(coefficient=LillyCoefficient(smagorinsky=C, reduction_factor=Cb))  lilly_coefficient.jl:47-112"""

    def __init__(self, coefficient=0.16, Pr=1.0, Cb=None):
        self.C, self.Pr, self.Cb = coefficient, Pr, Cb
        self.kind = "smagorinsky"

    def Pr_for(self, name):
        return self.Pr[name] if isinstance(self.Pr, dict) else self.Pr


def SmagorinskyLilly(C=0.16, Cb=1.0, Pr=1.0):
    return Smagorinsky(coefficient=C, Pr=Pr, Cb=Cb)


# ---------------------------------------------------------------------------------
# Viscosity / diffusivity "extractors"   abstract_scalar_diffusivity_closure.jl:310-332
# ---------------------------------------------------------------------------------
def _nu_at(ctx, nu, loc):
    """nu: number or quantity of a ccc field; loc in {'ccc','ffc','fcf','cff'}"""
    if not callable(nu):
        v = ctx.FT(nu)
        return lambda o: v
    q = nu
    # ℑxyᶠᶠᵃ = ℑyᶠ(ℑxᶠ(·)), ℑxzᶠᵃᶠ = ℑzᶠ(ℑxᶠ(·)), ℑyzᵃᶠᶠ = ℑzᶠ(ℑyᶠ(·))  (interpolation_operators.jl:45-56):
    # the lower dimension is the inner operator.
    for d in [d for d in range(3) if loc[d] == "f"]:
        q = iF(ctx, q, d)
    return q


def _kappa_at(ctx, kappa, d):
    if not callable(kappa):
        v = ctx.FT(kappa)
        return lambda o: v
    return iF(ctx, kappa, d)


def _strain_offdiag(ctx, U, a, b):
    """Σ_ab (a != b) at the location that is Face in a and b:  0.5 (∂b u_a + ∂a u_b)
    velocity_tracer_gradients.jl:31-42"""
    h = ctx.FT(0.5)
    lo, hi = (a, b) if a < b else (b, a)
    # source order: Σ12 = 0.5(∂y u + ∂x v); Σ13 = 0.5(∂z u + ∂x w); Σ23 = 0.5(∂z v + ∂y w)
    first = ddF(ctx, ctx.field(U[lo]), hi)
    second = ddF(ctx, ctx.field(U[hi]), lo)
    return lambda o: h * (first(o) + second(o))


def div_tau(ctx, closure_nu, U, comp):
    """∂ⱼ τ_{comp,j}   closure_kernel_operators.jl:22-41 with τ = -2 ν Σ  (:189-204)"""
    g, FT = ctx.g, ctx.FT
    total = None
    two = FT(2)
    for d in range(3):
        if d == comp:
            loc = "ccc"
            sig = ddC(ctx, ctx.field(U[comp]), comp)               # Σ_aa = ∂a u_a at ccc
        else:
            l = ["c", "c", "c"]
            l[d] = "f"
            l[comp] = "f"
            loc = "".join(l)
            sig = _strain_offdiag(ctx, U, comp, d)
        nu = _nu_at(ctx, closure_nu, loc)
        A = ctx.area(d, loc[2])                  # Ax_qᶜᶜᶜ/Ay_qᶠᶠᶜ/Az_qᶠᶜᶠ … at the flux point (closure_kernel_operators.jl:22-41)
        flux = (lambda nu, sig, A: (lambda o: A(o) * (-two * (nu(o) * sig(o)))))(nu, sig, A)
        term = (dF(ctx, flux, d) if d == comp else dC(ctx, flux, d))(O)
        total = term if total is None else total + term
    return ctx.rvol("f" if comp == 2 else "c")(O) * total


def div_q(ctx, kappa, c_f):
    """∇·q_c   closure_kernel_operators.jl:43-48 with q = -κ ∂c (:240-242)"""
    g = ctx.g
    c = ctx.field(c_f)
    total = None
    for d in range(3):
        kap = _kappa_at(ctx, kappa, d)
        grad = ddF(ctx, c, d)
        A = ctx.area(d, "f" if d == 2 else "c")
        flux = (lambda kap, grad, A: (lambda o: A(o) * (-(kap(o) * grad(o)))))(kap, grad, A)
        term = dC(ctx, flux, d)(O)
        total = term if total is None else total + term
    return ctx.rvol("c")(O) * total


# ---------------------------------------------------------------------------------
# AMD   anisotropic_minimum_dissipation.jl:154-351 ; velocity_tracer_gradients.jl:126-250
# ---------------------------------------------------------------------------------
def _sq(q):
    return lambda o: q(o) ** 2


def _mul(a, b):
    return lambda o: a(o) * b(o)


class _AMD:
    def __init__(self, ctx, U):
        g, FT = ctx.g, ctx.FT
        self.ctx = ctx
        u, v, w = (ctx.field(f) for f in U)
        # Δᶠ = 2Δ (:224-226); every Δᶠz_{loc}(i,j,k) is 2·Δzᶜᶜᶜ(i,j,k) at the index of the evaluation point (:228-234)
        Dfx, Dfy = FT(2) * g.dx, FT(2) * g.dy
        _dzc = ctx.dz("c")
        Dfz = lambda o: FT(2) * _dzc(o)
        self.Df = (Dfx, Dfy, Dfz)
        # normalised gradients (velocity_tracer_gradients.jl:126-154)
        self.dxu = ddC(ctx, u, 0)
        self.dyv = ddC(ctx, v, 1)
        self.dzw = ddC(ctx, w, 2)
        rxy, ryx = FT(Dfx / Dfy), FT(Dfy / Dfx)
        _dxv, _dyu = ddF(ctx, v, 0), ddF(ctx, u, 1)
        _dxw, _dzu = ddF(ctx, w, 0), ddF(ctx, u, 2)
        _dyw, _dzv = ddF(ctx, w, 1), ddF(ctx, v, 2)
        self.dxv = lambda o: rxy * _dxv(o)       # ffc
        self.dyu = lambda o: ryx * _dyu(o)       # ffc
        self.dxw = lambda o: (Dfx / Dfz(o)) * _dxw(o)       # fcf
        self.dzu = lambda o: (Dfz(o) / Dfx) * _dzu(o)       # fcf
        self.dyw = lambda o: (Dfy / Dfz(o)) * _dyw(o)       # cff
        self.dzv = lambda o: (Dfz(o) / Dfy) * _dzv(o)       # cff
        h = FT(0.5)
        self.S12 = lambda o: h * (self.dyu(o) + self.dxv(o))
        self.S13 = lambda o: h * (self.dzu(o) + self.dxw(o))
        self.S23 = lambda o: h * (self.dzv(o) + self.dyw(o))

    # ℑxyᶜᶜᵃ = ℑyᶜ(ℑxᶜ(·)),  ℑxzᶜᵃᶜ = ℑzᶜ(ℑxᶜ(·)),  ℑyzᵃᶜᶜ = ℑzᶜ(ℑyᶜ(·))  (interpolation_operators.jl:45-56)
    def Ixy(self, q):
        return iC(self.ctx, iC(self.ctx, q, 0), 1)

    def Ixz(self, q):
        return iC(self.ctx, iC(self.ctx, q, 0), 2)

    def Iyz(self, q):
        return iC(self.ctx, iC(self.ctx, q, 1), 2)

    def q_trace(self):
        """norm_tr_∇uᶜᶜᶜ  :285-306"""
        t = self.dxu(O) ** 2
        t = t + self.dyv(O) ** 2
        t = t + self.dzw(O) ** 2
        t = t + self.Ixy(_sq(self.dxv))(O)
        t = t + self.Ixy(_sq(self.dyu))(O)
        t = t + self.Ixz(_sq(self.dxw))(O)
        t = t + self.Ixz(_sq(self.dzu))(O)
        t = t + self.Iyz(_sq(self.dyw))(O)
        t = t + self.Iyz(_sq(self.dzv))(O)
        return t

    def r_term(self):
        """norm_uᵢₐ_uⱼₐ_Σᵢⱼᶜᶜᶜ  :240-279 (S11 = dxu, S22 = dyv, S33 = dzw)"""
        FT = self.ctx.FT
        two = FT(2)
        S11, S22, S33 = self.dxu(O), self.dyv(O), self.dzw(O)
        Ixy, Ixz, Iyz = self.Ixy, self.Ixz, self.Iyz
        t1 = S11 * self.dxu(O) ** 2
        t1 = t1 + S22 * Ixy(_sq(self.dxv))(O)
        t1 = t1 + S33 * Ixz(_sq(self.dxw))(O)
        t1 = t1 + two * self.dxu(O) * Ixy(_mul(self.dxv, self.S12))(O)
        t1 = t1 + two * self.dxu(O) * Ixz(_mul(self.dxw, self.S13))(O)
        t1 = t1 + two * Ixy(self.dxv)(O) * Ixz(self.dxw)(O) * Iyz(self.S23)(O)

        t2 = S11 * Ixy(_sq(self.dyu))(O)
        t2 = t2 + S22 * self.dyv(O) ** 2
        t2 = t2 + S33 * Iyz(_sq(self.dyw))(O)
        t2 = t2 + two * self.dyv(O) * Ixy(_mul(self.dyu, self.S12))(O)
        t2 = t2 + two * Ixy(self.dyu)(O) * Iyz(self.dyw)(O) * Ixz(self.S13)(O)
        t2 = t2 + two * self.dyv(O) * Iyz(_mul(self.dyw, self.S23))(O)

        t3 = S11 * Ixz(_sq(self.dzu))(O)
        t3 = t3 + S22 * Iyz(_sq(self.dzv))(O)
        t3 = t3 + S33 * self.dzw(O) ** 2
        t3 = t3 + two * Ixz(self.dzu)(O) * Iyz(self.dzv)(O) * Ixy(self.S12)(O)
        t3 = t3 + two * self.dzw(O) * Ixz(_mul(self.dzu, self.S13))(O)
        t3 = t3 + two * self.dzw(O) * Iyz(_mul(self.dzv, self.S23))(O)
        return t1 + t2 + t3

    def Cb_wi_bi(self, Cb, b):
        """Cb_norm_wᵢ_bᵢᶜᶜᶜ  :308-323 ; b = buoyancy_perturbationᶜᶜᶜ as a quantity (None without buoyancy: ∂b = 0)"""
        ctx, FT = self.ctx, self.ctx.FT
        if Cb is None or b is None:
            return FT(0)
        Dfx, Dfy, Dfz = self.Df
        bx = iC(ctx, ddF(ctx, b, 0), 0)(O)          # ℑxᶜᵃᵃ ∂xᶠᶜᶜ b
        by = iC(ctx, ddF(ctx, b, 1), 1)(O)          # ℑyᵃᶜᵃ ∂yᶜᶠᶜ b
        bz = iC(ctx, ddF(ctx, b, 2), 2)(O)          # ℑzᵃᵃᶜ ∂zᶜᶜᶠ b
        wx_bx = self.Ixz(self.dxw)(O) * Dfx * bx
        wy_by = self.Iyz(self.dyw)(O) * Dfy * by
        wz_bz = self.dzw(O) * Dfz(O) * bz            # norm_∂z_w = ∂z_w: the diagonal terms carry no width ratio (velocity_tracer_gradients.jl:126-128)
        return FT(Cb) * (wx_bx + wy_by + wz_bz)

    def delta2(self):
        FT = self.ctx.FT
        Dfx, Dfy, Dfz = self.Df
        return FT(3) / (FT(1) / Dfx ** 2 + FT(1) / Dfy ** 2 + FT(1) / Dfz(O) ** 2)

    def tracer_terms(self, c_f):
        ctx = self.ctx
        c = ctx.field(c_f)
        Dfx, Dfy, Dfz = self.Df
        _cx, _cy, _cz = ddF(ctx, c, 0), ddF(ctx, c, 1), ddF(ctx, c, 2)
        cx = lambda o: Dfx * _cx(o)
        cy = lambda o: Dfy * _cy(o)
        cz = lambda o: Dfz(o) * _cz(o)
        Ix = lambda q: iC(ctx, q, 0)
        Iy = lambda q: iC(ctx, q, 1)
        Iz = lambda q: iC(ctx, q, 2)
        sigma = Ix(_sq(cx))(O) + Iy(_sq(cy))(O) + Iz(_sq(cz))(O)          # norm_θᵢ²ᶜᶜᶜ :349-351
        Ixy, Ixz, Iyz = self.Ixy, self.Ixz, self.Iyz
        a = self.dxu(O) * Ix(_sq(cx))(O)
        a = a + Ixy(self.dxv)(O) * Ix(cx)(O) * Iy(cy)(O)
        a = a + Ixz(self.dxw)(O) * Ix(cx)(O) * Iz(cz)(O)
        b = Ixy(self.dyu)(O) * Iy(cy)(O) * Ix(cx)(O)
        b = b + self.dyv(O) * Iy(_sq(cy))(O)
        b = b + Ixz(self.dyw)(O) * Iy(cy)(O) * Iz(cz)(O)       # sic: ℑxzᶜᵃᶜ on norm_∂y_w (:336)
        cc = Ixz(self.dzu)(O) * Iz(cz)(O) * Ix(cx)(O)
        cc = cc + Iyz(self.dzv)(O) * Iz(cz)(O) * Iy(cy)(O)
        cc = cc + self.dzw(O) * Iz(_sq(cz))(O)
        return sigma, a + b + cc


def compute_amd(ctx, closure, U, tracers, nu_e, kappa_e, buoyancy=None):
    """_compute_AMD_viscosity! / _compute_AMD_diffusivity!  :154-197 over the window of ctx (interior)."""
    g, FT = ctx.g, ctx.FT
    amd = _AMD(ctx, U)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = amd.q_trace()
        r = amd.r_term()
        d2 = amd.delta2()
        b = buoyancy_q(ctx, buoyancy, tracers) if (buoyancy is not None and getattr(closure, "Cb", None) is not None) else None
        Cb_zeta = amd.Cb_wi_bi(getattr(closure, "Cb", None), b) / amd.Df[2](O)
        nu = -FT(closure.Cnu) * d2 * (r - Cb_zeta) / q
        nu = np.where(q == 0, FT(0), nu)
        nu_e.interior[...] = np.maximum(FT(0), nu)
        for name, c in tracers.items():
            sigma, theta = amd.tracer_terms(c)
            kap = -FT(closure.Ckappa_for(name)) * d2 * theta / sigma
            kap = np.where(sigma == 0, FT(0), kap)
            kappa_e[name].interior[...] = np.maximum(FT(0), kap)


# ---------------------------------------------------------------------------------
# Smagorinsky / SmagorinskyLilly   Smagorinskys/smagorinsky.jl:92-125, lilly_coefficient.jl:114-135
# ---------------------------------------------------------------------------------
def strain_double_dot_ccc(ctx, U):
    """ΣᵢⱼΣᵢⱼᶜᶜᶜ = tr_Σ² + 2 ℑxyᶜᶜᵃ Σ₁₂² + 2 ℑxzᶜᵃᶜ Σ₁₃² + 2 ℑyzᵃᶜᶜ Σ₂₃²   scale_invariant_operators.jl:10-13,
    velocity_tracer_gradients.jl:25-46,78"""
    FT = ctx.FT
    u, v, w = (ctx.field(f) for f in U)
    two = FT(2)
    tr = ddC(ctx, u, 0)(O) ** 2 + ddC(ctx, v, 1)(O) ** 2 + ddC(ctx, w, 2)(O) ** 2
    S12, S13, S23 = _strain_offdiag(ctx, U, 0, 1), _strain_offdiag(ctx, U, 0, 2), _strain_offdiag(ctx, U, 1, 2)
    Ixy = iC(ctx, iC(ctx, _sq(S12), 0), 1)
    Ixz = iC(ctx, iC(ctx, _sq(S13), 0), 2)
    Iyz = iC(ctx, iC(ctx, _sq(S23), 1), 2)
    return tr + two * Ixy(O) + two * Ixz(O) + two * Iyz(O)


def dz_b_q(ctx, buoyancy, tracers):
    """∂z_b at ccf: buoyancy_tracer.jl:16 ; seawater_buoyancy.jl:219-224 (LinearEquationOfState: α, β constants) ; no_buoyancy.jl:9"""
    FT = ctx.FT
    if buoyancy is None:
        return lambda o: ctx.zeros()
    if buoyancy.kind == "tracer":
        return ddF(ctx, ctx.field(tracers["b"]), 2)
    dT, dS = ddF(ctx, ctx.field(tracers["T"]), 2), ddF(ctx, ctx.field(tracers["S"]), 2)
    gg, al, be = FT(buoyancy.g), FT(buoyancy.alpha), FT(buoyancy.beta)
    return lambda o: gg * (al * dT(o) - be * dS(o))


def lilly_stability(FT, N2, S2, cb):
    """stability(N², Σ², cᵇ)  lilly_coefficient.jl:114-126"""
    N2p = np.maximum(FT(0), N2)
    with np.errstate(divide="ignore", invalid="ignore"):
        s2 = FT(1) - np.minimum(FT(1), cb * N2p / S2)
        return np.where(S2 == 0, FT(0), np.sqrt(s2)).astype(FT)


def compute_smagorinsky(ctx, closure, U, tracers, buoyancy, nu_e, kappa_e):
    """_compute_smagorinsky_viscosity!  smagorinsky.jl:92-108 ; κₑ = νₑ / Pr  (:44,154-156: the face diffusivities are
    ℑ(νₑ) / Pr — the oracle stores νₑ / Pr per tracer and interpolates that, identical up to one rounding)"""
    g, FT = ctx.g, ctx.FT
    S2 = strain_double_dot_ccc(ctx, U)
    D3 = g.D[0] * g.D[1] * ctx.dz("c")(O)                     # Δxᶜᶜᶜ Δyᶜᶜᶜ Δzᶜᶜᶜ (Flat: 1)
    Df = np.cbrt(np.asarray(D3, dtype=FT))
    if closure.Cb is None:
        cs2 = FT(closure.C) ** 2                             # square_smagorinsky_coefficient(::ConstantSmagorinsky) :110
    else:
        N2 = iC(ctx, dz_b_q(ctx, buoyancy, tracers), 2)(O)   # ℑzᵃᵃᶜ ∂z_b   lilly_coefficient.jl:130
        cs2 = lilly_stability(FT, N2, S2, FT(closure.Cb)) * FT(closure.C) ** 2
    nu = cs2 * Df ** 2 * np.sqrt(FT(2) * S2)
    nu_e.interior[...] = nu
    for name in tracers:
        kappa_e[name].interior[...] = nu / FT(closure.Pr_for(name))


# ---------------------------------------------------------------------------------
# Buoyancy
# ---------------------------------------------------------------------------------
def _validate_unit_vector(v):
    """validate_unit_vector  src/Grids/grid_utils.jl: three components, unit norm"""
    v = tuple(float(x) for x in v)
    if len(v) != 3 or not np.isclose(np.sqrt(sum(x * x for x in v)), 1.0):
        raise ValueError("unit vector must have three components and be unitary")
    return v


class SeawaterBuoyancy:
    """SeawaterBuoyancy with LinearEquationOfState: b = g (α T - β S).  gravity_unit_vector: BuoyancyForce(formulation;
    gravity_unit_vector)  src/BuoyancyFormulations/buoyancy_force.jl:47-58 (None = NegativeZDirection())"""

    def __init__(self, g=9.80665, alpha=1.67e-4, beta=7.8e-4, gravity_unit_vector=None):
        self.g, self.alpha, self.beta = g, alpha, beta
        self.kind = "seawater"
        self.required = ("T", "S")
        self.gravity_unit_vector = None if gravity_unit_vector is None else _validate_unit_vector(gravity_unit_vector)


class BuoyancyTracer:
    def __init__(self, gravity_unit_vector=None):
        self.kind = "tracer"
        self.required = ("b",)
        self.gravity_unit_vector = None if gravity_unit_vector is None else _validate_unit_vector(gravity_unit_vector)


def g_hat(buoyancy, d):
    """ĝ_x, ĝ_y, ĝ_z = −gravity_unit_vector (0, 0, 1 for NegativeZDirection)   buoyancy_force.jl:52-58"""
    guv = getattr(buoyancy, "gravity_unit_vector", None)
    if guv is None:
        return (0.0, 0.0, 1.0)[d]
    return -guv[d]


def buoyancy_q(ctx, buoyancy, tracers):
    FT = ctx.FT
    if buoyancy.kind == "tracer":
        return ctx.field(tracers["b"])
    T, S = ctx.field(tracers["T"]), ctx.field(tracers["S"])
    gg, al, be = FT(buoyancy.g), FT(buoyancy.alpha), FT(buoyancy.beta)
    return lambda o: gg * (al * T(o) - be * S(o))


# ---------------------------------------------------------------------------------
# Coriolis: FPlane with active-weighted interpolation
# ---------------------------------------------------------------------------------
def _inactive_cell(ctx, o):
    g = ctx.g
    bad = np.zeros((1, 1, 1), dtype=bool)
    for d in range(3):
        if g.bounded(d):
            idx = ctx.index(d, o)
            bad = bad | (idx < 1) | (idx > g.N[d])
    return bad


def _peripheral(ctx, o, loc):
    """peripheral_node for loc with exactly one Face (inactive_node.jl:152-158)"""
    res = _inactive_cell(ctx, o)
    for d in range(3):
        if loc[d] == "f":
            res = res | _inactive_cell(ctx, sh(o, d, -1))
    return res


class FPlane:
    """FPlane(f) or FPlane(rotation_rate, latitude): f = 2Ω sind(φ)   src/Coriolis/f_plane.jl:9-44"""

    def __init__(self, f=None, rotation_rate=None, latitude=None):
        use_f = f is not None
        use_planet = latitude is not None
        if use_f == use_planet or (use_f and rotation_rate is not None):
            raise ValueError("Either both keywords rotation_rate and latitude must be specified, *or* only f must be specified.")
        if use_planet:
            rotation_rate = 7.292115e-5 if rotation_rate is None else rotation_rate
            f = 2 * rotation_rate * _sind(latitude)
        self.f = f
        self.kind = "fplane"


class BetaPlane:
    """BetaPlane(f₀, β) or BetaPlane(rotation_rate, latitude, radius): f = f₀ + β y   src/Coriolis/beta_plane.jl:1-47"""

    def __init__(self, f0=None, beta=None, rotation_rate=None, latitude=None, radius=None):
        use_fb = f0 is not None and beta is not None
        use_planet = latitude is not None
        if use_fb == use_planet or ((f0 is None) != (beta is None)):
            raise ValueError("Either both keywords f₀ and β must be specified, *or* all of rotation_rate, latitude, and radius.")
        if use_planet:
            rotation_rate = 7.292115e-5 if rotation_rate is None else rotation_rate
            radius = 6371.0e3 if radius is None else radius
            f0 = 2 * rotation_rate * _sind(latitude)
            beta = 2 * rotation_rate * _cosd(latitude) / radius
        self.f0, self.beta = f0, beta
        self.kind = "betaplane"


class NonTraditionalBetaPlane:
    """NonTraditionalBetaPlane(fz, fy, β, γ, radius | rotation_rate, latitude, radius)   src/Coriolis/non_traditional_beta_plane.jl:16-77"""

    def __init__(self, fz=None, fy=None, beta=None, gamma=None, rotation_rate=None, latitude=None, radius=None):
        fs = (fz, fy, beta, gamma)
        use_f = (not all(c is None for c in fs)) and latitude is None
        use_planet = latitude is not None and all(c is None for c in fs)
        if use_f == use_planet:                                                                        # !xor(...)  :64-67
            raise ValueError("Either the keywords fz, fy, β, γ, and radius must be specified, *or* all of rotation_rate, latitude, and radius.")
        radius = 6371.0e3 if radius is None else radius
        if use_planet:
            om = 7.292115e-5 if rotation_rate is None else rotation_rate
            fz, fy = 2 * om * _sind(latitude), 2 * om * _cosd(latitude)
            beta, gamma = 2 * om * _cosd(latitude) / radius, -4 * om * _sind(latitude) / radius
        self.fz, self.fy, self.beta, self.gamma, self.R = fz, fy, beta, gamma, radius
        self.kind = "ntbetaplane"


class ConstantCartesianCoriolis:
    """ConstantCartesianCoriolis(fx, fy, fz | f, rotation_axis | latitude, rotation_rate)   src/Coriolis/constant_cartesian_coriolis.jl:11-67"""

    def __init__(self, fx=None, fy=None, fz=None, f=None, rotation_axis=None, latitude=None, rotation_rate=None):
        comps = (fx, fy, fz)
        if latitude is not None:
            if any(c is not None for c in comps) or f is not None:
                raise ValueError("Only `rotation_rate` can be specified when using `latitude`.")
            rotation_rate = 7.292115e-5 if rotation_rate is None else rotation_rate
            fx, fy, fz = 0.0, 2 * rotation_rate * _cosd(latitude), 2 * rotation_rate * _sind(latitude)
        elif f is not None:
            if any(c is not None for c in comps):
                raise ValueError("Only `rotation_axis` can be specified when using `f`.")
            if rotation_axis is None:
                fx, fy, fz = 0.0, 0.0, f
            else:
                ax = np.asarray(rotation_axis, dtype=np.float64)
                if ax.shape != (3,) or not np.isclose(np.sqrt((ax ** 2).sum()), 1.0):      # validate_unit_vector
                    raise ValueError("unit vector must be unitary")
                fx, fy, fz = (f * a for a in ax)
        elif all(c is not None for c in comps):
            pass
        else:
            raise ValueError("Either (i) `latitude`, or (ii) `f`, or (iii) `fx`, `fy` and `fz` must be specified.")
        self.fx, self.fy, self.fz = fx, fy, fz
        self.kind = "cartesian"


def _sind(deg):
    """sind / cosd of Julia are exact at multiples of 30° / 45° / 90°"""
    exact = {0: 0.0, 30: 0.5, 90: 1.0, 150: 0.5, 180: 0.0}
    d = deg % 360
    if d in exact:
        return exact[d]
    if d - 180 in exact:
        return -exact[d - 180]
    return float(np.sin(np.deg2rad(deg)))


def _cosd(deg):
    return _sind(deg + 90)


def coriolis_cross(ctx, cor, U, comp):
    """x_f_cross_U / y_f_cross_U / z_f_cross_U at the location of velocity component comp, over the window of ctx.
    FPlane f_plane.jl:50-52 ; BetaPlane beta_plane.jl:56-72 ; ConstantCartesianCoriolis constant_cartesian_coriolis.jl:70-81"""
    FT, g = ctx.FT, ctx.g
    if cor.kind == "fplane":
        return fplane_x(ctx, cor.f, U) if comp == 0 else (fplane_y(ctx, cor.f, U) if comp == 1 else ctx.zeros())
    if cor.kind == "betaplane":
        if comp == 2:
            return ctx.zeros()
        # ynode(i, j, k, grid, Face, Center, Center) for u ; (Center, Face, Center) for v
        off = 0.5 if comp == 0 else 0.0
        y = FT(g.x0[1]) + (ctx.index(1, O).astype(np.float64) - 1 + off) * float(g.D[1]) if not g.flat(1) else np.zeros((1, 1, 1))
        fy = FT(cor.f0) + FT(cor.beta) * y.astype(FT)
        unit = fplane_x(ctx, 1.0, U) if comp == 0 else fplane_y(ctx, 1.0, U)          # ∓ active_weighted_ℑxy(·)
        return fy * unit
    if cor.kind == "ntbetaplane":
        # two_Ωʸ = fy (1 − z/R) + γ y ; two_Ωᶻ = fz (1 + 2z/R) + β y at ynode / znode of the evaluation point   non_traditional_beta_plane.jl:79-96
        u, v, w = (ctx.field(f) for f in U)
        fzz, fyy, be, ga, R = FT(cor.fz), FT(cor.fy), FT(cor.beta), FT(cor.gamma), FT(cor.R)

        def node(d, off):
            if d == 2 and g.stretched:                 # grid.z.cᵃᵃᶜ[k] / grid.z.cᵃᵃᶠ[k]  (grid_generation.jl:33-94)
                tab = g._zC if off else g._zF
                return lambda o: tab[ctx.index(2, o) - 1 + g.H[2]].astype(FT)
            return lambda o: (FT(g.x0[d]) + ((ctx.index(d, o).astype(np.float64) - 1 + off) * float(g.D[d])).astype(FT)).astype(FT)

        def Oy(y, z):
            return lambda o: fyy * (FT(1) - z(o) / R) + ga * y(o)

        def Oz(y, z):
            return lambda o: fzz * (FT(1) + FT(2) * z(o) / R) + be * y(o)

        yc, yf, zc, zf = node(1, 0.5), node(1, 0.0), node(2, 0.5), node(2, 0.0)
        if comp == 0:
            a, b = iC(ctx, w, 2), iC(ctx, v, 1)
            oy, oz = Oy(yc, zc), Oz(yc, zc)
            return iF(ctx, lambda o: oy(o) * a(o) - oz(o) * b(o), 0)(O)
        if comp == 1:
            return Oz(yf, zc)(O) * iF(ctx, iC(ctx, u, 0), 1)(O)                    # ℑxyᶜᶠᵃ = ℑyᶠ(ℑxᶜ u)
        return -Oy(yc, zf)(O) * iF(ctx, iC(ctx, u, 0), 2)(O)                       # ℑxzᶜᵃᶠ = ℑzᶠ(ℑxᶜ u)
    fx, fy, fz = FT(cor.fx), FT(cor.fy), FT(cor.fz)
    u, v, w = (ctx.field(f) for f in U)
    if comp == 0:
        a, b = iC(ctx, w, 2), iC(ctx, v, 1)
        return iF(ctx, lambda o: fy * a(o) - fz * b(o), 0)(O)
    if comp == 1:
        a, b = iC(ctx, u, 0), iC(ctx, w, 2)
        return iF(ctx, lambda o: fz * a(o) - fx * b(o), 1)(O)
    a, b = iC(ctx, v, 1), iC(ctx, u, 0)
    return iF(ctx, lambda o: fx * a(o) - fy * b(o), 2)(O)


def fplane_x(ctx, f, U):
    """x_f_cross_U = -f * active_weighted_ℑxyᶠᶜᶜ(v)   f_plane.jl:50"""
    FT = ctx.FT
    v = ctx.field(U[1])
    num = iC(ctx, iF(ctx, v, 0), 1)(O)                      # ℑxyᶠᶜᵃ = ℑyᶜ(ℑxᶠ(v))
    act = lambda o: (~_peripheral(ctx, o, "cfc")).astype(FT) * np.ones(ctx.shape, FT)
    nodes = iC(ctx, iF(ctx, act, 0), 1)(O)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = np.where(nodes == 0, FT(0), num / nodes)
    return -FT(f) * val


def fplane_y(ctx, f, U):
    """y_f_cross_U = +f * active_weighted_ℑxyᶜᶠᶜ(u)   f_plane.jl:51"""
    FT = ctx.FT
    u = ctx.field(U[0])
    num = iF(ctx, iC(ctx, u, 0), 1)(O)                      # ℑxyᶜᶠᵃ = ℑyᶠ(ℑxᶜ(u))
    act = lambda o: (~_peripheral(ctx, o, "fcc")).astype(FT) * np.ones(ctx.shape, FT)
    nodes = iF(ctx, iC(ctx, act, 0), 1)(O)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = np.where(nodes == 0, FT(0), num / nodes)
    return FT(f) * val
