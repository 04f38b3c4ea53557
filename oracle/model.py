"""Oracle: NonhydrostaticModel state, tendencies, pressure solve and time stepping.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  src/Models/NonhydrostaticModels/nonhydrostatic_model.jl:115-244
  .../update_nonhydrostatic_model_state.jl:20-69, update_hydrostatic_pressure.jl:12-49
  .../compute_nonhydrostatic_tendencies.jl:18-184, nonhydrostatic_tendency_kernel_functions.jl:70-298
  .../solve_for_pressure.jl:12-18,78-95, pressure_correction.jl:8-53, set_nonhydrostatic_model.jl:33-60
  src/Solvers/fft_based_poisson_solver.jl:52-137, poisson_eigenvalues.jl:8-31, plan_transforms.jl:16-34,
      discrete_transforms.jl:26-34 (FFTW REDFT10 / REDFT01 * 1/2N  ==  scipy.fft.dct type 2 / type 3 * 1/2N)
  src/TimeSteppers/runge_kutta_3.jl:60-226, quasi_adams_bashforth_2.jl:74-175, store_tendencies.jl:6-22,
      clock.jl:128-143
  src/BoundaryConditions/compute_flux_bcs.jl:57-163
  src/Utils/kernel_launching.jl:145-195 (exclude_periphery work ranges)
"""
from fractions import Fraction

import numpy as np
import scipy.fft as sfft

from . import advection as adv
from . import closures as clo
from .grid import BC, Field, Grid, fill_halo_regions, SIDES
from .operators import Ctx, O, dF, ddF, div_ccc, iF, sh


def poisson_eigenvalues(N, L, topo):
    """poisson_eigenvalues.jl:8-31 — always Float64 (Int * π promotion)"""
    i = np.arange(1, N + 1, dtype=np.float64)
    L = float(L)
    if topo == "P":
        return (2.0 * np.sin((i - 1) * np.pi / N) / (L / N)) ** 2
    if topo == "B":
        return (2.0 * np.sin((i - 1) * np.pi / (2 * N)) / (L / N)) ** 2
    return np.zeros(N)


class Clock:
    def __init__(self):
        self.time, self.iteration, self.stage = 0.0, 0, 1
        self.last_dt, self.last_stage_dt = float("inf"), float("inf")

    def tick(self, dt, stage=False):
        """tick!  clock.jl:128-143"""
        self.time += dt
        if stage:
            self.stage += 1
            self.last_stage_dt = dt
        else:
            self.iteration += 1
            self.stage = 1
            self.last_dt = dt
            self.last_stage_dt = dt


class OracleModel:
    """NonhydrostaticModel(; grid, advection, closure, tracers, buoyancy, coriolis, timestepper, boundary_conditions)."""

    def __init__(self, grid, advection=None, closure=None, tracers=(), buoyancy=None, coriolis_f=None,
                 timestepper="RungeKutta3", boundary_conditions=None, chi=0.1, coriolis=None):
        FT = grid.FT
        tracers = () if tracers is None else tuple(tracers)
        if coriolis is None and coriolis_f is not None:
            coriolis = clo.FPlane(f=coriolis_f)
        if advection is None:
            advection = adv.Centered(FT, 2)
        # adapt_advection_order, then inflate_grid_halo_size (per direction)  nonhydrostatic_model.jl:175-184,248-262
        advection = adv.adapt_advection_order(advection, grid)
        closures = () if closure is None else (tuple(closure) if isinstance(closure, (tuple, list)) else (closure,))
        cneed = 1
        for c in closures:
            cneed = max(cneed, 2 if c.kind in ("amd", "smagorinsky") else 1)   # AbstractScalarDiffusivity{…, 2}: smagorinsky.jl:31
        H = tuple(max(grid.H[d], adv.scheme_of(advection, d).buffer, cneed) if not grid.flat(d) else 0 for d in range(3))
        if H != grid.H:
            grid = grid.with_halo(H)
        for d in range(3):
            if not grid.flat(d):
                assert grid.N[d] >= grid.H[d], "halo must be <= size (validate_halo)"
                # the d-scheme interpolates velocities along every other direction c with its advecting_velocity_scheme: where the
                # adapted grid's halo H[c] is smaller than that stencil the reference reads outside the halo — not restated
                sd = adv.scheme_of(advection, d)
                deep = sd.buffer if sd.kind == "centered" else (sd.advecting_velocity_scheme.buffer if hasattr(sd, "advecting_velocity_scheme") else 1)
                for c in range(3):
                    assert c == d or grid.flat(c) or grid.H[c] >= deep, "adapted scheme reads outside the halo in the reference"
        self.grid, self.FT = grid, FT
        self.advection, self.closures, self.buoyancy, self.coriolis = advection, closures, buoyancy, coriolis
        bcs = boundary_conditions or {}
        self.u = Field(grid, "fcc", bcs.get("u"), "u")
        self.v = Field(grid, "cfc", bcs.get("v"), "v")
        self.w = Field(grid, "ccf", bcs.get("w"), "w")
        self.U = (self.u, self.v, self.w)
        self.tracers = {n: Field(grid, "ccc", bcs.get(n), n) for n in tracers}
        if buoyancy is not None:
            for n in buoyancy.required:
                assert n in self.tracers, f"buoyancy needs tracer {n}"
        self.pNHS = Field(grid, "ccc", None, "pNHS")
        # pHY′ allocated iff buoyancy !== nothing (nonhydrostatic_model.jl:147-153); gravity is -z
        self.pHY = Field(grid, "ccc", None, "pHY") if buoyancy is not None else None
        self.nu_e = self.kappa_e = None
        for c in closures:
            if c.kind in ("amd", "smagorinsky"):
                assert self.nu_e is None, "one eddy-viscosity closure per model"
                # boundary conditions of the diffusivity fields: (νₑ = …, κₑ = (tracer = …,)) in the reference's model_bcs
                # (test/test_boundary_conditions_integration.jl:62-65; build_diffusivity_fields, anisotropic_minimum_dissipation.jl:364-384)
                self.nu_e = Field(grid, "ccc", bcs.get("nu_e"), "nu_e")
                self.kappa_e = {n: Field(grid, "ccc", (bcs.get("kappa_e") or {}).get(n), "kappa_e_" + n) for n in tracers}
        for c in closures:
            if c.kind == "scalar" and getattr(c, "is_array", False):
                def center_field(v, label):
                    f = Field(grid, "ccc", None, label)
                    f.set(v if isinstance(v, np.ndarray) else float(v))
                    fill_halo_regions(f)
                    return f
                c.nu_field = center_field(c.nu, "nu")
                c.kappa_fields = {n: center_field(c.kappa_for(n), "kappa_" + n) for n in tracers}
        self.fields = {"u": self.u, "v": self.v, "w": self.w, **self.tracers}
        self.Gn = {n: f.like("Gn_" + n) for n, f in self.fields.items()}
        self.Gm = {n: f.like("Gm_" + n) for n, f in self.fields.items()}
        self.timestepper = timestepper
        self.chi = chi
        self.clock = Clock()
        # FFTBasedPoissonSolver eigenvalues
        self.lam = tuple(poisson_eigenvalues(grid.N[d], grid.L[d], grid.topo[d]) for d in range(3))
        # RK3 coefficients converted from rationals to FT (runge_kutta_3.jl:69-78)
        r = lambda a, b: FT(float(Fraction(a, b)))
        self.g1, self.g2, self.g3 = r(8, 15), r(5, 12), r(3, 4)
        self.z2, self.z3 = r(-17, 60), r(-5, 12)
        self.update_state(compute_tendencies=False)

    # ------------------------------------------------------------------ helpers
    def _ctx_for(self, f):
        """work range with exclude_periphery=true (kernel_launching.jl:145-195)"""
        g = self.grid
        rng = []
        for d in range(3):
            lo = 2 if (f.loc[d] == "f" and g.bounded(d) and g.N[d] > 1) else 1
            rng.append((lo, g.N[d]))
        return Ctx(g, *rng)

    def _ctx_full(self):
        g = self.grid
        return Ctx(g, (1, g.Nx), (1, g.Ny), (1, g.Nz))

    @staticmethod
    def _target(f, ctx):
        g = f.grid
        sl = tuple(slice(ctx.r[d][0] + g.H[d] - 1, ctx.r[d][1] + g.H[d]) for d in range(3))
        return f.data[sl]

    # ------------------------------------------------------------------ set!
    def set(self, enforce_incompressibility=True, **kw):
        """set!(model; kwargs...)  set_nonhydrostatic_model.jl:33-60"""
        for name, value in kw.items():
            f = self.fields[name]
            f.set(value)
            fill_halo_regions(f)
        self.update_state(compute_tendencies=False)
        if enforce_incompressibility:
            one = self.FT(1)
            self.compute_pressure_correction(one)
            self.make_pressure_correction(one)
            self.update_state(compute_tendencies=False)

    # ------------------------------------------------------------------ update_state!
    def update_state(self, compute_tendencies=True):
        """update_state!  update_nonhydrostatic_model_state.jl:20-56"""
        for f in self.fields.values():
            fill_halo_regions(f, fill_open_bcs=False)
        self.compute_auxiliaries()
        if self.nu_e is not None:
            fill_halo_regions(self.nu_e)
            for f in self.kappa_e.values():
                fill_halo_regions(f)
        if compute_tendencies:
            self.compute_tendencies()

    def compute_auxiliaries(self):
        """compute_auxiliaries!  :58-69 — diffusivities then hydrostatic pressure"""
        for c in self.closures:
            if c.kind == "amd":
                clo.compute_amd(self._ctx_full(), c, self.U, self.tracers, self.nu_e, self.kappa_e, self.buoyancy)
            elif c.kind == "smagorinsky":
                clo.compute_smagorinsky(self._ctx_full(), c, self.U, self.tracers, self.buoyancy, self.nu_e, self.kappa_e)
        self.update_hydrostatic_pressure()

    def update_hydrostatic_pressure(self):
        """_update_hydrostatic_pressure!  update_hydrostatic_pressure.jl:12-49"""
        g = self.grid
        if self.pHY is None or g.flat(2):
            return
        ir = (1, g.Nx) if g.flat(0) else (0, g.Nx + 1)
        jr = (1, g.Ny) if g.flat(1) else (0, g.Ny + 1)
        Nz = g.Nz
        ctx = Ctx(g, ir, jr, (Nz, Nz))
        b = clo.buoyancy_q(ctx, self.buoyancy, self.tracers)
        bf = iF(ctx, b, 2)                                # z_dot_g_bᶜᶜᶠ = ĝ_z ℑzᵃᵃᶠ(b)   g_dot_b.jl:3
        if self.buoyancy.gravity_unit_vector is not None:
            gz, bf0 = self.FT(clo.g_hat(self.buoyancy, 2)), bf
            bf = lambda o: gz * bf0(o)
        p = ctx.field(self.pHY)
        tgt = lambda kk: self._target(self.pHY, Ctx(g, ir, jr, (kk, kk)))
        dzf = ctx.dz("f")                                 # Δzᶜᶜᶠ(k+1)
        tgt(Nz)[...] = -bf((0, 0, 1)) * dzf((0, 0, 1))
        for k in range(Nz - 1, 0, -1):
            off = k - Nz
            tgt(k)[...] = p((0, 0, off + 1)) - bf((0, 0, off + 1)) * dzf((0, 0, off + 1))

    # ------------------------------------------------------------------ tendencies
    def _closure_nu_kappa(self, ctx, c, name):
        if c.kind == "scalar" and getattr(c, "is_array", False):
            return ctx.field(c.nu_field), (ctx.field(c.kappa_fields[name]) if name is not None else None)
        if c.kind == "scalar":
            return c.nu, (c.kappa_for(name) if name is not None else None)
        return ctx.field(self.nu_e), (ctx.field(self.kappa_e[name]) if name is not None else None)

    def compute_tendencies(self):
        """compute_tendencies! -> compute_Gu!/Gv!/Gw!/Gc!  compute_nonhydrostatic_tendencies.jl:18-163"""
        g, FT = self.grid, self.FT
        sch = self.advection
        for comp, name in enumerate(("u", "v", "w")):
            f = self.fields[name]
            ctx = self._ctx_for(f)
            G = -adv.div_momentum(ctx, sch, self.U, comp)
            # buoyancy: gravity = -ẑ  =>  x̂·g b = ŷ·g b = 0 ; ẑ·g b only when pHY′ is nothing (:168-170,222)
            if comp == 2 and self.buoyancy is not None and self.pHY is None:
                G = G + iF(ctx, clo.buoyancy_q(ctx, self.buoyancy, self.tracers), 2)(O)
            # tilted gravity: x_dot_g_bᶠᶜᶜ = ĝ_x ℑxᶠ b, y_dot_g_bᶜᶠᶜ = ĝ_y ℑyᶠ b   g_dot_b.jl:1-2 (:95,157)
            if comp < 2 and self.buoyancy is not None and self.buoyancy.gravity_unit_vector is not None:
                G = G + FT(clo.g_hat(self.buoyancy, comp)) * iF(ctx, clo.buoyancy_q(ctx, self.buoyancy, self.tracers), comp)(O)
            if self.coriolis is not None:
                if comp < 2 or self.coriolis.kind in ("cartesian", "ntbetaplane"):       # z_f_cross_U = 0 for FPlane / BetaPlane
                    G = G - clo.coriolis_cross(ctx, self.coriolis, self.U, comp)
            if self.pHY is not None and comp < 2:
                G = G - ddF(ctx, ctx.field(self.pHY), comp)(O)      # hydrostatic_pressure_gradient_x/y
            for c in self.closures:
                nu, _ = self._closure_nu_kappa(ctx, c, None)
                G = G - clo.div_tau(ctx, nu, self.U, comp)
            self._target(self.Gn[name], ctx)[...] = G
        for name, cf in self.tracers.items():
            ctx = self._ctx_for(cf)
            G = -adv.div_tracer(ctx, sch, self.U, cf)
            for c in self.closures:
                _, kap = self._closure_nu_kappa(ctx, c, name)
                G = G - clo.div_q(ctx, kap, cf)
            self._target(self.Gn[name], ctx)[...] = G

    def compute_flux_bc_tendencies(self):
        """compute_flux_bc_tendencies!  compute_nonhydrostatic_tendencies.jl:170-184 ; compute_flux_bcs.jl:116-163"""
        g, FT = self.grid, self.FT
        for name, f in self.fields.items():
            G = self.Gn[name]
            for d in range(3):
                for side in (0, 1):
                    bc = f.bcs[SIDES[2 * d + side]]
                    if bc.kind != "flux" or bc.value is None:
                        continue
                    idx = 1 if side == 0 else g.N[d]
                    rng = [(1, g.N[e]) for e in range(3)]
                    rng[d] = (idx, idx)
                    cx = Ctx(g, *rng)
                    tgt = self._target(G, cx)
                    val = bc.get(FT)
                    if not np.isscalar(val):
                        val = np.expand_dims(val, d)
                    # A(…, flip(L_d)) / volume(…, LX, LY, LZ)   compute_flux_bcs.jl:116-161
                    zl = f.loc[2]
                    contrib = val * cx.area(d, zl)(O) / cx.vol(zl)(O)
                    if side == 0:
                        tgt[...] = tgt + contrib
                    else:
                        tgt[...] = tgt - contrib

    # ------------------------------------------------------------------ substeps
    def rk3_substep(self, dt, gamma, zeta):
        """rk3_substep! / rk3_substep_field!  runge_kutta_3.jl:179-226"""
        FT = self.FT
        dt = FT(dt)
        for name, f in self.fields.items():
            ctx = self._ctx_for(f)
            U = self._target(f, ctx)
            Gn = self._target(self.Gn[name], ctx)
            if zeta is None:
                U[...] = U + dt * gamma * Gn
            else:
                Gm = self._target(self.Gm[name], ctx)
                U[...] = U + dt * (gamma * Gn + zeta * Gm)

    def ab2_step(self, dt, chi):
        """ab2_step! / ab2_step_field!  quasi_adams_bashforth_2.jl:127-175"""
        FT = self.FT
        dt, chi = FT(dt), FT(chi)
        not_euler = FT(1) if chi != FT(-0.5) else FT(0)
        for name, f in self.fields.items():
            ctx = self._ctx_for(f)
            U = self._target(f, ctx)
            Gn = self._target(self.Gn[name], ctx)
            Gm = self._target(self.Gm[name], ctx)
            with np.errstate(invalid="ignore"):
                Gu = (FT(1.5) + chi) * Gn - (FT(0.5) + chi) * Gm * not_euler
            U[...] = U + dt * Gu

    def cache_previous_tendencies(self):
        """cache_previous_tendencies!  store_tendencies.jl:6-22 (whole :xyz interior of the grid)"""
        ctx = self._ctx_full()
        for name in self.fields:
            self._target(self.Gm[name], ctx)[...] = self._target(self.Gn[name], ctx)

    # ------------------------------------------------------------------ pressure
    def solve_poisson(self, rhs):
        """solve!(ϕ, ::FFTBasedPoissonSolver)  fft_based_poisson_solver.jl:95-125.
        rhs: (Nx,Ny,Nz) real array in FT.  Bounded dims: DCT-II forward, DCT-III * 1/(2N) backward;
        Periodic: DFT / inverse DFT; Flat: nothing.  λ in Float64."""
        g, FT = self.grid, self.FT
        CT = np.complex128 if FT is np.float64 else np.complex64
        b = rhs.astype(CT)
        bdims = [d for d in range(3) if g.bounded(d)]
        pdims = [d for d in range(3) if g.topo[d] == "P"]
        for d in bdims:          # plan_r2r!(REDFT10) acts on real and imaginary parts separately
            b = (sfft.dct(b.real, type=2, axis=d) + 1j * sfft.dct(b.imag, type=2, axis=d)).astype(CT)
        if pdims:
            b = sfft.fftn(b, axes=pdims).astype(CT)
        lx, ly, lz = self.lam
        lam = lx[:, None, None] + ly[None, :, None] + lz[None, None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            phi = (-b / lam).astype(CT)
        phi[0, 0, 0] = 0
        if pdims:
            phi = sfft.ifftn(phi, axes=pdims).astype(CT)
        for d in bdims:
            phi = (sfft.dct(phi.real, type=3, axis=d) + 1j * sfft.dct(phi.imag, type=3, axis=d)).astype(CT)
            phi = (phi * (1.0 / (2 * g.N[d]))).astype(CT)
        return phi.real.astype(FT)

    def _tridiagonal_setup(self):
        """FourierTridiagonalPoissonSolver(grid) for a z-stretched grid  fourier_tridiagonal_poisson_solver.jl:74-131,
        compute_main_diagonal! (HomogeneousZFormulation) :172-185, compute_lower_diagonal! :198-202"""
        g, FT = self.grid, self.FT
        Nz = g.Nz
        k = np.arange(1, Nz + 1)
        dzc = g.dz_at("c", k)
        one = FT(1)
        lam = (self.lam[0][:, None] + self.lam[1][None, :])[:, :, None]          # Float64
        D = np.empty((g.Nx, g.Ny, Nz), dtype=FT)
        D[:, :, 0] = (-one / g.dz_at("f", 2) - dzc[0] * lam[:, :, 0]).astype(FT)
        D[:, :, Nz - 1] = (-one / g.dz_at("f", Nz) - dzc[Nz - 1] * lam[:, :, 0]).astype(FT)
        for kk in range(2, Nz):
            D[:, :, kk - 1] = (-(one / g.dz_at("f", kk + 1) + one / g.dz_at("f", kk)) - dzc[kk - 1] * lam[:, :, 0]).astype(FT)
        self._tri_D = D
        self._tri_lower = np.asarray([one / g.dz_at("f", q + 1) for q in range(1, Nz)], dtype=FT)    # = upper diagonal
        CT = np.complex128 if FT is np.float64 else np.complex64
        self._tri_phi = np.zeros((g.Nx, g.Ny, Nz), dtype=CT)                     # solver.storage (persists between solves)
        self._tri_t = np.zeros((g.Nx, g.Ny, Nz), dtype=FT)                       # scratch

    def solve_poisson_tridiagonal(self, rhs):
        """solve!(x, ::FourierTridiagonalPoissonSolver)  fourier_tridiagonal_poisson_solver.jl:204-231 with
        solve_batched_tridiagonal_system_z!  batched_tridiagonal_solver.jl:220-243.
        rhs: Δzᶜᶜᶜ·div (Nx,Ny,Nz) real FT (solve_for_pressure.jl:36-42,69-76)."""
        g, FT = self.grid, self.FT
        if not hasattr(self, "_tri_D"):
            self._tridiagonal_setup()
        CT = self._tri_phi.dtype.type
        f = rhs.astype(CT)
        bdims = [d for d in (0, 1) if g.bounded(d)]
        pdims = [d for d in (0, 1) if g.topo[d] == "P"]
        for d in bdims:
            f = (sfft.dct(f.real, type=2, axis=d) + 1j * sfft.dct(f.imag, type=2, axis=d)).astype(CT)
        if pdims:
            f = sfft.fftn(f, axes=pdims).astype(CT)
        a, b, c, t, phi = self._tri_lower, self._tri_D, self._tri_lower, self._tri_t, self._tri_phi
        Nz = g.Nz
        eps10 = 10 * np.finfo(FT).eps
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            beta = b[:, :, 0].copy()
            phi[:, :, 0] = f[:, :, 0] / beta
            for k in range(1, Nz):
                t[:, :, k] = c[k - 1] / beta
                beta = b[:, :, k] - a[k - 1] * t[:, :, k]
                dominant = np.abs(beta) > eps10
                star = (f[:, :, k] - a[k - 1] * phi[:, :, k - 1]) / beta
                phi[:, :, k] = np.where(dominant, star, phi[:, :, k])
            for k in range(Nz - 2, -1, -1):
                phi[:, :, k] = phi[:, :, k] - t[:, :, k + 1] * phi[:, :, k + 1]
        out = phi
        if pdims:
            out = sfft.ifftn(out, axes=pdims).astype(CT)
        for d in bdims:
            out = (sfft.dct(out.real, type=3, axis=d) + 1j * sfft.dct(out.imag, type=3, axis=d)).astype(CT)
            out = (out * (1.0 / (2 * g.N[d]))).astype(CT)
        out = (out - out.mean()).astype(CT)                      # ϕ .= ϕ .- mean(ϕ)
        self._tri_phi[...] = out                                 # transforms and the mean removal act in place on storage
        return out.real.astype(FT)

    def compute_pressure_correction(self, dt):
        """compute_pressure_correction!  pressure_correction.jl:8-20"""
        for f in self.U:
            fill_halo_regions(f)                                   # open BCs filled here
        ctx = self._ctx_full()
        rhs = div_ccc(ctx, self.u, self.v, self.w)                 # NOT divided by Δt (solve_for_pressure.jl:12-18)
        if self.grid.stretched:
            # _fourier_tridiagonal_source_term!: Δzᶜᶜᶜ · div   (solve_for_pressure.jl:36-42)
            self.pNHS.interior[...] = self.solve_poisson_tridiagonal(ctx.dz("c")(O) * rhs)
        else:
            self.pNHS.interior[...] = self.solve_poisson(rhs)
        fill_halo_regions(self.pNHS)

    def make_pressure_correction(self, dt):
        """make_pressure_correction!  pressure_correction.jl:31-53 (incl. the stale-halo quirk: only the
        interior of pNHS is divided by Δt)."""
        FT = self.FT
        ctx = self._ctx_full()
        p = ctx.field(self.pNHS)
        for d, f in enumerate(self.U):
            tgt = self._target(f, ctx)
            tgt[...] = tgt - ddF(ctx, p, d)(O)
        dtp = max(float(np.finfo(FT).eps), float(dt))            # Float64 after promotion
        self.pNHS.interior[...] = (self.pNHS.interior.astype(np.float64) / dtp).astype(FT)

    # ------------------------------------------------------------------ time stepping
    def time_step(self, dt, euler=False):
        if self.timestepper == "RungeKutta3":
            self._time_step_rk3(dt)
        else:
            self._time_step_ab2(dt, euler)

    def _time_step_rk3(self, dt):
        """time_step!(model::AbstractModel{<:RungeKutta3TimeStepper}, Δt)  runge_kutta_3.jl:93-170"""
        dt = float(dt)
        if self.clock.iteration == 0:
            self.update_state(compute_tendencies=True)
        # stage_Δt(Δt, γ, ζ) = Δt * (γ + ζ) with γ, ζ of type FT  (:176-177)
        dt1, dt2, dt3 = dt * float(self.g1), dt * float(self.g2 + self.z2), dt * float(self.g3 + self.z3)
        tn1 = self.clock.time + dt
        for stage, (gam, zet, sdt) in enumerate(((self.g1, None, dt1), (self.g2, self.z2, dt2), (self.g3, self.z3, dt3))):
            self.compute_flux_bc_tendencies()
            self.rk3_substep(dt, gam, zet)
            if stage < 2:
                self.clock.tick(sdt, stage=True)
            else:
                corrected = tn1 - self.clock.time
                self.clock.tick(sdt)
                self.clock.last_stage_dt = corrected
                self.clock.last_dt = dt
            self.compute_pressure_correction(sdt)
            self.make_pressure_correction(sdt)
            if stage < 2:
                self.cache_previous_tendencies()
            self.update_state(compute_tendencies=True)

    def _time_step_ab2(self, dt, euler=False):
        """time_step!(model::AbstractModel{<:QuasiAdamsBashforth2TimeStepper}, Δt)  quasi_adams_bashforth_2.jl:74-120"""
        dt = float(dt)
        if self.clock.iteration == 0:
            self.update_state(compute_tendencies=True)
        euler = euler or (dt != self.clock.last_dt)
        chi = -0.5 if euler else self.chi
        self.compute_flux_bc_tendencies()
        self.ab2_step(dt, chi)
        self.clock.tick(dt)
        self.compute_pressure_correction(dt)
        self.make_pressure_correction(dt)
        self.cache_previous_tendencies()
        self.update_state(compute_tendencies=True)

    # ------------------------------------------------------------------ diagnostics
    def divergence(self):
        return div_ccc(self._ctx_full(), self.u, self.v, self.w)
