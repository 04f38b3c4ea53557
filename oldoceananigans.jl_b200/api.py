"""Host-side mirror of the reference API for the NonhydrostaticModel path.

Julia is not available in the build image, so the host layer above the C ABI is Python; it mirrors the
reference's names, argument meaning and error behaviour (``!`` becomes a trailing underscore):

    RectilinearGrid(arch, FT; size, extent | x,y,z, topology, halo)   src/Grids/rectilinear_grid.jl:264-291
    NonhydrostaticModel(; grid, advection, closure, tracers, buoyancy, coriolis, timestepper,
                        boundary_conditions)                          src/Models/NonhydrostaticModels/nonhydrostatic_model.jl:115-244
    set_(model; u=…, T=…)            set!         set_nonhydrostatic_model.jl:33-60
    time_step_(model, Δt)            time_step!   src/TimeSteppers/runge_kutta_3.jl:93, quasi_adams_bashforth_2.jl:74
    Simulation(model; Δt, stop_iteration, stop_time), run_(sim)      src/Simulations/run.jl:92-176
    update_state_, compute_tendencies_, fill_halo_regions_, …        staged entry points

The Julia glue with the identical mapping is in julia/OceananigansB200Ext.jl and INTEGRATION.md.
All numerics run in liboceananigans_b200.so (CUDA, sm_100a); nothing here computes on the CPU.
"""
import ctypes as C
import math
from fractions import Fraction

import numpy as np

from . import _lib as L

__all__ = [
    "B200", "Distributed", "Partition", "Periodic", "Bounded", "Flat", "Center", "Face", "RectilinearGrid",
    "Centered", "FluxFormAdvection", "adapt_advection_order", "required_halo_size", "UpwindBiased", "WENO", "ScalarDiffusivity", "AnisotropicMinimumDissipation", "Smagorinsky", "SmagorinskyLilly", "LillyCoefficient",
    "SeawaterBuoyancy", "LinearEquationOfState", "BuoyancyTracer", "BuoyancyForce", "FPlane", "BetaPlane", "ConstantCartesianCoriolis", "NonTraditionalBetaPlane",
    "FluxBoundaryCondition", "ValueBoundaryCondition", "GradientBoundaryCondition", "OpenBoundaryCondition",
    "FieldBoundaryConditions", "NonhydrostaticModel", "Field", "OutputTicket", "set_", "time_step_", "update_state_",
    "compute_tendencies_", "compute_flux_bc_tendencies_", "rk3_substep_", "ab2_step_", "cache_previous_tendencies_",
    "compute_pressure_correction_", "make_pressure_correction_", "fill_halo_regions_", "solve_poisson",
    "interior", "parent", "Simulation", "run_", "Clock", "Checkpointer", "OceananigansB200Error",
    "cell_advection_timescale", "cell_diffusion_timescale", "hasnan", "step_diagnostics", "TimeStepWizard",
]

OceananigansB200Error = L.OceananigansB200Error


class B200:
    """Architecture tag: `struct B200 <: AbstractSerialArchitecture` (cf. GPU(), src/Architectures.jl:44-46)."""

    def __init__(self, device=0):
        self.device = device


class Partition:
    """Partition(x, y, z): ranks per dimension (src/DistributedComputations/distributed_architectures.jl:14-18).
    The B200 path implements slabs, Partition(1, R) and Partition(R, 1), and pencils, Partition(Rx, Ry), for Periodic and Bounded
    x, y and z; z is never partitioned (the reference's solver cannot either, distributed_fft_based_poisson_solver.jl:211-229)."""

    def __init__(self, x=1, y=1, z=1):
        if z != 1:
            raise NotImplementedError("z is never partitioned (distributed_fft_based_poisson_solver.jl:211-229)")
        if int(x) < 1 or int(y) < 1:
            raise ValueError("ranks per dimension must be positive")
        self.x, self.y, self.z = int(x), int(y), int(z)


class Distributed:
    """Distributed(B200(device); partition=Partition(1, R)): one process per GPU
    (src/DistributedComputations/distributed_architectures.jl:242-302).  rank / nranks default to torch.distributed's.
    Without `partition` the ranks form slabs in y, Partition(1, R) — the decomposition this library is tuned for (peer-memory transposes,
    overlapped exchanges); the reference's default is Partition(R), slabs in x (:257-261), which is available by asking for it.
    `exchange` is the TEST-ONLY host transport of the host simulation (a Python callable, see tests/dist_worker.py); the
    CUDA library communicates with NCCL over NVLink."""

    def __init__(self, child_architecture=None, partition=None, rank=None, nranks=None, exchange=None):
        self.child = child_architecture or B200()
        self.device = self.child.device
        if rank is None or nranks is None:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise ValueError("Distributed needs rank/nranks or an initialised torch.distributed process group")
            rank, nranks = dist.get_rank(), dist.get_world_size()
        self.rank, self.nranks = int(rank), int(nranks)
        self.partition = partition or Partition(1, self.nranks)
        if self.partition.x * self.partition.y != self.nranks:
            raise ValueError("partition does not match the number of ranks")
        # rank -> (index along x, index along y): x-major, rank2index (distributed_architectures.jl:354-362)
        self.rx, self.ry = self.rank // self.partition.y, self.rank % self.partition.y
        self.exchange = exchange

    def local_range(self, d, n_global):
        """first index and length of this rank's share of a dimension of n_global cells"""
        if d == 0:
            n = n_global // self.partition.x
            return self.rx * n, n
        if d == 1:
            n = n_global // self.partition.y
            return self.ry * n, n
        return 0, n_global


class _Topo:
    def __init__(self, name, code):
        self.name, self.code = name, code

    def __repr__(self):
        return self.name


Periodic, Bounded, Flat = _Topo("Periodic", L.OC_PERIODIC), _Topo("Bounded", L.OC_BOUNDED), _Topo("Flat", L.OC_FLAT)
Center, Face = "Center", "Face"


def _tuple(x, n=None):
    if x is None:
        return None
    t = tuple(x) if isinstance(x, (tuple, list)) else (x,)
    return t


class RectilinearGrid:
    """RectilinearGrid, regular or vertically stretched.  `size`, `extent`/`x,y,z`, `halo` list only the non-Flat dimensions.
    `z` may be a 2-tuple (interval, regular spacing), or a vector of Nz+1 increasing faces / a function of the face index
    (variably spaced Bounded z: rectilinear_grid.jl:264-291, grid_generation.jl:33-94), which selects the
    FourierTridiagonalPoissonSolver.  Stretched x or y are out of scope."""

    def __init__(self, architecture=None, FT=np.float64, *, size, extent=None, x=None, y=None, z=None,
                 topology=(Periodic, Periodic, Bounded), halo=None):
        if isinstance(architecture, type) or (architecture is not None and not isinstance(architecture, (B200, Distributed))):
            # RectilinearGrid(FT; ...) form
            FT, architecture = architecture, None
        self.architecture = architecture or B200()
        self.FT = np.dtype(FT).type
        if self.FT not in (np.float32, np.float64):
            raise ValueError("eltype(grid) must be Float32 or Float64")
        self.topology = tuple(topology)
        nonflat = [d for d in range(3) if self.topology[d] is not Flat]
        size = _tuple(size)
        if len(size) != len(nonflat):
            raise ValueError(f"size={size} must have {len(nonflat)} elements for topology {self.topology}")  # input_validation.jl
        bounds = [None] * 3
        if extent is not None:
            extent = _tuple(extent)
            if len(extent) != len(nonflat):
                raise ValueError("extent must list the non-Flat dimensions")
            for n, d in enumerate(nonflat):
                bounds[d] = (0.0, float(extent[n])) if d < 2 else (-float(extent[n]), 0.0)
        else:
            for d, b in enumerate((x, y, z)):
                if b is not None:
                    if callable(b) or not isinstance(b, tuple) or len(b) != 2:
                        if d < 2:
                            raise NotImplementedError("stretched x / y coordinates are out of scope (only z can be variably spaced)")
                        continue                                     # stretched z: handled below
                    bounds[d] = (float(b[0]), float(b[1]))
        self.z_faces = None
        if extent is None and z is not None and (callable(z) or not isinstance(z, tuple) or len(z) != 2):
            if self.topology[2] is not Bounded:
                raise ValueError("a variably spaced z needs the Bounded topology")      # fourier_tridiagonal_poisson_solver.jl:86-90
            nz = int(size[nonflat.index(2)])
            zf = [z(i) for i in range(1, nz + 2)] if callable(z) else list(z)
            if len(zf) != nz + 1:
                raise ValueError("z must list Nz+1 faces")
            zf = np.asarray(zf, dtype=self.FT)                      # interior_face_nodes = zeros(FT, N+1)
            if not np.all(np.diff(zf) > 0):
                raise ValueError("The elements of z must be increasing!")              # grid_generation.jl:45-48
            self.z_faces = np.ascontiguousarray(zf.astype(np.float64))
            bounds[2] = (float(zf[0]), float(zf[-1]))
        halo = _tuple(halo)
        N, H, Lx, D, x0 = [1, 1, 1], [0, 0, 0], [1.0] * 3, [1.0] * 3, [0.0] * 3
        for n, d in enumerate(nonflat):
            if bounds[d] is None:
                raise ValueError("missing domain extent for a non-Flat dimension")
            N[d] = int(size[n])
            H[d] = int(halo[n]) if halo is not None else min(3, N[d])            # input_validation.jl:71-77
            a, b = bounds[d]
            if not b > a:
                raise ValueError("domain end must be larger than its start")
            delta = (Fraction(b) - Fraction(a)) / N[d]                          # grid_generation.jl:105-110
            D[d] = float(self.FT(float(delta)))
            Lx[d] = float(self.FT(float(Fraction(b) - Fraction(a))))
            x0[d] = a
        self.N, self.H, self.L, self.D, self.x0 = tuple(N), tuple(H), tuple(Lx), tuple(D), tuple(x0)
        self.Nx, self.Ny, self.Nz = self.N
        self.Hx, self.Hy, self.Hz = self.H
        self.Lx, self.Ly, self.Lz = self.L
        self.dx, self.dy, self.dz = self.D
        if self.z_faces is not None:
            self.L = (self.L[0], self.L[1], float(self.FT(self.FT(self.z_faces[-1]) - self.FT(self.z_faces[0]))))
            self.Lz = self.L[2]
            self.D = (self.D[0], self.D[1], float("nan"))          # no constant Δz
            self.dz = None

    def with_halo(self, halo):
        g = RectilinearGrid.__new__(RectilinearGrid)
        g.__dict__.update(self.__dict__)
        g.H = tuple(0 if self.topology[d] is Flat else int(halo[d]) for d in range(3))
        g.Hx, g.Hy, g.Hz = g.H
        return g

    def nodes(self, d, loc):
        n = self.N[d] + (1 if (loc == Face and self.topology[d] is Bounded) else 0)
        if d == 2 and self.z_faces is not None:
            zf = self.z_faces.astype(self.FT)
            return (zf if loc == Face else ((zf[1:] + zf[:-1]) / self.FT(2))).astype(np.float64)[:n]
        return self.x0[d] + (np.arange(n) + (0.5 if loc == Center else 0.0)) * self.D[d]

    def __repr__(self):
        return f"RectilinearGrid{{{self.FT.__name__}}}(size={self.N}, halo={self.H}, topology={self.topology})"


# ------------------------------------------------------------------------------------------ physics descriptors
class Centered:
    """Centered(order = 2 | 4)   src/Advection/centered_reconstruction.jl:5-55"""

    def __init__(self, FT=np.float64, order=2):
        if order not in (2, 4):
            raise NotImplementedError("Centered: orders 2 and 4 are implemented on B200 (higher orders: SURVEY §8f)")
        self.order, self.buffer = order, order // 2
        self.code = L.OC_CENTERED2 if order == 2 else L.OC_CENTERED4


class UpwindBiased:
    """UpwindBiased(order = 1 | 3 | 5)   src/Advection/upwind_biased_reconstruction.jl:57-86"""

    def __init__(self, FT=np.float64, order=3):
        if order % 2 == 0:
            raise ValueError("UpwindBiased reconstruction scheme is defined only for odd orders")     # :59
        if order not in (1, 3, 5):
            raise NotImplementedError("UpwindBiased: orders 1, 3 and 5 are implemented on B200")
        self.order, self.buffer = order, (order + 1) // 2
        self.code = {1: L.OC_UPWIND1, 3: L.OC_UPWIND3, 5: L.OC_UPWIND5}[order]


class WENO:
    """WENO(order = 3 | 5)   src/Advection/weno_reconstruction.jl:7-93"""

    def __init__(self, FT=np.float64, order=5, bounds=None):
        if order % 2 == 0:
            raise ValueError("WENO reconstruction scheme is defined only for odd orders")         # weno_reconstruction.jl:81
        if order not in (3, 5, 7, 9) or bounds is not None:
            raise NotImplementedError("WENO: orders 3, 5, 7 and 9 without bounds are implemented on B200")
        self.order, self.buffer = order, (order + 1) // 2
        self.code = {3: L.OC_WENO3, 5: L.OC_WENO5, 7: L.OC_WENO7, 9: L.OC_WENO9}[order]


class _NoAdvection:
    """advection = nothing: no advective fluxes (momentum_advection_operators.jl:86-95)"""
    order, buffer, code = 0, 1, L.OC_ADVECTION_NONE


class FluxFormAdvection:
    """FluxFormAdvection(x, y, z): one scheme per flux direction (src/Advection/flux_form_advection.jl) — what adapt_advection_order
    returns when it had to lower the scheme in some direction."""

    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z
        self.dirs = (x, y, z)
        self.buffer = max(s.buffer for s in self.dirs)
        self.code = max(self.dirs, key=lambda s: s.buffer).code


def required_halo_size(advection, d):
    """required_halo_size_x / _y / _z(advection)"""
    return advection.dirs[d].buffer if isinstance(advection, FluxFormAdvection) else advection.buffer


def adapt_advection_order(advection, grid):
    """adapt_advection_order(advection, grid)  (src/Advection/adapt_advection_order.jl:18-96): where the grid has fewer points than the
    scheme's buffer the scheme is lowered in that direction — Centered(order = 2N), UpwindBiased(order = 2N − 1), WENO(order = 2N − 1)
    (WENO(order = 1) is UpwindBiased(order = 1): weno_reconstruction.jl:83-85); Flat directions are left alone.  Returns the scheme
    itself when nothing changed, a FluxFormAdvection otherwise."""
    if isinstance(advection, _NoAdvection):
        return advection
    dirs = list(advection.dirs) if isinstance(advection, FluxFormAdvection) else [advection] * 3
    changed = False
    for d in range(3):
        sch, N = dirs[d], grid.N[d]
        if grid.topology[d] is Flat or isinstance(sch, _NoAdvection) or N >= sch.buffer:
            continue
        if isinstance(sch, Centered):
            new = Centered(order=2 * N)
        elif isinstance(sch, UpwindBiased):
            new = UpwindBiased(order=2 * N - 1)
        else:
            new = WENO(order=2 * N - 1) if 2 * N - 1 >= 3 else UpwindBiased(order=1)
        dirs[d], changed = new, True
    return FluxFormAdvection(*dirs) if changed else advection


class ScalarDiffusivity:
    """ScalarDiffusivity(ν, κ): numbers — or arrays of the grid's size, located at (Center, Center, Center), for ν and / or for κ (one array,
    or a dict tracer -> array | number): the reference interpolates array coefficients to the flux points exactly like eddy viscosities
    (abstract_scalar_diffusivity_closure.jl:323-332).  Halos follow the default boundary conditions of a Center field (periodic / no flux)."""

    def __init__(self, FT=np.float64, nu=0.0, kappa=0.0, **kw):
        self.nu = kw.get("ν", nu)
        self.kappa = kw.get("κ", kappa)
        vals = [self.nu] + (list(self.kappa.values()) if isinstance(self.kappa, dict) else [self.kappa])
        if any(callable(v) for v in vals):
            raise NotImplementedError("function-valued diffusivities are out of scope")
        self.is_array = any(isinstance(v, np.ndarray) for v in vals)


class AnisotropicMinimumDissipation:
    def __init__(self, FT=np.float64, C=1.0 / 3.0, Cnu=None, Ckappa=None, Cb=None):
        # Cb: buoyancy modification multiplier; None turns the term off (anisotropic_minimum_dissipation.jl:35,62-68; the reference
        # warns that the modification is unvalidated, :137)
        self.Cnu = C if Cnu is None else Cnu
        self.Ckappa = C if Ckappa is None else Ckappa
        self.Cb = Cb


class LillyCoefficient:
    """LillyCoefficient(smagorinsky=0.16, reduction_factor=1)  Smagorinskys/lilly_coefficient.jl:47-48"""

    def __init__(self, FT=np.float64, smagorinsky=0.16, reduction_factor=1.0):
        self.smagorinsky, self.reduction_factor = smagorinsky, reduction_factor


class Smagorinsky:
    """Smagorinsky(coefficient=0.16, Pr=1.0)  Smagorinskys/smagorinsky.jl:76-80; `coefficient` a number or a LillyCoefficient
    (DynamicCoefficient is out of scope); `Pr` a number or a dict tracer -> number (tracer_diffusivities)."""

    def __init__(self, FT=np.float64, coefficient=0.16, Pr=1.0):
        if not isinstance(coefficient, (int, float, LillyCoefficient)):
            raise NotImplementedError("Smagorinsky: coefficient must be a number or a LillyCoefficient (DynamicCoefficient is out of scope)")
        self.coefficient, self.Pr = coefficient, Pr


def SmagorinskyLilly(FT=np.float64, C=0.16, Cb=1.0, Pr=1.0):
    """SmagorinskyLilly(C=0.16, Cb=1, Pr=1)  Smagorinskys/lilly_coefficient.jl:107-112"""
    return Smagorinsky(FT, coefficient=LillyCoefficient(FT, smagorinsky=C, reduction_factor=Cb), Pr=Pr)


class LinearEquationOfState:
    def __init__(self, FT=np.float64, thermal_expansion=1.67e-4, haline_contraction=7.8e-4):
        self.thermal_expansion, self.haline_contraction = thermal_expansion, haline_contraction


class SeawaterBuoyancy:
    def __init__(self, FT=np.float64, gravitational_acceleration=9.80665, equation_of_state=None,
                 constant_temperature=None, constant_salinity=None):
        if constant_temperature is not None or constant_salinity is not None:
            raise NotImplementedError("constant_temperature / constant_salinity are out of scope")
        self.g = gravitational_acceleration
        self.eos = equation_of_state or LinearEquationOfState()
        if not isinstance(self.eos, LinearEquationOfState):
            raise NotImplementedError("only LinearEquationOfState (TEOS-10 coefficients are not in the reference tree)")
        self.required = ("T", "S")


class BuoyancyTracer:
    required = ("b",)


class BuoyancyForce:
    """BuoyancyForce(formulation; gravity_unit_vector=NegativeZDirection())   src/BuoyancyFormulations/buoyancy_force.jl:47-58"""

    def __init__(self, formulation, gravity_unit_vector=None):
        if not isinstance(formulation, (SeawaterBuoyancy, BuoyancyTracer)):
            raise NotImplementedError("BuoyancyForce: SeawaterBuoyancy (LinearEquationOfState) or BuoyancyTracer")
        self.formulation = formulation
        self.gravity_unit_vector = None
        if gravity_unit_vector is not None:
            v = tuple(float(x) for x in gravity_unit_vector)
            if len(v) != 3 or not np.isclose(math.sqrt(sum(x * x for x in v)), 1.0):           # validate_unit_vector
                raise ValueError("gravity_unit_vector must have three components and be unitary")
            self.gravity_unit_vector = v
        self.required = formulation.required


_OMEGA_EARTH = 7.292115e-5      # Oceananigans.defaults.planet_rotation_rate
_R_EARTH = 6371.0e3             # Oceananigans.defaults.planet_radius


def _sind(deg):
    """Julia's sind / cosd are exact at multiples of 30° / 90°"""
    exact = {0: 0.0, 30: 0.5, 90: 1.0, 150: 0.5, 180: 0.0}
    d = deg % 360
    if d in exact:
        return exact[d]
    if d - 180 in exact:
        return -exact[d - 180]
    return math.sin(math.radians(deg))


def _cosd(deg):
    return _sind(deg + 90)


class FPlane:
    """FPlane(f=…) or FPlane(rotation_rate=Ω_Earth, latitude=φ): f = 2Ω sind(φ)   src/Coriolis/f_plane.jl:9-44"""

    def __init__(self, FT=np.float64, f=None, rotation_rate=None, latitude=None):
        use_f, use_planet = f is not None, latitude is not None
        if use_f == use_planet or (use_f and rotation_rate is not None):
            raise ValueError("Either both keywords rotation_rate and latitude must be specified, *or* only f must be specified.")
        if use_planet:
            f = 2 * (_OMEGA_EARTH if rotation_rate is None else rotation_rate) * _sind(latitude)
        self.f = f


class BetaPlane:
    """BetaPlane(f₀=…, β=…) or BetaPlane(rotation_rate, latitude, radius): f = f₀ + β y   src/Coriolis/beta_plane.jl:1-47"""

    def __init__(self, FT=np.float64, f0=None, beta=None, rotation_rate=None, latitude=None, radius=None, **kw):
        f0, beta = kw.pop("f₀", f0), kw.pop("β", beta)
        if kw:
            raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
        use_fb, use_planet = f0 is not None and beta is not None, latitude is not None
        if use_fb == use_planet or ((f0 is None) != (beta is None)):
            raise ValueError("Either both keywords f₀ and β must be specified, *or* all of rotation_rate, latitude, and radius.")
        if use_planet:
            om = _OMEGA_EARTH if rotation_rate is None else rotation_rate
            f0, beta = 2 * om * _sind(latitude), 2 * om * _cosd(latitude) / (_R_EARTH if radius is None else radius)
        self.f0, self.beta = f0, beta


class NonTraditionalBetaPlane:
    """NonTraditionalBetaPlane(fz, fy, β, γ, radius | rotation_rate, latitude, radius)   src/Coriolis/non_traditional_beta_plane.jl:16-77"""

    def __init__(self, FT=np.float64, fz=None, fy=None, beta=None, gamma=None, rotation_rate=None, latitude=None, radius=None, **kw):
        beta, gamma = kw.pop("β", beta), kw.pop("γ", gamma)
        if kw:
            raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
        fs = (fz, fy, beta, gamma)
        use_f = (not all(c is None for c in fs)) and latitude is None
        use_planet = latitude is not None and all(c is None for c in fs)
        if use_f == use_planet:
            raise ValueError("Either the keywords fz, fy, β, γ, and radius must be specified, *or* all of rotation_rate, latitude, and radius.")
        radius = _R_EARTH if radius is None else radius
        if use_planet:
            om = _OMEGA_EARTH if rotation_rate is None else rotation_rate
            fz, fy = 2 * om * _sind(latitude), 2 * om * _cosd(latitude)
            beta, gamma = 2 * om * _cosd(latitude) / radius, -4 * om * _sind(latitude) / radius
        if any(c is None for c in (fz, fy, beta, gamma)):
            raise ValueError("NonTraditionalBetaPlane: fz, fy, β and γ must all be given")      # the reference would fail converting nothing to FT
        self.fz, self.fy, self.beta, self.gamma, self.R = fz, fy, beta, gamma, radius


class ConstantCartesianCoriolis:
    """ConstantCartesianCoriolis(fx, fy, fz | f, rotation_axis | latitude, rotation_rate)   src/Coriolis/constant_cartesian_coriolis.jl:11-67"""

    def __init__(self, FT=np.float64, fx=None, fy=None, fz=None, f=None, rotation_axis=None, latitude=None, rotation_rate=None):
        comps = (fx, fy, fz)
        if latitude is not None:
            if any(c is not None for c in comps) or f is not None:
                raise ValueError("Only `rotation_rate` can be specified when using `latitude`.")
            om = _OMEGA_EARTH if rotation_rate is None else rotation_rate
            fx, fy, fz = 0.0, 2 * om * _cosd(latitude), 2 * om * _sind(latitude)
        elif f is not None:
            if any(c is not None for c in comps):
                raise ValueError("Only `rotation_axis` can be specified when using `f`.")
            if rotation_axis is None:                                   # ZDirection()
                fx, fy, fz = 0.0, 0.0, f
            else:
                ax = np.asarray(rotation_axis, dtype=np.float64)
                if ax.shape != (3,) or not np.isclose(float(np.sqrt((ax ** 2).sum())), 1.0):      # validate_unit_vector
                    raise ValueError("unit vector must be unitary")
                fx, fy, fz = (f * float(a) for a in ax)
        elif not all(c is not None for c in comps):
            raise ValueError("Either (i) `latitude`, or (ii) `f`, or (iii) `fx`, `fy` and `fz` must be specified.")
        self.fx, self.fy, self.fz = fx, fy, fz


class _BC:
    def __init__(self, kind, value):
        self.array = None
        if callable(value):
            raise NotImplementedError("function-valued boundary conditions are out of scope (Julia closures cannot cross the C ABI)")
        if value is not None and not np.isscalar(value):
            if kind not in (L.OC_BC_FLUX, L.OC_BC_VALUE, L.OC_BC_GRADIENT):
                raise NotImplementedError("array-valued boundary conditions: Flux, Value or Gradient")
            self.array = np.asarray(value)
            if self.array.ndim != 2:
                raise ValueError("an array-valued boundary condition needs a 2-D array over the two tangential dimensions")
            value = 0.0
        self.kind, self.value = kind, value


def FluxBoundaryCondition(value):
    return _BC(L.OC_BC_FLUX, value)


def ValueBoundaryCondition(value):
    return _BC(L.OC_BC_VALUE, value)


def GradientBoundaryCondition(value):
    return _BC(L.OC_BC_GRADIENT, value)


def OpenBoundaryCondition(value=None):
    return _BC(L.OC_BC_OPEN, value)


_SIDES = ("west", "east", "south", "north", "bottom", "top")


class FieldBoundaryConditions:
    def __init__(self, **sides):
        for k in sides:
            if k not in _SIDES:
                raise ValueError(f"unknown side {k}")
        self.sides = sides


class Clock:
    def __init__(self, model):
        self._m = model

    def _get(self):
        c = L.oc_clock()
        self._m._lib.check(self._m._lib.oc_get_clock(self._m._h, C.byref(c)))
        return c

    time = property(lambda s: s._get().time)
    iteration = property(lambda s: s._get().iteration)
    stage = property(lambda s: s._get().stage)
    last_Δt = property(lambda s: s._get().last_dt)
    last_dt = property(lambda s: s._get().last_dt)
    last_stage_Δt = property(lambda s: s._get().last_stage_dt)


class UploadTicket:
    """One asynchronous `set!(field, array)` in flight (oc_upload_begin): the values travel from a page-locked copy of `array` to the device
    on a separate stream and land in the field in stream order with the time stepping.  `.wait()` returns once the field holds them."""

    def __init__(self, model, fid, array):
        lib = model._lib
        self._lib, self._h = lib, model._h
        FT = model.grid.FT
        a = np.asfortranarray(np.asarray(array, dtype=FT))
        self.nbytes = a.nbytes
        self._ptr = C.c_void_p()
        lib.check(lib.oc_host_alloc(C.byref(self._ptr), max(self.nbytes, 1)))
        C.memmove(self._ptr, a.ctypes.data, self.nbytes)
        t = C.c_int()
        try:
            lib.check(lib.oc_upload_begin(self._h, fid, self._ptr, self.nbytes, C.byref(t)))
        except Exception:
            lib.oc_host_free(self._ptr)
            self._ptr = None
            raise
        self._ticket = t.value

    def done(self):
        if self._ptr is None:
            return True
        d = C.c_int()
        self._lib.check(self._lib.oc_output_test(self._h, self._ticket, C.byref(d)))
        return bool(d.value)

    def wait(self):
        if self._ptr is not None:
            self._lib.check(self._lib.oc_output_wait(self._h, self._ticket))
            self._lib.oc_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.wait()
        except Exception:
            pass


class OutputTicket:
    """One asynchronous output in flight (oc_output_begin / oc_output_wait / oc_output_test)."""

    def __init__(self, model, fid, lo, n):
        lib = model._lib
        self._lib, self._h, self.shape = lib, model._h, tuple(n)
        FT = model.grid.FT
        self.nbytes = int(np.prod(n)) * np.dtype(FT).itemsize
        self._ptr = C.c_void_p()
        lib.check(lib.oc_host_alloc(C.byref(self._ptr), max(self.nbytes, 1)))            # page-locked: the copy is truly asynchronous
        ctype = C.c_double if FT is np.float64 else C.c_float
        self._view = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(ctype)), shape=(int(np.prod(n)),))
        t = C.c_int()
        try:
            lib.check(lib.oc_output_begin(self._h, fid, (C.c_int * 3)(*lo), (C.c_int * 3)(*n), self._ptr, self.nbytes, C.byref(t)))
        except Exception:
            lib.oc_host_free(self._ptr)
            self._ptr = None
            raise
        self._ticket, self._result = t.value, None

    def done(self):
        if self._result is not None:
            return True
        d = C.c_int()
        self._lib.check(self._lib.oc_output_test(self._h, self._ticket, C.byref(d)))
        return bool(d.value)

    def wait(self):
        """block until the host copy is complete; returns the array indexed [i, j, k]"""
        if self._result is None:
            self._lib.check(self._lib.oc_output_wait(self._h, self._ticket))
            self._result = np.array(self._view, copy=True).reshape(self.shape, order="F")
            self._lib.oc_host_free(self._ptr)
            self._ptr = None
        return self._result

    def __del__(self):
        try:
            if getattr(self, "_ptr", None) is not None and self._result is None:
                self._lib.oc_output_wait(self._h, self._ticket)
                self._lib.oc_host_free(self._ptr)
        except Exception:
            pass


class Field:
    """A handle to a device field.  `interior(f)` / `parent(f)` return host copies (numpy, indexed [i, j, k])."""

    def __init__(self, model, fid, name):
        self.model, self.id, self.name = model, fid, name

    def info(self):
        info = L.oc_field_info()
        self.model._lib.check(self.model._lib.oc_field_info_get(self.model._h, self.id, C.byref(info)))
        return info

    @property
    def location(self):
        return tuple(Face if x else Center for x in self.info().location)

    def _download(self, parent_):
        info = self.info()
        shape = tuple(info.parent_size if parent_ else info.interior_size)
        a = np.empty(shape, dtype=self.model.grid.FT, order="F")
        fn = self.model._lib.oc_download_parent if parent_ else self.model._lib.oc_download_interior
        self.model._lib.check(fn(self.model._h, self.id, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    def interior(self):
        return self._download(False)

    def begin_output(self, indices=None):
        """Asynchronous `Array(interior(field)[indices...])` for output writers (oc_output_begin): snapshots the index box now, in
        stream order with the time stepping, and copies it to page-locked host memory on a separate stream.  `indices`: a 3-tuple of
        slices / integers over the interior (0-based; None = the whole interior).  Returns an OutputTicket; `.wait()` gives the array."""
        info = self.info()
        size = tuple(info.interior_size)
        idx = (slice(None),) * 3 if indices is None else tuple(indices)
        lo, n = [], []
        for d in range(3):
            if isinstance(idx[d], slice):
                a, b, st = idx[d].indices(size[d])
                if st != 1:
                    raise ValueError("output slices must have unit stride")
                lo.append(a); n.append(b - a)
            else:
                lo.append(int(idx[d])); n.append(1)
        return OutputTicket(self.model, self.id, lo, n)

    def begin_set(self, array):
        """Asynchronous `set!(field, array)` (oc_upload_begin): returns an UploadTicket at once; the interior values are in place — in
        stream order — for every entry point called afterwards.  Like `set_parent` / oc_upload_interior it neither fills halos nor
        projects: follow it with `oc_set_finalize` (set_(model) does both synchronously)."""
        info = self.info()
        shape = tuple(info.interior_size)
        a = np.asarray(array)
        if a.shape != shape:
            a = a.reshape(shape)
        return UploadTicket(self.model, self.id, a)

    def maximum_abs(self):
        """maximum(abs, interior(field)), reduced on the device (oc_field_maximum_abs): an 8-byte copy instead of the field"""
        out = C.c_double()
        self.model._lib.check(self.model._lib.oc_field_maximum_abs(self.model._h, self.id, C.byref(out)))
        return out.value

    def parent(self):
        return self._download(True)

    def set(self, value):
        """set!(field, array | number | f(x, y, z))   src/Fields/set!.jl:34-121"""
        g = self.model.grid
        info = self.info()
        shape = tuple(info.interior_size)
        if callable(value):
            locs = [Face if x else Center for x in info.location]
            nodes = [g.nodes(d, locs[d]) if g.topology[d] is not Flat else np.zeros(1) for d in range(3)]
            if self.model.distributed:      # this rank's rows
                # (the last rank of a Bounded partitioned dimension also owns the wall face of a Face field there: interior_size = N_l + 1)
                for d in (0, 1):
                    j0, _ = g.architecture.local_range(d, g.N[d])
                    nodes[d] = nodes[d][j0:j0 + info.interior_size[d]]
            X = np.meshgrid(*nodes, indexing="ij")
            args = [A for d, A in enumerate(X) if g.topology[d] is not Flat]
            value = np.broadcast_to(np.asarray(value(*args), dtype=np.float64), X[0].shape)
        if np.isscalar(value):
            a = np.full(shape, value, dtype=g.FT, order="F")
        else:
            a = np.asarray(value)
            if a.size != int(np.prod(shape)):
                raise ValueError(f"cannot set {self.name} of size {shape} from an array of shape {a.shape}")   # set!.jl:105-112
            a = np.asfortranarray(a.reshape(shape).astype(g.FT))
        self.model._lib.check(self.model._lib.oc_upload_interior(self.model._h, self.id, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def set_parent(self, a):
        info = self.info()
        a = np.asfortranarray(np.asarray(a).reshape(tuple(info.parent_size)).astype(self.model.grid.FT))
        self.model._lib.check(self.model._lib.oc_upload_parent(self.model._h, self.id, a.ctypes.data_as(C.c_void_p), a.nbytes))


def interior(f):
    return f.interior()


def parent(f):
    return f.parent()


class _NT(dict):
    """NamedTuple-like: attribute and key access"""
    __getattr__ = dict.__getitem__


class NonhydrostaticModel:
    def __init__(self, *, grid, advection="default", closure=None, tracers=(), buoyancy=None, coriolis=None,
                 timestepper="RungeKutta3", boundary_conditions=None, forcing=None, stokes_drift=None,
                 background_fields=None, biogeochemistry=None, particles=None, library=None):
        for name, val in (("forcing", forcing), ("stokes_drift", stokes_drift), ("background_fields", background_fields),
                          ("biogeochemistry", biogeochemistry), ("particles", particles)):
            if val:
                raise NotImplementedError(f"{name} is out of scope of the B200 path (Julia closures / other subsystems)")
        self._lib = library if library is not None else L.load()
        # tracers = :c, (:T, :S), () or nothing   (tracernames: src/Fields/field_tuples.jl)
        tracers = () if tracers is None else ((tracers,) if isinstance(tracers, str) else tuple(tracers))
        if boundary_conditions is not None and not isinstance(boundary_conditions, dict):
            raise TypeError("boundary_conditions must be a NamedTuple-like mapping field name -> FieldBoundaryConditions")
        if len(tracers) > L.OC_MAX_TRACERS:
            raise ValueError("too many tracers")
        # advection = Centered() by default (nonhydrostatic_model.jl:117); advection = nothing (None) switches advection off
        if isinstance(advection, str) and advection == "default":
            advection = Centered()
        elif advection is None:
            advection = _NoAdvection()
        closures = () if closure is None else (tuple(closure) if isinstance(closure, (tuple, list)) else (closure,))
        # adapt_advection_order, then inflate_grid_halo_size — per direction   nonhydrostatic_model.jl:175-184,248-262
        advection = adapt_advection_order(advection, grid)
        cneed = 1
        for c in closures:
            cneed = max(cneed, 2 if isinstance(c, (AnisotropicMinimumDissipation, Smagorinsky)) else 1)
        H = tuple(max(grid.H[d], required_halo_size(advection, d), cneed) if grid.topology[d] is not Flat else 0 for d in range(3))
        if isinstance(advection, FluxFormAdvection):
            # The scheme of flux direction d interpolates the advecting velocity ALONG every other direction c (`_advective_momentum_flux_Uv
            # (…, scheme.x, …)`: ℑy of U with the x scheme, upwind_biased_advective_fluxes.jl:47-53) — Centered(4), two points deep, for
            # the fifth-order schemes and Centered(4).  The adapted scheme only gives c the halo of ITS lowered scheme: with H[c] = 1 the
            # reference reads outside the halo (unspecified values).  Nothing to be bit-compatible with: refused.
            for d in range(3):
                sd = advection.dirs[d]        # its advecting_velocity_scheme is Centered(order - 1) (Centered: itself): buffer (order - 1) / 2
                deep = 1 if isinstance(sd, _NoAdvection) else (sd.order // 2 if isinstance(sd, Centered) else max(1, (sd.order - 1) // 2))
                for c in range(3):
                    if c != d and grid.topology[c] is not Flat and grid.topology[d] is not Flat and H[c] < deep:
                        raise NotImplementedError(f"adapt_advection_order: the {'xyz'[d]}-scheme interpolates velocities two points deep along "
                                                  f"{'xyz'[c]}, where the adapted grid has a halo of {H[c]} (the reference reads outside the halo)")
        if H != grid.H:
            grid = grid.with_halo(H)
        self.grid, self.advection, self.closure, self.buoyancy, self.coriolis = grid, advection, closure, buoyancy, coriolis
        self.timestepper_name = {"RungeKutta3": "RungeKutta3", "QuasiAdamsBashforth2": "QuasiAdamsBashforth2"}.get(timestepper)
        if self.timestepper_name is None:
            raise ValueError(f"unknown timestepper {timestepper}")
        cfg = L.oc_config()
        self._lib.oc_config_init(C.byref(cfg))
        cfg.float_type = L.OC_F64 if grid.FT is np.float64 else L.OC_F32
        for d in range(3):
            cfg.N[d], cfg.H[d], cfg.topology[d] = grid.N[d], grid.H[d], grid.topology[d].code
            cfg.delta[d], cfg.extent[d] = grid.D[d], grid.L[d]
        if grid.z_faces is not None:
            self._z_faces = grid.z_faces                            # keep the buffer alive across oc_model_create
            cfg.z_stretched = 1
            cfg.z_faces = self._z_faces.ctypes.data_as(C.POINTER(C.c_double))
            cfg.delta[2] = 0.0
        arch = grid.architecture
        self.distributed = isinstance(arch, Distributed) and arch.nranks > 1
        if self.distributed:
            # local grid of this rank: Ny / R rows (distributed_grids.jl:75-126); the spacing and the global extent stay
            part = arch.partition
            if grid.N[0] % part.x != 0 or grid.N[1] % part.y != 0:
                raise ValueError("Nx and Ny must be divisible by the number of ranks along x and y")
            cfg.N[0], cfg.N[1] = grid.N[0] // part.x, grid.N[1] // part.y
            cfg.dist_rank, cfg.dist_nranks, cfg.dist_ranks_x = arch.rank, arch.nranks, part.x
        cfg.advection = advection.code
        if isinstance(advection, FluxFormAdvection):
            cfg.has_advection_dir = 1
            for d in range(3):
                cfg.advection_dir[d] = advection.dirs[d].code
        cfg.timestepper = L.OC_RK3 if timestepper == "RungeKutta3" else L.OC_AB2
        cfg.n_tracers = len(tracers)
        sd = [c for c in closures if isinstance(c, ScalarDiffusivity)]
        amd = [c for c in closures if isinstance(c, AnisotropicMinimumDissipation)]
        smag = [c for c in closures if isinstance(c, Smagorinsky)]
        sda = [c for c in sd if c.is_array]              # array-valued coefficients: they live in the diffusivity fields (νₑ, κₑ)
        sd = [c for c in sd if not c.is_array]
        if len(sd) > 1 or len(sda) > 1 or len(amd) + len(smag) + len(sda) > 1 or len(sd) + len(sda) + len(amd) + len(smag) != len(closures):
            raise NotImplementedError("closure must be ScalarDiffusivity, AnisotropicMinimumDissipation, Smagorinsky or a tuple of a constant "
                                      "ScalarDiffusivity and one of: an eddy-viscosity closure, an array-valued ScalarDiffusivity")
        if sda:
            cfg.array_diffusivity = 1
        pick = lambda v, n: float(v[n]) if isinstance(v, dict) else float(v)
        if sd:
            cfg.has_scalar_diffusivity, cfg.nu = 1, float(sd[0].nu)
            for t, n in enumerate(tracers):
                cfg.kappa[t] = pick(sd[0].kappa, n)
        if amd:
            cfg.has_amd, cfg.amd_Cnu = 1, float(amd[0].Cnu)
            if amd[0].Cb is not None:
                cfg.amd_has_Cb, cfg.amd_Cb = 1, float(amd[0].Cb)
            for t, n in enumerate(tracers):
                cfg.amd_Ckappa[t] = pick(amd[0].Ckappa, n)
        if smag:
            co = smag[0].coefficient
            if isinstance(co, LillyCoefficient):
                cfg.smagorinsky, cfg.smag_C, cfg.smag_Cb = 2, float(co.smagorinsky), float(co.reduction_factor)
            else:
                cfg.smagorinsky, cfg.smag_C = 1, float(co)
            for t, n in enumerate(tracers):
                cfg.smag_Pr[t] = pick(smag[0].Pr, n)
        if isinstance(buoyancy, BuoyancyForce):
            if buoyancy.gravity_unit_vector is not None:
                cfg.tilted_gravity = 1
                for d in range(3):
                    cfg.gravity_unit_vector[d] = buoyancy.gravity_unit_vector[d]
            buoyancy = buoyancy.formulation
        if buoyancy is not None:
            for n in buoyancy.required:
                if n not in tracers:
                    raise ValueError(f"buoyancy model requires tracer {n}")         # validate_buoyancy
            if isinstance(buoyancy, SeawaterBuoyancy):
                cfg.buoyancy = L.OC_BUOYANCY_SEAWATER_LINEAR
                cfg.gravity, cfg.thermal_expansion, cfg.haline_contraction = buoyancy.g, buoyancy.eos.thermal_expansion, buoyancy.eos.haline_contraction
                cfg.tracer_T, cfg.tracer_S = tracers.index("T"), tracers.index("S")
            elif isinstance(buoyancy, BuoyancyTracer):
                cfg.buoyancy, cfg.tracer_b = L.OC_BUOYANCY_TRACER, tracers.index("b")
            else:
                raise NotImplementedError("unsupported buoyancy model")
        if coriolis is not None:
            if isinstance(coriolis, FPlane):
                cfg.has_coriolis, cfg.coriolis_f = L.OC_CORIOLIS_FPLANE, float(coriolis.f)
            elif isinstance(coriolis, BetaPlane):
                cfg.has_coriolis, cfg.coriolis_f, cfg.coriolis_beta = L.OC_CORIOLIS_BETAPLANE, float(coriolis.f0), float(coriolis.beta)
                cfg.origin_y = float(grid.x0[1])
            elif isinstance(coriolis, ConstantCartesianCoriolis):
                cfg.has_coriolis = L.OC_CORIOLIS_CARTESIAN
                cfg.coriolis_fxyz[0], cfg.coriolis_fxyz[1], cfg.coriolis_fxyz[2] = float(coriolis.fx), float(coriolis.fy), float(coriolis.fz)
            elif isinstance(coriolis, NonTraditionalBetaPlane):
                cfg.has_coriolis = L.OC_CORIOLIS_NONTRADITIONAL_BETAPLANE
                cfg.coriolis_fxyz[1], cfg.coriolis_fxyz[2] = float(coriolis.fy), float(coriolis.fz)
                cfg.coriolis_beta, cfg.coriolis_gamma, cfg.coriolis_radius = float(coriolis.beta), float(coriolis.gamma), float(coriolis.R)
                cfg.origin_y, cfg.origin_z = float(grid.x0[1]), float(grid.x0[2])
            else:
                raise NotImplementedError("Coriolis must be FPlane, BetaPlane, ConstantCartesianCoriolis or NonTraditionalBetaPlane "
                                          "(HydrostaticSphericalCoriolis is out of scope)")
        names = ("u", "v", "w") + tracers
        bcs = boundary_conditions or {}
        # (νₑ = …, κₑ = (tracer = …,)): boundary conditions of the diffusivity fields (build_diffusivity_fields), applied after creation
        diff_bcs = {k: bcs[k] for k in ("νₑ", "nu_e", "κₑ", "kappa_e") if k in bcs}
        bcs = {k: v for k, v in bcs.items() if k not in diff_bcs}
        if diff_bcs and not (amd or smag):
            raise ValueError("boundary conditions for νₑ / κₑ need a closure with diffusivity fields (AnisotropicMinimumDissipation, Smagorinsky)")
        for k in bcs:
            if k not in names:
                raise ValueError(f"boundary_conditions given for unknown field {k}")
        for f, n in enumerate(names):
            fb = bcs.get(n)
            if fb is None:
                continue
            for s, side in enumerate(_SIDES):
                bc = fb.sides.get(side)
                if bc is not None:
                    cfg.bcs[f][s].kind = bc.kind
                    cfg.bcs[f][s].has_value = 0 if bc.value is None else 1
                    cfg.bcs[f][s].value = 0.0 if bc.value is None else float(bc.value)
        cfg.device = grid.architecture.device
        self._cfg = cfg
        h = C.c_void_p()
        self._lib.check(self._lib.oc_model_create(C.byref(cfg), C.byref(h)))
        self._h = h
        if self.distributed:
            self._attach_transport(arch)           # before anything that fills halos (array-valued boundary conditions do)
        for f, n in enumerate(names):              # FluxBoundaryCondition(array), ValueBoundaryCondition(array), GradientBoundaryCondition(array)
            fb = bcs.get(n)
            for s, side in enumerate(_SIDES):
                bc = fb.sides.get(side) if fb is not None else None
                if bc is not None and bc.array is not None:
                    self.set_boundary_condition_array(f, s, bc.array)
        for key, fb in diff_bcs.items():
            per_field = {L.OC_FIELD_NU_E: fb} if key in ("νₑ", "nu_e") else \
                {L.OC_FIELD_KAPPA_E0 + tracers.index(n): b for n, b in (fb.items() if isinstance(fb, dict) else vars(fb).items())}
            for fid, b in per_field.items():
                for s, side in enumerate(_SIDES):
                    bc = b.sides.get(side)
                    if bc is None:
                        continue
                    if bc.array is not None:
                        raise NotImplementedError("array-valued boundary conditions on diffusivity fields")
                    self._lib.check(self._lib.oc_set_diffusivity_bc(self._h, fid, s, bc.kind, 0.0 if bc.value is None else float(bc.value)))
        self.tracer_names = tracers
        self.velocities = _NT(u=Field(self, 0, "u"), v=Field(self, 1, "v"), w=Field(self, 2, "w"))
        self.tracers = _NT({n: Field(self, 3 + t, n) for t, n in enumerate(tracers)})
        self.pressures = _NT(pNHS=Field(self, L.OC_FIELD_PNHS, "pNHS"),
                             pHY=Field(self, L.OC_FIELD_PHY, "pHY′") if buoyancy is not None else None)
        self.diffusivity_fields = None
        if amd or smag or sda:
            self.diffusivity_fields = _NT(nu_e=Field(self, L.OC_FIELD_NU_E, "νₑ"),
                                          kappa_e=_NT({n: Field(self, L.OC_FIELD_KAPPA_E0 + t, "κₑ." + n) for t, n in enumerate(tracers)}))
        self.fields = _NT({**self.velocities, **self.tracers})
        self.timestepper = _NT(Gn=_NT({n: Field(self, L.OC_FIELD_GN0 + f, "Gⁿ." + n) for f, n in enumerate(names)}),
                               Gm=_NT({n: Field(self, L.OC_FIELD_GM0 + f, "G⁻." + n) for f, n in enumerate(names)}))
        self.clock = Clock(self)
        if sda:
            # the array coefficients: interior values, then halos like fill_halo_regions!(ν) on a Center field with default BCs
            shape = tuple(self.grid.N)
            local = (slice(None),) * 3
            if self.distributed:                   # coefficient arrays cover the global grid: this rank's share
                (i0, ni), (j0, nj) = arch.local_range(0, shape[0]), arch.local_range(1, shape[1])
                local = (slice(i0, i0 + ni), slice(j0, j0 + nj), slice(None))
            full = lambda v: (np.asarray(v, dtype=self.grid.FT) if isinstance(v, np.ndarray) else np.full(shape, float(v), dtype=self.grid.FT))[local]
            self.diffusivity_fields.nu_e.set(full(sda[0].nu))
            ids = [L.OC_FIELD_NU_E]
            for t, n in enumerate(tracers):
                k = sda[0].kappa
                self.diffusivity_fields.kappa_e[n].set(full(k[n] if isinstance(k, dict) else k))
                ids.append(L.OC_FIELD_KAPPA_E0 + t)
            self._lib.check(self._lib.oc_fill_halo_regions(self._h, (C.c_int * len(ids))(*ids), len(ids), 1))
        # constructor tail: update_state!(model; compute_tendencies=false)   nonhydrostatic_model.jl:241
        self._lib.check(self._lib.oc_update_state(self._h, 0))

    def set_boundary_condition_array(self, field, side, array):
        """(re)load the N₁×N₂ values of an array-valued Flux / Value / Gradient boundary condition (oc_set_bc_array); `array` is indexed [i₁, i₂]"""
        d = side // 2
        t1, t2 = (1 if d == 0 else 0), (1 if d == 2 else 2)
        shape = (self.grid.N[t1], self.grid.N[t2])
        a = np.asarray(array)
        if a.shape != shape:
            raise ValueError(f"boundary-condition array of shape {a.shape}; expected {shape}")
        if self.distributed:                       # this rank's share of the side (a collective call: every rank makes it)
            arch = self.grid.architecture
            (i0, n1), (j0, n2) = arch.local_range(t1, shape[0]), arch.local_range(t2, shape[1])
            a = a[i0:i0 + n1, j0:j0 + n2]
        a = np.asfortranarray(a.astype(self.grid.FT))
        self._lib.check(self._lib.oc_set_bc_array(self._h, field, side, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def _attach_transport(self, arch):
        if arch.exchange is not None:
            fn = arch.exchange

            def trampoline(user, n, sp, rp, tg, sptr, sb, rptr, rb):
                try:
                    return int(fn([(sp[i], rp[i], tg[i], sptr[i], sb[i], rptr[i], rb[i]) for i in range(n)]) or 0)
                except Exception as e:          # never unwind through C
                    import traceback
                    traceback.print_exc()
                    return 1
            self._exchange_cb = L.EXCHANGE_FN(trampoline)      # keep alive
            self._lib.check(self._lib.oc_dist_attach_host(self._h, C.cast(self._exchange_cb, C.c_void_p), None))
            return
        import torch
        import torch.distributed as dist
        ident = (C.c_uint8 * 128)()
        if arch.rank == 0:
            self._lib.check(self._lib.oc_dist_unique_id(ident))
        backend = dist.get_backend()
        t = torch.tensor(list(ident), dtype=torch.uint8, device="cuda" if backend == "nccl" else "cpu")
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        self._lib.check(self._lib.oc_dist_attach_nccl(self._h, ident))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.oc_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # measurement helpers
    def sync(self):
        self._lib.check(self._lib.oc_sync(self._h))

    def launch_count(self):
        return int(self._lib.oc_launch_count(self._h))

    def device_bytes(self):
        b = C.c_int64()
        self._lib.check(self._lib.oc_device_bytes(self._h, C.byref(b)))
        return b.value

    def timers(self, enable=None, reset=False):
        if enable is not None:
            self._lib.check(self._lib.oc_timers_enable(self._h, 1 if enable else 0))
        if reset:
            self._lib.check(self._lib.oc_timers_reset(self._h))
            return None
        ms = (C.c_double * len(L.OC_TIMER_NAMES))()
        n = (C.c_int64 * len(L.OC_TIMER_NAMES))()
        self._lib.check(self._lib.oc_timers_get(self._h, ms, n))
        return {name: (ms[i], n[i]) for i, name in enumerate(L.OC_TIMER_NAMES)}


# ------------------------------------------------------------------------------------------ model methods
def set_(model, enforce_incompressibility=True, **kwargs):
    """set!(model; enforce_incompressibility=true, kwargs...)"""
    for name, value in kwargs.items():
        if name not in model.fields:
            raise ValueError(f"name {name} not found in model.velocities or model.tracers.")   # set_nonhydrostatic_model.jl:41
        model.fields[name].set(value)
    model._lib.check(model._lib.oc_set_finalize(model._h, 1 if enforce_incompressibility else 0))


def time_step_(model, dt, euler=False, callbacks=()):
    """time_step!(model, Δt; callbacks=[], euler=false)"""
    if callbacks:
        raise NotImplementedError("mid-step callbacks need the staged entry points; use update_state_/… directly")
    if model.timestepper_name == "RungeKutta3":
        model._lib.check(model._lib.oc_time_step_rk3(model._h, float(dt)))
    else:
        model._lib.check(model._lib.oc_time_step_ab2(model._h, float(dt), 1 if euler else 0))


def update_state_(model, compute_tendencies=True):
    model._lib.check(model._lib.oc_update_state(model._h, 1 if compute_tendencies else 0))


def compute_tendencies_(model):
    model._lib.check(model._lib.oc_compute_tendencies(model._h))


def compute_flux_bc_tendencies_(model):
    model._lib.check(model._lib.oc_compute_flux_bc_tendencies(model._h))


def rk3_substep_(model, dt, stage):
    model._lib.check(model._lib.oc_rk3_substep(model._h, float(dt), int(stage)))


def ab2_step_(model, dt, chi=0.1):
    model._lib.check(model._lib.oc_ab2_step(model._h, float(dt), float(chi)))


def cache_previous_tendencies_(model):
    model._lib.check(model._lib.oc_cache_previous_tendencies(model._h))


def compute_pressure_correction_(model, dt):
    model._lib.check(model._lib.oc_compute_pressure_correction(model._h, float(dt)))


def make_pressure_correction_(model, dt):
    model._lib.check(model._lib.oc_make_pressure_correction(model._h, float(dt)))


def fill_halo_regions_(fields, fill_open_bcs=True):
    """fill_halo_regions!(field | fields)"""
    fields = [fields] if isinstance(fields, Field) else list(fields)
    model = fields[0].model
    ids = (C.c_int * len(fields))(*[f.id for f in fields])
    model._lib.check(model._lib.oc_fill_halo_regions(model._h, ids, len(fields), 1 if fill_open_bcs else 0))


def solve_poisson(model, rhs):
    """solve!(ϕ, model.pressure_solver, rhs) for a host right-hand side of shape (Nx, Ny, Nz); on a distributed model every rank
    passes (and receives) its own slab of rows, (Nx, Ny / R, Nz) — a collective call, like the reference's solve! on a
    DistributedFFTBasedPoissonSolver (distributed_fft_based_poisson_solver.jl:141-178)"""
    g = model.grid
    shape = list(g.N)
    if getattr(model, "distributed", False):
        shape[0] //= g.architecture.partition.x
        shape[1] //= g.architecture.partition.y
    shape = tuple(shape)
    a = np.asfortranarray(np.asarray(rhs).reshape(shape).astype(g.FT))
    out = np.empty(shape, dtype=g.FT, order="F")
    model._lib.check(model._lib.oc_poisson_solve(model._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), a.nbytes))
    return out


# ------------------------------------------------------------------------------------------ diagnostics
def step_diagnostics(model):
    """One on-device reduction pass: (cell_advection_timescale, max|u|, max|v|, max|w|, hasnan(u)); distributed models reduce across
    ranks (all_reduce(min / max), src/Simulations/time_step_wizard.jl:112)."""
    d = L.oc_diagnostics()
    model._lib.check(model._lib.oc_compute_diagnostics(model._h, C.byref(d)))
    out = dict(cell_advection_timescale=d.cell_advection_timescale, max_abs_u=d.max_abs_u, max_abs_v=d.max_abs_v,
               max_abs_w=d.max_abs_w, has_nan=bool(d.has_nan))
    if getattr(model, "distributed", False):
        import torch
        import torch.distributed as dist
        if dist.is_initialized():
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([-out["cell_advection_timescale"], out["max_abs_u"], out["max_abs_v"], out["max_abs_w"],
                              float(out["has_nan"])], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t = t.cpu().tolist()
            out = dict(cell_advection_timescale=-t[0], max_abs_u=t[1], max_abs_v=t[2], max_abs_w=t[3], has_nan=bool(t[4]))
    return out


def cell_advection_timescale(model):
    """cell_advection_timescale(model)   src/Advection/cell_advection_timescale.jl:13-34"""
    return step_diagnostics(model)["cell_advection_timescale"]


def hasnan(model):
    """hasnan(model) = hasnan(first(fields(model)))   src/Diagnostics/nan_checker.jl"""
    return step_diagnostics(model)["has_nan"]


def cell_diffusion_timescale(model):
    """cell_diffusion_timescale(model)   src/TurbulenceClosures/turbulence_closure_diagnostics.jl:23-25,41-46,57-69,83-84:
    Δ² / max diffusivity with Δ = min(Δx, Δy, Δz) over the non-Flat dimensions (ThreeDimensionalFormulation); closure tuples take the
    minimum; no closure: Inf.  The maxima of νₑ / κₑ are reduced on the device (oc_field_maximum_abs)."""
    g = model.grid
    spac = [g.D[d] for d in range(3) if g.topology[d] is not Flat and not (d == 2 and g.z_faces is not None)]
    if g.z_faces is not None:
        spac.append(float(np.min(np.diff(g.z_faces.astype(g.FT)))))
    delta2 = min(spac) ** 2 if spac else math.inf
    closures = model.closure if isinstance(model.closure, (tuple, list)) else (() if model.closure is None else (model.closure,))
    div = lambda a, b: math.inf if b == 0 else a / b
    pick = lambda v: max(v.values()) if isinstance(v, dict) else v
    out = math.inf
    for c in closures:
        if isinstance(c, ScalarDiffusivity):
            kap = [c.kappa[n] if isinstance(c.kappa, dict) else c.kappa for n in model.tracer_names] or [0.0]
            out = min(out, div(delta2, float(c.nu)), div(delta2, float(max(kap))))
        elif isinstance(c, Smagorinsky):
            prs = [c.Pr[n] if isinstance(c.Pr, dict) else c.Pr for n in model.tracer_names] or [1.0]
            out = min(out, div(delta2, model.diffusivity_fields.nu_e.maximum_abs() * max(1.0, 1.0 / min(prs))))
        elif isinstance(c, AnisotropicMinimumDissipation):
            out = min(out, div(delta2, model.diffusivity_fields.nu_e.maximum_abs()))
            for f in model.diffusivity_fields.kappa_e.values():
                out = min(out, div(delta2, f.maximum_abs()))
    return out


class TimeStepWizard:
    """TimeStepWizard(cfl=0.2, diffusive_cfl=Inf, max_change=1.1, min_change=0.5, max_Δt=Inf, min_Δt=0)
    src/Simulations/time_step_wizard.jl:65-116"""

    def __init__(self, cfl=0.2, diffusive_cfl=math.inf, max_change=1.1, min_change=0.5, max_Δt=math.inf, min_Δt=0.0, max_dt=None, min_dt=None):
        self.diffusive_cfl = diffusive_cfl
        if min_change >= 1:
            raise ValueError(f"min_change must be < 1. You provided min_change = {min_change}.")
        if max_change <= 1:
            raise ValueError(f"max_change must be > 1. You provided max_change = {max_change}.")
        self.cfl, self.max_change, self.min_change = cfl, max_change, min_change
        self.max_dt = max_Δt if max_dt is None else max_dt
        self.min_dt = min_Δt if min_dt is None else min_dt

    def new_time_step(self, old_dt, model):
        new_dt = self.cfl * cell_advection_timescale(model)
        if math.isfinite(self.diffusive_cfl):        # :77-79: the diffusion timescale is evaluated only when a diffusive CFL is given
            new_dt = min(new_dt, self.diffusive_cfl * cell_diffusion_timescale(model))
        new_dt = min(self.max_change * old_dt, new_dt)
        new_dt = max(self.min_change * old_dt, new_dt)
        return min(max(new_dt, self.min_dt), self.max_dt)

    def __call__(self, simulation):
        simulation.Δt = self.new_time_step(simulation.Δt, simulation.model)


# ------------------------------------------------------------------------------------------ Simulation
class Checkpointer:
    """Checkpointer(model; …): the parent arrays of the prognostic fields, of the time stepper's G⁻ and the clock — what the
    reference's Checkpointer writes and `set!(model, filepath)` picks up (src/OutputWriters/checkpointer.jl:161-262).  The container
    here is an .npz file with the reference's property names ("u", "v", "w", tracers, "timestepper/G⁻/<name>", "clock/…"); a JLD2 file
    holds the same arrays (x fastest, halos included), so the Julia extension can hand them over unchanged."""

    def __init__(self, model, prefix="checkpoint"):
        self.model, self.prefix = model, prefix

    def state(self):
        m = self.model
        out = {}
        for n, f in m.fields.items():
            out[n] = f.parent()
        for n, f in m.timestepper.Gm.items():
            out["timestepper/G⁻/" + n] = f.parent()          # (brings the tendencies up to date first, like update_state!)
        for n, f in m.timestepper.Gn.items():                 # the reference's checkpoints hold Gⁿ too (checkpointer.jl:161-200); a pickup
            out["timestepper/Gⁿ/" + n] = f.parent()          # recomputes it from the restored state, so it is written for layout parity only
        c = m.clock._get()
        out["clock/time"], out["clock/iteration"], out["clock/stage"] = np.float64(c.time), np.int64(c.iteration), np.int32(c.stage)
        out["clock/last_Δt"], out["clock/last_stage_Δt"] = np.float64(c.last_dt), np.float64(c.last_stage_dt)
        return out

    def write(self, path=None):
        prefix = self.prefix
        if getattr(self.model, "distributed", False):      # one file per rank: prefix *= "_rank$rank" (output_writer_utils.jl:240-241)
            prefix += f"_rank{self.model.grid.architecture.rank}"
        path = path or f"{prefix}_iteration{self.model.clock.iteration}.npz"
        np.savez(path, **self.state())
        return path

    @staticmethod
    def pickup(model, source):
        """set!(model, filepath): restore fields, G⁻ and the clock (checkpointer.jl:202-262)."""
        z = np.load(source) if isinstance(source, str) else source
        lib, h = model._lib, model._h
        for n, f in model.fields.items():
            f.set_parent(z[n])
        for fidx, n in enumerate(model.fields):
            a = np.asfortranarray(np.asarray(z["timestepper/G⁻/" + n], dtype=model.grid.FT))
            lib.check(lib.oc_restore_previous_tendency(h, fidx, a.ctypes.data_as(C.c_void_p), a.nbytes))
        c = L.oc_clock()
        c.time, c.iteration, c.stage = float(z["clock/time"]), int(z["clock/iteration"]), int(z["clock/stage"])
        c.last_dt, c.last_stage_dt = float(z["clock/last_Δt"]), float(z["clock/last_stage_Δt"])
        lib.check(lib.oc_set_clock(h, C.byref(c)))
        return model


class Simulation:
    """Simulation(model; Δt, stop_iteration=Inf, stop_time=Inf)   src/Simulations/simulation.jl"""

    def __init__(self, model, Δt=None, dt=None, stop_iteration=math.inf, stop_time=math.inf, align_time_step=True):
        self.model = model
        self.Δt = Δt if Δt is not None else dt
        if self.Δt is None:
            raise ValueError("Simulation needs Δt")
        self.stop_iteration, self.stop_time, self.align_time_step = stop_iteration, stop_time, align_time_step
        self.callbacks = {}
        self.running = False

    def add_callback(self, name, fn, every=1):
        self.callbacks[name] = (fn, every)


def run_(sim):
    """run!(simulation)  src/Simulations/run.jl:92-176 — time_step!(sim) until a stop criterion fires."""
    model = sim.model
    sim.running = True
    while sim.running:
        clk = model.clock._get()
        if clk.iteration >= sim.stop_iteration or clk.time >= sim.stop_time:
            break
        dt = sim.Δt
        if sim.align_time_step and math.isfinite(sim.stop_time):       # aligned_time_step  run.jl:41-57
            dt = min(dt, sim.stop_time - clk.time)
        time_step_(model, dt)
        it = clk.iteration + 1
        for fn, every in sim.callbacks.values():
            if it % every == 0:
                fn(sim)
    sim.running = False
