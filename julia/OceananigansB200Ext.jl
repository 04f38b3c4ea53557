# OceananigansB200Ext.jl — Julia glue between Oceananigans.jl (v0.100.5) and liboceananigans_b200.so.
#
# STATUS: written against the reference sources under /root/reference but NOT executed — the build image has no `julia`
# binary (SURVEY.md §0).  The Python host layer (oldoceananigans.jl_b200/api.py) implements the same mapping and is what the
# test-suite runs.  Every `ccall` below binds one symbol of include/oceananigans_b200.h; the struct layouts mirror that header
# field by field (checked for the Python twin by tests/test_cabi.py::test_struct_layouts_match_header).
#
# Design: `B200 <: AbstractSerialArchitecture` keeps a HOST mirror of every field (plain `Array`s, so every host-side feature of
# Oceananigans — output writers, diagnostics, checkpointer, `interior(field)` — keeps working unchanged) and a DEVICE twin of the
# model that owns the state between `fetch!` calls.  The methods the time-stepper calls are overloaded for models on `B200`, so
# no KernelAbstractions kernel is ever launched for them (cf. ext/OceananigansCUDAExt.jl:37-138, which plays the same role for
# CUDA.jl).  There is no fallback: an unsupported configuration throws at model construction.
module OceananigansB200Ext

using Oceananigans
using Oceananigans.Architectures: AbstractSerialArchitecture
using Oceananigans.Grids: RectilinearGrid, Periodic, Bounded, Flat, topology, halo_size
using Oceananigans.Fields: interior
using Oceananigans.TimeSteppers: RungeKutta3TimeStepper, QuasiAdamsBashforth2TimeStepper
using Oceananigans.Models.NonhydrostaticModels: NonhydrostaticModel
using Oceananigans.Advection: Centered, WENO, UpwindBiased, required_halo_size_x
using Oceananigans.TurbulenceClosures: ScalarDiffusivity, AnisotropicMinimumDissipation
using Oceananigans.TurbulenceClosures.Smagorinskys: Smagorinsky, LillyCoefficient
using Oceananigans.BuoyancyFormulations: SeawaterBuoyancy, BuoyancyTracer, LinearEquationOfState, BuoyancyForce
using Oceananigans.Coriolis: FPlane, BetaPlane, ConstantCartesianCoriolis, NonTraditionalBetaPlane
using Oceananigans.BoundaryConditions: BoundaryCondition, Flux, Value, Gradient, Open, Periodic as PeriodicBC

import Oceananigans.Architectures as AC
import Oceananigans.TimeSteppers: time_step!, update_state!
import Oceananigans.Fields: set!

const LIB = get(ENV, "OCEANANIGANS_B200_LIB", "liboceananigans_b200.so")

# ---- architecture ----------------------------------------------------------------------------------------------------------
"`B200(device=0)`: the B200-native architecture.  Host arrays are `Array`s (the mirror); the device twin lives in the library."
struct B200 <: AbstractSerialArchitecture
    device::Int32
end
B200() = B200(0)
AC.array_type(::B200) = Array                       # src/Architectures.jl:59-66
AC.architecture(::B200) = B200()
AC.on_architecture(::B200, a::Array) = a            # ext/OceananigansCUDAExt.jl:66-76
AC.on_architecture(::B200, a::Number) = a
AC.device!(a::B200, i) = nothing
Base.summary(::B200) = "B200"

# ---- C structs (include/oceananigans_b200.h) -----------------------------------------------------------------------------------
const OC_MAX_TRACERS = 8
const OC_MAX_FIELDS = 3 + OC_MAX_TRACERS

struct OcBC
    kind::Int32; has_value::Int32; value::Float64
end

mutable struct OcConfig               # `oc_config`, same field order; NTuple for C arrays
    abi_version::Int32; float_type::Int32
    N::NTuple{3,Int32}; H::NTuple{3,Int32}; topology::NTuple{3,Int32}
    delta::NTuple{3,Float64}; extent::NTuple{3,Float64}
    advection::Int32; timestepper::Int32; ab2_chi::Float64
    n_tracers::Int32
    has_scalar_diffusivity::Int32; nu::Float64; kappa::NTuple{OC_MAX_TRACERS,Float64}
    has_amd::Int32; amd_Cnu::Float64; amd_Ckappa::NTuple{OC_MAX_TRACERS,Float64}
    buoyancy::Int32; gravity::Float64; thermal_expansion::Float64; haline_contraction::Float64
    tracer_T::Int32; tracer_S::Int32; tracer_b::Int32
    has_coriolis::Int32; coriolis_f::Float64
    bcs::NTuple{OC_MAX_FIELDS,NTuple{6,OcBC}}
    device::Int32; dist_rank::Int32; dist_nranks::Int32
    z_stretched::Int32; z_faces::Ptr{Float64}      # ABI v2: vertically stretched grid (Nz+1 faces, read during oc_model_create only)
    smagorinsky::Int32; amd_has_Cb::Int32            # ABI v3: 0 none, 1 Smagorinsky(coefficient::Number), 2 LillyCoefficient
    smag_C::Float64; smag_Cb::Float64; smag_Pr::NTuple{OC_MAX_TRACERS,Float64}
    coriolis_beta::Float64; origin_y::Float64; coriolis_fxyz::NTuple{3,Float64}   # ABI v3: BetaPlane, ConstantCartesianCoriolis
    tilted_gravity::Int32; reserved2::Int32; gravity_unit_vector::NTuple{3,Float64} # ABI v3: BuoyancyForce(…; gravity_unit_vector)
    amd_Cb::Float64                                 # ABI v4: AnisotropicMinimumDissipation(; Cb), with amd_has_Cb
    coriolis_gamma::Float64; coriolis_radius::Float64; origin_z::Float64  # ABI v4: NonTraditionalBetaPlane (has_coriolis = 4)
    has_advection_dir::Int32; advection_dir::NTuple{3,Int32}            # ABI v5: FluxFormAdvection(x, y, z) from adapt_advection_order
    array_diffusivity::Int32                                            # ABI v5: ScalarDiffusivity with array-valued ν / κ
    dist_ranks_x::Int32                                                 # ABI v6: Partition(Rx, Ry); 0 / 1 = slabs in y
    OcConfig() = new()
end

struct OcClock
    time::Float64; iteration::Int64; stage::Int32; last_dt::Float64; last_stage_dt::Float64
end

check(status) = status == 0 || error("liboceananigans_b200: ", unsafe_string(ccall((:oc_last_error, LIB), Cstring, ())))

# ---- the device twin ----------------------------------------------------------------------------------------------------------
mutable struct DeviceModel
    handle::Ptr{Cvoid}
    function DeviceModel(cfg::OcConfig)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:oc_model_create, LIB), Cint, (Ref{OcConfig}, Ref{Ptr{Cvoid}}), cfg, h))     # oc_model_create
        dm = new(h[])
        finalizer(d -> ccall((:oc_model_destroy, LIB), Cint, (Ptr{Cvoid},), d.handle), dm)       # oc_model_destroy
        return dm
    end
end

const TWINS = WeakKeyDict{Any,DeviceModel}()

topo_code(::Type{Periodic}) = Int32(0); topo_code(::Type{Bounded}) = Int32(1); topo_code(::Type{Flat}) = Int32(2)

bc_record(bc::BoundaryCondition{<:Flux, Nothing}) = OcBC(2, 0, 0.0)
bc_record(bc::BoundaryCondition{<:Flux, <:Number}) = OcBC(2, 1, bc.condition)
bc_record(bc::BoundaryCondition{<:Flux, <:AbstractArray}) = OcBC(2, 1, 0.0)      # values follow through oc_set_bc_array (twin)
bc_record(bc::BoundaryCondition{<:Value, <:AbstractArray}) = OcBC(3, 1, 0.0)
bc_record(bc::BoundaryCondition{<:Gradient, <:AbstractArray}) = OcBC(4, 1, 0.0)
bc_record(bc::BoundaryCondition{<:Value, <:Number}) = OcBC(3, 1, bc.condition)
bc_record(bc::BoundaryCondition{<:Gradient, <:Number}) = OcBC(4, 1, bc.condition)
bc_record(bc::BoundaryCondition{<:Open, Nothing}) = OcBC(5, 0, 0.0)
bc_record(bc::BoundaryCondition{<:PeriodicBC}) = OcBC(1, 0, 0.0)
bc_record(::Nothing) = OcBC(6, 0, 0.0)
bc_record(bc) = throw(ArgumentError("B200: boundary condition $bc is out of scope (functions / arrays / mixed)"))

"Translate `NonhydrostaticModel(; grid, advection, closure, tracers, buoyancy, coriolis, timestepper)` into `oc_config`."
function config(model::NonhydrostaticModel)
    grid = model.grid
    grid isa RectilinearGrid || throw(ArgumentError("B200: only RectilinearGrid"))
    FT = eltype(grid)
    cfg = OcConfig()
    ccall((:oc_config_init, LIB), Cvoid, (Ref{OcConfig},), cfg)                                   # oc_config_init
    cfg.float_type = FT == Float64 ? 0 : 1
    TX, TY, TZ = topology(grid)
    cfg.topology = (topo_code(TX), topo_code(TY), topo_code(TZ))
    cfg.N = Int32.((grid.Nx, grid.Ny, grid.Nz)); cfg.H = Int32.(halo_size(grid))
    grid.Δxᶜᵃᵃ isa Number && grid.Δyᵃᶜᵃ isa Number ||
        throw(ArgumentError("B200: stretched x / y are out of scope (only z can be variably spaced)"))
    zfaces = nothing
    if grid.z.Δᵃᵃᶜ isa Number
        cfg.delta = Float64.((grid.Δxᶜᵃᵃ, grid.Δyᵃᶜᵃ, grid.z.Δᵃᵃᶜ))
    else    # vertically stretched: pass the Nz+1 interior faces; the library regenerates the halo spacings like generate_coordinate
        TZ === Bounded || throw(ArgumentError("B200: a stretched z needs the Bounded topology"))
        zfaces = Float64.(Array(grid.z.cᵃᵃᶠ[1:grid.Nz+1]))
        cfg.delta = Float64.((grid.Δxᶜᵃᵃ, grid.Δyᵃᶜᵃ, 0.0)); cfg.z_stretched = 1
    end
    cfg.extent = Float64.((grid.Lx, grid.Ly, grid.Lz))
    arch = architecture(grid)
    if arch isa Oceananigans.DistributedComputations.Distributed
        # Distributed(B200(); partition = Partition(Rx, Ry)): the local grid's sizes are already this rank's (distributed_grids.jl:75-126),
        # its extents are local too — the library wants the global ones, and the GLOBAL topology (the local LeftConnected /
        # RightConnected / FullyConnected come from a Bounded or Periodic global dimension, :75-126): both from
        # reconstruct_global_grid (distributed_grids.jl:192-233).  Ranks are x-major on both sides (rank2index,
        # distributed_architectures.jl:354-362).
        Rx, Ry, Rz = arch.ranks
        Rz == 1 || throw(ArgumentError("B200: z is never partitioned"))
        gg = Oceananigans.DistributedComputations.reconstruct_global_grid(grid)
        GX, GY, GZ = topology(gg)
        cfg.dist_rank = arch.local_rank; cfg.dist_nranks = Rx * Ry; cfg.dist_ranks_x = Rx
        cfg.extent = Float64.((gg.Lx, gg.Ly, gg.Lz))
        cfg.topology = (topo_code(GX), topo_code(GY), topo_code(GZ))
    end
    # oc_advection code of one scheme (by its buffer = required halo)
    scheme_code(a) = a === nothing ? Int32(7) :
                     a isa Centered && required_halo_size_x(a) == 1 ? Int32(0) : a isa Centered && required_halo_size_x(a) == 2 ? Int32(2) :
                     a isa WENO ? Int32((0, 5, 1, 9, 10)[required_halo_size_x(a)]) :                      # WENO(3), (5), (7), (9); WENO(1) is UpwindBiased(1)
                     a isa UpwindBiased ? Int32((6, 3, 4)[required_halo_size_x(a)]) :
                     throw(ArgumentError("B200: advection must be Centered(order<=4), UpwindBiased(order<=5), WENO(order<=9) or nothing"))
    adv = model.advection.momentum
    if adv isa Oceananigans.Advection.FluxFormAdvection          # adapt_advection_order lowered the scheme in some direction (ABI v5)
        cfg.has_advection_dir = 1
        cfg.advection_dir = (scheme_code(adv.x), scheme_code(adv.y), scheme_code(adv.z))
        cfg.advection = maximum(cfg.advection_dir)              # informational; the per-direction codes are what the kernels use
    else
        cfg.has_advection_dir = 0
        cfg.advection = scheme_code(adv)
    end
    cfg.timestepper = model.timestepper isa RungeKutta3TimeStepper ? 0 :
                      model.timestepper isa QuasiAdamsBashforth2TimeStepper ? 1 : throw(ArgumentError("B200: unsupported time stepper"))
    model.timestepper isa QuasiAdamsBashforth2TimeStepper && (cfg.ab2_chi = model.timestepper.χ)
    names = keys(model.tracers); cfg.n_tracers = length(names)
    closures = model.closure isa Tuple ? model.closure : (model.closure,)
    for c in closures
        c === nothing && continue
        if c isa ScalarDiffusivity
            cfg.has_scalar_diffusivity = 1; cfg.nu = c.ν
            cfg.kappa = ntuple(t -> t <= length(names) ? Float64(c.κ[t]) : 0.0, OC_MAX_TRACERS)
        elseif c isa AnisotropicMinimumDissipation
            cfg.has_amd = 1; cfg.amd_Cnu = c.Cν
            if c.Cb !== nothing      # buoyancy modification (anisotropic_minimum_dissipation.jl:62-68): ABI v4
                cfg.amd_has_Cb = 1; cfg.amd_Cb = c.Cb
            end
            cfg.amd_Ckappa = ntuple(t -> t <= length(names) ? Float64(c.Cκ[t]) : 0.0, OC_MAX_TRACERS)
        elseif c isa Smagorinsky && (c.coefficient isa Number || c.coefficient isa LillyCoefficient)
            # Smagorinsky(coefficient, Pr) / SmagorinskyLilly(C, Cb, Pr)  (Smagorinskys/smagorinsky.jl:31-84, lilly_coefficient.jl:47-112)
            if c.coefficient isa LillyCoefficient
                cfg.smagorinsky = 2; cfg.smag_C = c.coefficient.smagorinsky; cfg.smag_Cb = c.coefficient.reduction_factor
            else
                cfg.smagorinsky = 1; cfg.smag_C = c.coefficient
            end
            cfg.smag_Pr = ntuple(t -> t <= length(names) ? Float64(c.Pr[t]) : 1.0, OC_MAX_TRACERS)
        else
            throw(ArgumentError("B200: closure $(summary(c)) is out of scope"))
        end
    end
    b = model.buoyancy isa BuoyancyForce ? model.buoyancy.formulation : model.buoyancy
    if model.buoyancy isa BuoyancyForce && !(model.buoyancy.gravity_unit_vector isa Oceananigans.Grids.NegativeZDirection)
        cfg.tilted_gravity = 1; cfg.gravity_unit_vector = Float64.(Tuple(model.buoyancy.gravity_unit_vector))   # buoyancy_force.jl:47-58
    end
    if b isa SeawaterBuoyancy
        eos = b.equation_of_state; eos isa LinearEquationOfState || throw(ArgumentError("B200: only LinearEquationOfState"))
        cfg.buoyancy = 2; cfg.gravity = b.gravitational_acceleration
        cfg.thermal_expansion = eos.thermal_expansion; cfg.haline_contraction = eos.haline_contraction
        cfg.tracer_T = findfirst(==(:T), names) - 1; cfg.tracer_S = findfirst(==(:S), names) - 1
    elseif b isa BuoyancyTracer
        cfg.buoyancy = 1; cfg.tracer_b = findfirst(==(:b), names) - 1
    elseif b !== nothing
        throw(ArgumentError("B200: unsupported buoyancy formulation"))
    end
    if model.coriolis isa FPlane
        cfg.has_coriolis = 1; cfg.coriolis_f = model.coriolis.f
    elseif model.coriolis isa BetaPlane            # f = f₀ + β ynode  (src/Coriolis/beta_plane.jl:56-72)
        cfg.has_coriolis = 2; cfg.coriolis_f = model.coriolis.f₀; cfg.coriolis_beta = model.coriolis.β
        cfg.origin_y = TY === Flat ? 0.0 : Float64(grid.yᵃᶠᵃ[1])
    elseif model.coriolis isa ConstantCartesianCoriolis     # src/Coriolis/constant_cartesian_coriolis.jl:70-81
        cfg.has_coriolis = 3; cfg.coriolis_fxyz = Float64.((model.coriolis.fx, model.coriolis.fy, model.coriolis.fz))
    elseif model.coriolis isa NonTraditionalBetaPlane       # src/Coriolis/non_traditional_beta_plane.jl:79-96 (ABI v4)
        c = model.coriolis
        cfg.has_coriolis = 4; cfg.coriolis_fxyz = (0.0, Float64(c.fy), Float64(c.fz))
        cfg.coriolis_beta = c.β; cfg.coriolis_gamma = c.γ; cfg.coriolis_radius = c.R
        cfg.origin_y = Float64(grid.yᵃᶠᵃ[1]); cfg.origin_z = Float64(grid.z.cᵃᵃᶠ[1])
    elseif model.coriolis !== nothing
        throw(ArgumentError("B200: Coriolis must be FPlane, BetaPlane, ConstantCartesianCoriolis or NonTraditionalBetaPlane"))
    end
    fields = (model.velocities..., model.tracers...)
    cfg.bcs = ntuple(OC_MAX_FIELDS) do f
        f > length(fields) && return ntuple(_ -> OcBC(0, 0, 0.0), 6)
        bcs = fields[f].boundary_conditions
        map(bc_record, (bcs.west, bcs.east, bcs.south, bcs.north, bcs.bottom, bcs.top))
    end
    cfg.device = AC.architecture(grid).device
    return cfg, zfaces
end

function twin(model)
    get!(TWINS, model) do
        cfg, zfaces = config(model)
        dm = if zfaces === nothing
            DeviceModel(cfg)
        else
            GC.@preserve zfaces begin                  # oc_model_create copies the faces; the pointer is not kept
                cfg.z_faces = pointer(zfaces)
                DeviceModel(cfg)
            end
        end
        arch = architecture(model.grid)
        if arch isa Oceananigans.DistributedComputations.Distributed
            # one process per GPU: rank 0 makes the 128-byte NCCL id, MPI carries it, every rank joins the communicator inside the library
            # (replaces the MPI halo / transpose machinery of src/DistributedComputations; must precede anything that fills halos)
            id = zeros(UInt8, 128)
            arch.local_rank == 0 && check(ccall((:oc_dist_unique_id, LIB), Cint, (Ptr{UInt8},), id))
            Oceananigans.DistributedComputations.MPI.Bcast!(id, 0, arch.communicator)
            check(ccall((:oc_dist_attach_nccl, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), dm.handle, id))
        end
        # Flux / Value / GradientBoundaryCondition(A::AbstractArray): upload the N₁×N₂ values (getbc(bc, i, j, …) = A[i, j])
        for (f, field) in enumerate((model.velocities..., model.tracers...))
            bcs = field.boundary_conditions
            for (s, bc) in enumerate((bcs.west, bcs.east, bcs.south, bcs.north, bcs.bottom, bcs.top))
                bc isa BoundaryCondition{<:Union{Flux, Value, Gradient}, <:AbstractArray} || continue
                J = Array{eltype(model.grid)}(bc.condition)
                GC.@preserve J check(ccall((:oc_set_bc_array, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Csize_t),
                                           dm.handle, f - 1, s - 1, J, sizeof(J)))
            end
        end
        dm
    end
end

prognostic(model) = (model.velocities..., model.tracers...)

"Host mirror -> device (set!(u::Field, a::Array), src/Fields/set!.jl:101-121)."
function push!(model)
    dm = twin(model)
    for (f, field) in enumerate(prognostic(model))
        a = Array(interior(field))
        GC.@preserve a check(ccall((:oc_upload_interior, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t), dm.handle, f - 1, a, sizeof(a)))
    end
end

"Device -> host mirror: prognostic fields (parent arrays, halos included), pNHS, and the clock."
function fetch!(model)
    dm = twin(model)
    ids = (collect(0:length(prognostic(model))-1)..., 32)
    for (id, field) in zip(ids, (prognostic(model)..., model.pressures.pNHS))
        p = parent(field)
        GC.@preserve p check(ccall((:oc_download_parent, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t), dm.handle, id, p, sizeof(p)))
    end
    clk = Ref{OcClock}()
    check(ccall((:oc_get_clock, LIB), Cint, (Ptr{Cvoid}, Ref{OcClock}), dm.handle, clk))
    c = model.clock
    c.time = clk[].time; c.iteration = clk[].iteration; c.stage = clk[].stage; c.last_Δt = clk[].last_dt; c.last_stage_Δt = clk[].last_stage_dt
    return nothing
end

const B200Model = NonhydrostaticModel{<:Any, <:Any, <:B200}

# set!(model; u=…, v=…, T=…)  src/Models/NonhydrostaticModels/set_nonhydrostatic_model.jl:33-60
function set!(model::B200Model; enforce_incompressibility=true, kwargs...)
    for (name, value) in kwargs
        set!(getproperty(name in keys(model.velocities) ? model.velocities : model.tracers, name), value)   # host mirror (CPU path of set!)
    end
    push!(model)
    check(ccall((:oc_set_finalize, LIB), Cint, (Ptr{Cvoid}, Cint), twin(model).handle, enforce_incompressibility))
    fetch!(model)
    return nothing
end

# time_step!(model, Δt)  src/TimeSteppers/runge_kutta_3.jl:93, quasi_adams_bashforth_2.jl:74
function time_step!(model::NonhydrostaticModel{<:RungeKutta3TimeStepper, <:Any, <:B200}, Δt; callbacks=[])
    isempty(callbacks) || throw(ArgumentError("B200: mid-step callbacks need the staged entry points (oc_update_state, …)"))
    check(ccall((:oc_time_step_rk3, LIB), Cint, (Ptr{Cvoid}, Cdouble), twin(model).handle, Δt))
    clk = Ref{OcClock}(); check(ccall((:oc_get_clock, LIB), Cint, (Ptr{Cvoid}, Ref{OcClock}), twin(model).handle, clk))
    model.clock.time = clk[].time; model.clock.iteration = clk[].iteration; model.clock.last_Δt = clk[].last_dt
    return nothing
end
function time_step!(model::NonhydrostaticModel{<:QuasiAdamsBashforth2TimeStepper, <:Any, <:B200}, Δt; callbacks=[], euler=false)
    isempty(callbacks) || throw(ArgumentError("B200: mid-step callbacks need the staged entry points"))
    check(ccall((:oc_time_step_ab2, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cint), twin(model).handle, Δt, euler))
    clk = Ref{OcClock}(); check(ccall((:oc_get_clock, LIB), Cint, (Ptr{Cvoid}, Ref{OcClock}), twin(model).handle, clk))
    model.clock.time = clk[].time; model.clock.iteration = clk[].iteration; model.clock.last_Δt = clk[].last_dt
    return nothing
end

# update_state!(model, callbacks; compute_tendencies)  update_nonhydrostatic_model_state.jl:20
update_state!(model::B200Model, callbacks=[]; compute_tendencies=true) =
    check(ccall((:oc_update_state, LIB), Cint, (Ptr{Cvoid}, Cint), twin(model).handle, compute_tendencies))

# ---- asynchronous output (SURVEY §8f item 4): what JLD2Writer's fetch_and_convert_output needs, without stalling the step loop --------
"Start copying `interior(field)[i, j, k]` (unit-stride ranges) of the device field `id` to a page-locked host array; returns (ticket, array)."
function begin_output(model::B200Model, id, i::UnitRange, j::UnitRange, k::UnitRange)
    FT = eltype(model.grid)
    n = Cint.((length(i), length(j), length(k))); lo = Cint.((first(i) - 1, first(j) - 1, first(k) - 1))
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:oc_host_alloc, LIB), Cint, (Ref{Ptr{Cvoid}}, Csize_t), p, prod(n) * sizeof(FT)))
    a = unsafe_wrap(Array, Ptr{FT}(p[]), Int.(n))          # free with oc_host_free after use
    t = Ref{Cint}(0)
    check(ccall((:oc_output_begin, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{NTuple{3,Cint}}, Ref{NTuple{3,Cint}}, Ptr{Cvoid}, Csize_t, Ref{Cint}),
                twin(model).handle, id, lo, n, p[], sizeof(a), t))
    return t[], a
end
wait_output(model::B200Model, ticket) = check(ccall((:oc_output_wait, LIB), Cint, (Ptr{Cvoid}, Cint), twin(model).handle, ticket))

# ---- on-device step diagnostics (SURVEY §8f item 2): no full-field device-to-host copies for the TimeStepWizard / NaNChecker ----------
struct OcDiagnostics
    cell_advection_timescale::Float64; max_abs_u::Float64; max_abs_v::Float64; max_abs_w::Float64; has_nan::Int32; pad::Int32
end
function step_diagnostics(model::B200Model)
    d = Ref{OcDiagnostics}()
    check(ccall((:oc_compute_diagnostics, LIB), Cint, (Ptr{Cvoid}, Ref{OcDiagnostics}), twin(model).handle, d))
    return d[]
end
# cell_advection_timescale(model)  src/Advection/cell_advection_timescale.jl:13-34 (TimeStepWizard: src/Simulations/time_step_wizard.jl:101-115)
Oceananigans.Advection.cell_advection_timescale(model::B200Model) = step_diagnostics(model).cell_advection_timescale

"maximum(abs, interior(field)) reduced on the device; `id` = field id of include/oceananigans_b200.h (νₑ = 34, κₑ = 40 + t)"
function field_maximum_abs(model::B200Model, id)
    out = Ref{Cdouble}(0)
    check(ccall((:oc_field_maximum_abs, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Cdouble}), twin(model).handle, id, out))
    return out[]
end
# cell_diffusion_timescale(model)  src/TurbulenceClosures/turbulence_closure_diagnostics.jl:23-25,57-69: the host method is kept for
# ScalarDiffusivity (numbers); for the eddy-viscosity closures maximum(νₑ), maximum(κₑ) come from the device
function Oceananigans.Diagnostics.cell_diffusion_timescale(model::B200Model)
    closures = model.closure isa Tuple ? model.closure : (model.closure,)
    Δ² = Oceananigans.TurbulenceClosures.min_Δxyz(model.grid, Oceananigans.TurbulenceClosures.ThreeDimensionalFormulation())^2
    nt = length(model.tracers)
    τ = Inf
    for c in closures
        if c isa ScalarDiffusivity
            τ = min(τ, Δ² / c.ν, (Δ² / κ for κ in c.κ)...)
        elseif c isa Smagorinsky
            minPr = nt == 0 ? 1 : minimum(c.Pr)
            τ = min(τ, Δ² / (field_maximum_abs(model, 34) * max(1, 1 / minPr)))
        elseif c isa AnisotropicMinimumDissipation
            τ = min(τ, Δ² / field_maximum_abs(model, 34), (Δ² / field_maximum_abs(model, 40 + t - 1) for t in 1:nt)...)
        end
    end
    return τ
end

end # module
