/*
 * oceananigans_b200.h — C ABI of liboceananigans_b200.so
 *
 * A B200-native (sm_100a) implementation of ONE path of Oceananigans.jl v0.100.5: the
 * NonhydrostaticModel time step on a RectilinearGrid (regular, or vertically stretched with the
 * FourierTridiagonalPoissonSolver).  The reference has no FFI; its seam
 * is Julia multiple dispatch on the architecture type (ext/OceananigansCUDAExt.jl:37-138).  Every
 * entry point below names the reference method (file:line under /root/reference) it stands in for;
 * INTEGRATION.md shows the Julia `ccall` binding for each one.
 *
 * Conventions
 *  - plain C symbols, opaque handle, every call returns 0 on success or a negative oc_status;
 *    oc_last_error() returns the message of the last failure on the calling thread.
 *  - unsupported configurations are ERRORS at oc_model_create, never fallbacks (mirrors the
 *    ArgumentErrors of nonhydrostatic_model.jl:141-169).  There is no CPU path in this library.
 *  - the library owns all device memory.  Host buffers are dense, x fastest ("parent" layout of
 *    src/Grids/new_data.jl:15-73): interior buffers are Nx'×Ny'×Nz' with N' = N (+1 for a Face
 *    location in a Bounded dimension); parent buffers add the halos (N' + 2H; Flat: N=1, H=0).
 *  - one caller thread per model; work is asynchronous on the model's stream; oc_sync() waits.
 */
#ifndef OCEANANIGANS_B200_H
#define OCEANANIGANS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OC_ABI_VERSION 6
#define OC_MAX_TRACERS 8
#define OC_MAX_FIELDS (3 + OC_MAX_TRACERS)

typedef enum { OC_OK = 0, OC_ERR_INVALID = -1, OC_ERR_UNSUPPORTED = -2, OC_ERR_CUDA = -3, OC_ERR_STATE = -4 } oc_status;

typedef enum { OC_F64 = 0, OC_F32 = 1 } oc_float_type;
typedef enum { OC_PERIODIC = 0, OC_BOUNDED = 1, OC_FLAT = 2 } oc_topology;          /* src/Grids/Grids.jl:68-104 */
/* Centered(order=2), WENO(order=5): the BASELINE schemes (TMA-staged z-marching kernel); Centered(order=4),
 * UpwindBiased(order=3|5|1), WENO(order=3), advection=nothing: the rest of the family up to order 5
 * (src/Advection/{centered,upwind_biased,weno}_reconstruction.jl), in the general tile kernel */
typedef enum { OC_CENTERED2 = 0, OC_WENO5 = 1, OC_CENTERED4 = 2, OC_UPWIND3 = 3, OC_UPWIND5 = 4, OC_WENO3 = 5, OC_UPWIND1 = 6,
               OC_ADVECTION_NONE = 7, OC_WENO7 = 9, OC_WENO9 = 10 /* 8 is internal (FluxFormAdvection dispatch) */ } oc_advection;
typedef enum { OC_RK3 = 0, OC_AB2 = 1 } oc_timestepper;                              /* src/TimeSteppers */
typedef enum { OC_CORIOLIS_NONE = 0, OC_CORIOLIS_FPLANE = 1, OC_CORIOLIS_BETAPLANE = 2, OC_CORIOLIS_CARTESIAN = 3,
               OC_CORIOLIS_NONTRADITIONAL_BETAPLANE = 4 } oc_coriolis;   /* oc_config.has_coriolis */
typedef enum { OC_BUOYANCY_NONE = 0, OC_BUOYANCY_TRACER = 1, OC_BUOYANCY_SEAWATER_LINEAR = 2 } oc_buoyancy;

/* boundary_condition.jl:8,83-110.  OC_BC_DEFAULT resolves by topology and location
 * (field_boundary_conditions.jl:15-60): Periodic -> periodic, Bounded+Center -> Flux(nothing),
 * Bounded+Face -> Open(nothing) i.e. impenetrable, Flat -> none. */
typedef enum { OC_BC_DEFAULT = 0, OC_BC_PERIODIC = 1, OC_BC_FLUX = 2, OC_BC_VALUE = 3, OC_BC_GRADIENT = 4,
               OC_BC_OPEN = 5, OC_BC_NONE = 6 } oc_bc_kind;
typedef struct { int32_t kind; int32_t has_value; double value; } oc_bc;             /* scalar-valued BCs only */
enum { OC_WEST = 0, OC_EAST = 1, OC_SOUTH = 2, OC_NORTH = 3, OC_BOTTOM = 4, OC_TOP = 5 };

/* Field identifiers.  Prognostic fields first, in the order of prognostic_fields(model). */
enum { OC_FIELD_U = 0, OC_FIELD_V = 1, OC_FIELD_W = 2, OC_FIELD_TRACER0 = 3,
       OC_FIELD_PNHS = 32, OC_FIELD_PHY = 33, OC_FIELD_NU_E = 34, OC_FIELD_KAPPA_E0 = 40,
       OC_FIELD_GN0 = 64 /* + prognostic index */, OC_FIELD_GM0 = 96 /* + prognostic index */ };

/* The plain-data description of `NonhydrostaticModel(; grid, advection, closure, tracers, buoyancy,
 * coriolis, timestepper, boundary_conditions)` (nonhydrostatic_model.jl:115-244). */
typedef struct {
    int32_t abi_version;          /* OC_ABI_VERSION */
    int32_t float_type;           /* oc_float_type: eltype(grid) */
    int32_t N[3];                 /* grid size; 1 in Flat dimensions */
    int32_t H[3];                 /* halo size; 0 in Flat dimensions (already inflated, :184,248-262) */
    int32_t topology[3];          /* oc_topology */
    double  delta[3];             /* Δx, Δy, Δz as the reference computed them (FT value widened); 1 for Flat */
    double  extent[3];            /* Lx, Ly, Lz (for poisson_eigenvalues.jl:8-31) */
    int32_t advection;            /* oc_advection */
    int32_t timestepper;          /* oc_timestepper */
    double  ab2_chi;              /* quasi_adams_bashforth_2.jl:39 (default 0.1) */
    int32_t n_tracers;
    int32_t has_scalar_diffusivity; double nu; double kappa[OC_MAX_TRACERS];        /* ScalarDiffusivity(ν, κ) */
    int32_t has_amd; double amd_Cnu; double amd_Ckappa[OC_MAX_TRACERS];             /* AnisotropicMinimumDissipation */
    int32_t buoyancy;             /* oc_buoyancy */
    double  gravity, thermal_expansion, haline_contraction;                         /* SeawaterBuoyancy + LinearEquationOfState */
    int32_t tracer_T, tracer_S, tracer_b;  /* tracer indices used by the buoyancy model (-1 if unused) */
    int32_t has_coriolis; double coriolis_f;   /* oc_coriolis kind (0 = nothing); FPlane(f) / BetaPlane f₀ */
    oc_bc   bcs[OC_MAX_FIELDS][6];         /* per prognostic field × side */
    int32_t device;               /* CUDA device ordinal */
    /* Distributed(arch; partition = Partition(1, R)): slab decomposition in y (distributed_architectures.jl:242-302).
     * N[1] is the LOCAL size, extent[1] the GLOBAL extent; dist_nranks <= 1 means a serial model.  topology[] and bcs[] describe the
     * GLOBAL domain: with a Bounded y only rank 0 has the south wall and only rank R-1 the north one (the RightConnected /
     * FullyConnected / LeftConnected local grids of distributed_grids.jl:75-126) — the other sides' entries are ignored there, and
     * rank R-1 owns the wall face of y-Face fields (interior_size[1] = N[1] + 1). */
    int32_t dist_rank, dist_nranks;
    /* Vertically stretched grid: RectilinearGrid(…; z = faces::AbstractVector)  (rectilinear_grid.jl:264-291,
     * grid_generation.jl:33-94).  z_stretched = 1: z_faces points to N[2]+1 increasing face positions (FT values widened to
     * double; host memory, read during oc_model_create only); topology[2] must be Bounded; delta[2] is ignored and
     * extent[2] = z_faces[N] - z_faces[0].  The pressure solver is then the FourierTridiagonalPoissonSolver
     * (src/Solvers/fourier_tridiagonal_poisson_solver.jl:74-131; NonhydrostaticModels.jl:35-40). */
    int32_t z_stretched;
    const double* z_faces;
    /* Smagorinsky(coefficient = C, Pr)  (src/TurbulenceClosures/turbulence_closure_implementations/Smagorinskys/smagorinsky.jl:31-84):
     * smagorinsky = 1; SmagorinskyLilly(C, Cb, Pr) = Smagorinsky(coefficient = LillyCoefficient(smagorinsky = C, reduction_factor = Cb))
     * (Smagorinskys/lilly_coefficient.jl:47-112): smagorinsky = 2.  νₑ = ς (C Δᶠ)² √(2Σ²), κₑ = νₑ / Pr[tracer]; νₑ and κₑ are the fields
     * OC_FIELD_NU_E / OC_FIELD_KAPPA_E0 + t, like AMD's.  DynamicCoefficient is not implemented (rejected). */
    int32_t smagorinsky;          /* 0 = none, 1 = constant coefficient, 2 = LillyCoefficient */
    int32_t amd_has_Cb;           /* ABI v4 (the slot was reserved in v3): 1 = AMD buoyancy modification on, see amd_Cb below */
    double  smag_C, smag_Cb;
    double  smag_Pr[OC_MAX_TRACERS];
    /* The Coriolis family (ABI v3).  BetaPlane(f₀, β): f = f₀ + β y at the y-node of the velocity point (src/Coriolis/beta_plane.jl:56-72);
     * origin_y = y of the south face of the GLOBAL domain (ynode = origin_y + (j − ½ | j − 1) Δy).  ConstantCartesianCoriolis(fx, fy, fz)
     * (src/Coriolis/constant_cartesian_coriolis.jl:70-81): all three momentum equations.  Both run in the general tile kernel. */
    double  coriolis_beta, origin_y;
    double  coriolis_fxyz[3];
    /* BuoyancyForce(formulation; gravity_unit_vector = g̃)  (src/BuoyancyFormulations/buoyancy_force.jl:47-58, g_dot_b.jl:1-3):
     * tilted_gravity = 1: ĝ = −g̃ multiplies ℑxᶠ b / ℑyᶠ b in the u / v equations and ℑzᶠ b in the pHY′ integral; general tile kernel. */
    int32_t tilted_gravity;
    int32_t reserved2;
    double  gravity_unit_vector[3];
    /* ABI v4.  AnisotropicMinimumDissipation(; Cb): the buoyancy modification multiplier (anisotropic_minimum_dissipation.jl:62-68,
     * 137; `Cb = nothing` = amd_has_Cb 0).  νₑ = max(0, −Cν δ² (r − Cb ζ) / q) with Cb ζ = Cb_norm_wᵢ_bᵢᶜᶜᶜ / Δᶠz (:168-172, :310-323),
     * b = buoyancy_perturbationᶜᶜᶜ of the model's buoyancy formulation (0 without buoyancy). */
    double  amd_Cb;
    /* ABI v4.  NonTraditionalBetaPlane(fz, fy, β, γ, R)  (src/Coriolis/non_traditional_beta_plane.jl:16-77): has_coriolis = 4 with
     * fy = coriolis_fxyz[1], fz = coriolis_fxyz[2], β = coriolis_beta, γ = coriolis_gamma, R = coriolis_radius;
     * 2Ωʸ = fy (1 − z/R) + γ y, 2Ωᶻ = fz (1 + 2z/R) + β y at the y- and z-nodes of the evaluation points (:79-96); origin_y as for
     * BetaPlane, origin_z = z of the bottom face of the domain (regular z; a stretched grid takes its z-nodes from z_faces).
     * General tile kernel; serial and slab-decomposed models; no Flat y / z. */
    double  coriolis_gamma, coriolis_radius, origin_z;
    /* ABI v5.  adapt_advection_order (src/Advection/adapt_advection_order.jl:18-96): on a grid with fewer points than the scheme's buffer
     * in some direction the reference lowers the scheme THERE (Centered(2N), UpwindBiased(2N-1), WENO(2N-1); WENO(1) = UpwindBiased(1))
     * and steps with FluxFormAdvection(x, y, z).  has_advection_dir = 1: advection_dir[d] is the oc_advection code of the scheme that
     * computes the fluxes through the faces normal to d (`advection` stays the user's scheme); H[d] >= that scheme's buffer and
     * N[d] >= H[d] are then required per direction.  General tile kernel. */
    int32_t has_advection_dir;
    int32_t advection_dir[3];
    /* ABI v5.  ScalarDiffusivity(ν = A, κ = (T = B, …)) with AbstractArray / Field coefficients located at (Center, Center, Center)
     * (abstract_scalar_diffusivity_closure.jl:323-332: νᶜᶜᶜ = ν[i, j, k], νᶠᶠᶜ = ℑxy ν, κᶠᶜᶜ = ℑx κ, … — the interpolations the eddy-viscosity
     * closures use).  array_diffusivity = 1: the model owns the fields OC_FIELD_NU_E and OC_FIELD_KAPPA_E0 + t, never computes them, and
     * the caller fills them with oc_upload_interior / _parent (+ oc_fill_halo_regions, like fill_halo_regions!(ν) in the reference);
     * fields never uploaded are zero.  May be combined with a constant ScalarDiffusivity (a closure tuple), not with AMD / Smagorinsky. */
    int32_t array_diffusivity;
    /* ABI v6.  Partition(Rx, Ry): ranks along x (distributed_architectures.jl:14-18,242-302); 0 or 1 = the slab decomposition in y.
     * dist_nranks = Rx·Ry, dist_rank = rx·Ry + ry (x-major, rank2index :354-362); N[0] and N[1] are LOCAL sizes, extent[] global.
     * The distributed solver needs Nz % Ry == 0 and (N[1]·Ry) % Rx == 0 (distributed_fft_based_poisson_solver.jl:211-229). */
    int32_t dist_ranks_x;
} oc_config;

typedef struct oc_model oc_model;

typedef struct {
    int32_t location[3];          /* 0 = Center, 1 = Face */
    int32_t interior_size[3];     /* N' */
    int32_t parent_size[3];       /* N' + 2H */
    void*   device_ptr;           /* borrowed: element (i,j,k)=(1,1,1) of the field in the INTERNAL padded layout */
    int64_t stride_y, stride_z;   /* internal strides in elements (stride_x = 1) */
} oc_field_info;

typedef struct { double time; int64_t iteration; int32_t stage; double last_dt, last_stage_dt; } oc_clock;

/* ---- life cycle ----  NonhydrostaticModel(...) constructor hooks: nonhydrostatic_pressure_solver
 * (NonhydrostaticModels.jl:35-40), new_data/zeros(arch,…) (src/Grids/new_data.jl:68-73) */
const char* oc_last_error(void);
int  oc_abi_version(void);
void oc_config_init(oc_config* cfg);                                   /* defaults of the reference constructors */
int  oc_model_create(const oc_config* cfg, oc_model** out);
int  oc_model_destroy(oc_model* m);
int  oc_sync(oc_model* m);                                             /* sync_device! ext/OceananigansCUDAExt.jl:136-138 */

/* ---- data movement ----  set!(u::Field, a::Array) src/Fields/set!.jl:101-121 ; Array(interior(f)) / Array(parent(f));
 * on_architecture(CPU(), ·) ext/OceananigansCUDAExt.jl:66-76 */
int  oc_field_info_get(oc_model* m, int field, oc_field_info* info);
int  oc_upload_interior(oc_model* m, int field, const void* host, size_t nbytes);
int  oc_download_interior(oc_model* m, int field, void* host, size_t nbytes);
int  oc_upload_parent(oc_model* m, int field, const void* host, size_t nbytes);
int  oc_download_parent(oc_model* m, int field, void* host, size_t nbytes);

/* Array-valued boundary condition — FluxBoundaryCondition(A::AbstractArray), ValueBoundaryCondition(A), GradientBoundaryCondition(A) —
 * on one side of a prognostic field (src/BoundaryConditions/boundary_condition.jl: getbc(bc, i, j, …) = bc.condition[i, j];
 * compute_flux_bcs.jl:116-163; fill_halo_regions_value_gradient.jl:7-119).  `host` holds N₁×N₂ values of the model's float type over
 * the two tangential interior dimensions (first one fastest: Nx×Ny for bottom / top, Nx×Nz for south / north, Ny×Nz for west / east).
 * The side must have been created with that kind (OC_BC_FLUX / OC_BC_VALUE / OC_BC_GRADIENT) in oc_config; its scalar is then ignored.
 * May be called again to update the values (time-dependent forcing from the host).  Distributed models: the LOCAL N₁×N₂ share; a
 * collective call (Value / Gradient arrays refresh halos), made by every rank after the transport is attached — ranks on which the side
 * is connected to a neighbour ignore the values. */
int  oc_set_bc_array(oc_model* m, int field, int side, const void* host, size_t nbytes);

/* Boundary condition of a DIFFUSIVITY field — `boundary_conditions = (κₑ = (b = FieldBoundaryConditions(bottom = ValueBoundaryCondition(κ₀)),),)`
 * (build_diffusivity_fields, anisotropic_minimum_dissipation.jl:358-372; test/test_boundary_conditions_integration.jl:54-103): `field` is
 * OC_FIELD_NU_E or OC_FIELD_KAPPA_E0 + t of a model with an eddy-viscosity closure, `side` 0 … 5 (west, east, south, north, bottom, top) of a
 * Bounded dimension, `kind` OC_BC_FLUX (the default: no flux, halo = interior), OC_BC_VALUE or OC_BC_GRADIENT with the scalar `value`.
 * The halo fill of the diffusivities (compute_diffusivities! -> fill_halo_regions!) applies it from the next update_state! on. */
int  oc_set_diffusivity_bc(oc_model* m, int field, int side, int kind, double value);

/* ---- staged entry points (the methods time_step! calls; used by parity tests and mid-step callbacks) ---- */
/* fill_halo_regions!(fields...; fill_open_bcs)  src/BoundaryConditions/fill_halo_regions.jl:25-36; `fields` lists field ids */
int  oc_fill_halo_regions(oc_model* m, const int* fields, int nfields, int fill_open_bcs);
/* update_state!(model; compute_tendencies)  update_nonhydrostatic_model_state.jl:20-56 */
int  oc_update_state(oc_model* m, int compute_tendencies);
/* compute_tendencies!(model)  compute_nonhydrostatic_tendencies.jl:18-46 (Gⁿ, without flux-BC terms) */
int  oc_compute_tendencies(oc_model* m);
/* compute_flux_bc_tendencies!  :170-184 */
int  oc_compute_flux_bc_tendencies(oc_model* m);
/* rk3_substep!(model, Δt, γ, ζ)  runge_kutta_3.jl:179-203 ; stage = 1, 2, 3 selects (γ, ζ) */
int  oc_rk3_substep(oc_model* m, double dt, int stage);
/* ab2_step!(model, Δt) with χ  quasi_adams_bashforth_2.jl:127-154 */
int  oc_ab2_step(oc_model* m, double dt, double chi);
/* cache_previous_tendencies!  store_tendencies.jl:12-22 */
int  oc_cache_previous_tendencies(oc_model* m);
/* compute_pressure_correction!(model, Δt)  pressure_correction.jl:8-20 (halo fill of U with open BCs,
 * source term, FFT solve, halo fill of pNHS) */
int  oc_compute_pressure_correction(oc_model* m, double dt);
/* make_pressure_correction!(model, Δt)  pressure_correction.jl:40-53 */
int  oc_make_pressure_correction(oc_model* m, double dt);
/* solve!(ϕ, ::FFTBasedPoissonSolver, b)  fft_based_poisson_solver.jl:95-125 : host rhs (Nx×Ny×Nz, model FT) -> host ϕ;
 * on a stretched grid solve!(ϕ, ::FourierTridiagonalPoissonSolver, b)  fourier_tridiagonal_poisson_solver.jl:204-246
 * (b is multiplied by Δzᶜᶜᶜ inside, like set_source_term!; the volume mean of ϕ is removed) */
int  oc_poisson_solve(oc_model* m, const void* rhs_host, void* phi_host, size_t nbytes);

/* ---- the hot path ---- */
/* set!(model; ...) tail: fill halos, update_state!, projection with Δt = 1, update_state!  set_nonhydrostatic_model.jl:44-57 */
int  oc_set_finalize(oc_model* m, int enforce_incompressibility);
/* time_step!(model::AbstractModel{<:RungeKutta3TimeStepper}, Δt)  runge_kutta_3.jl:93-170 — fused schedule */
int  oc_time_step_rk3(oc_model* m, double dt);
/* time_step!(model::AbstractModel{<:QuasiAdamsBashforth2TimeStepper}, Δt; euler)  quasi_adams_bashforth_2.jl:74-114 */
int  oc_time_step_ab2(oc_model* m, double dt, int euler);
int  oc_get_clock(oc_model* m, oc_clock* clock);
int  oc_set_clock(oc_model* m, const oc_clock* clock);                 /* Checkpointer pickup: checkpointer.jl:202-228 */
/* Checkpointer pickup of the time stepper: set!(timestepper.G⁻, file) (src/OutputWriters/checkpointer.jl:230-262).  `field` is a
 * prognostic index; the host buffer is the PARENT array of G⁻ as oc_download_parent(m, OC_FIELD_GM0 + field, …) returned it.  Call it
 * after the state has been restored with oc_upload_parent: the next time step then uses this G⁻ (QuasiAdamsBashforth2 needs it; Gⁿ is
 * always recomputed from the restored state, like the reference's update_state! after a pickup). */
int  oc_restore_previous_tendency(oc_model* m, int field, const void* parent_host, size_t nbytes);

/* ---- asynchronous output ----
 * What an output writer needs from the device without stalling the step loop: `Array(interior(field))[i₁:i₂, j₁:j₂, k₁:k₂]` of
 * JLD2Writer / NetCDFWriter (src/OutputWriters/jld2_writer.jl: fetch_and_convert_output → on_architecture(CPU(), ·)) as a
 * two-stage copy.  oc_output_begin snapshots the box lo[d] … lo[d] + n[d] − 1 (0-based interior indices; may reach into the halos) into a
 * device staging buffer IN STREAM ORDER with the time stepping (a device-to-device copy at HBM speed), then streams it to `host`
 * (page-locked memory from oc_host_alloc for a truly asynchronous copy) on a separate copy stream; time steps issued afterwards overlap
 * that transfer and do not disturb the snapshot.  oc_output_wait blocks until the host buffer is complete; oc_output_test polls.
 * A ticket is valid until waited for; at most 64 may be in flight. */
int  oc_output_begin(oc_model* m, int field, const int lo[3], const int n[3], void* host, size_t nbytes, int* ticket);
/* The mirror image for input: `set!(field, host_array)` (src/Fields/set!.jl:101-121) without stalling the step loop.  The interior of a
 * prognostic field is streamed from `host` (page-locked) into a device staging buffer on a separate copy stream, and copied into the
 * field IN STREAM ORDER with the time stepping once it has arrived — time steps issued before the call are not disturbed, entry points
 * called afterwards see the new values.  The ticket (same space, same oc_output_wait / oc_output_test) completes when the field has been
 * written: `host` and the staging buffer may be reused then.  Like oc_upload_interior it does not fill halos or project: call
 * oc_set_finalize afterwards.  With two tickets per field in flight the transfer of step n+1's inputs overlaps step n. */
int  oc_upload_begin(oc_model* m, int field, const void* host, size_t nbytes, int* ticket);
int  oc_output_wait(oc_model* m, int ticket);
int  oc_output_test(oc_model* m, int ticket, int* done);

/* ---- on-device step diagnostics (one reduction pass, a 40-byte device-to-host copy) ----
 * cell_advection_timescale(grid, velocities)  src/Advection/cell_advection_timescale.jl:13-34  (TimeStepWizard,
 * src/Simulations/time_step_wizard.jl:101-115); maximum(abs, u|v|w) for progress messages; hasnan(u)  src/Diagnostics/nan_checker.jl.
 * Distributed models return the LOCAL values: reduce them across ranks as the reference does (all_reduce(min, …)). */
typedef struct { double cell_advection_timescale; double max_abs_u, max_abs_v, max_abs_w; int32_t has_nan; int32_t pad; } oc_diagnostics;
int  oc_compute_diagnostics(oc_model* m, oc_diagnostics* out);
/* maximum(abs, interior(field)) of one field, reduced on the device (an 8-byte device-to-host copy); NaN if the field holds a NaN.
 * What cell_diffusion_timescale(model) needs for the eddy-viscosity closures — maximum(νₑ), maximum(κₑ)
 * (src/TurbulenceClosures/turbulence_closure_diagnostics.jl:57-69; the TimeStepWizard's diffusive_cfl,
 * src/Simulations/time_step_wizard.jl:101-108) — and progress messages need for tracers.  Local value on distributed models. */
int  oc_field_maximum_abs(oc_model* m, int field, double* out);

/* ---- multi-GPU: one process per GPU, slab decomposition in y (x, y, z each Periodic or Bounded) ----
 * Replaces Distributed(...) + fill_halo_regions! on distributed fields (src/DistributedComputations/halo_communication.jl:87-333)
 * and DistributedFFTBasedPoissonSolver (distributed_fft_based_poisson_solver.jl:92-188).  Create the model with
 * cfg.dist_rank / cfg.dist_nranks set, then attach a transport before the first halo fill:
 *   oc_dist_unique_id   rank 0 makes the 128-byte NCCL id; the host layer broadcasts it (MPI.bcast / torch.distributed)
 *   oc_dist_attach_nccl ncclCommInitRank inside the library; halo exchange and FFT transposes are grouped ncclSend/ncclRecv
 *   oc_dist_attach_host TEST-ONLY (host simulation build): transfers go through a host callback so that the decomposition logic
 *                       can be exercised by world_size-2 gloo tests on CPU; the CUDA library returns OC_ERR_UNSUPPORTED.
 * Collective calls — every rank makes them, in the same order, like fill_halo_regions! / solve! on distributed fields in the reference:
 * everything that fills halos or solves for the pressure (oc_set_finalize, oc_update_state, oc_fill_halo_regions, oc_time_step_*, the
 * staged entry points, oc_poisson_solve, oc_set_bc_array) and, on grids with a Flat dimension, oc_upload_interior / _parent (they
 * refresh the internal copies of the Flat direction).  Downloads, clocks, timers and diagnostics are local. */
typedef int (*oc_exchange_fn)(void* user, int nmsg, const int* send_peers, const int* recv_peers, const int* tags,
                              void* const* send_ptrs, const size_t* send_bytes, void* const* recv_ptrs, const size_t* recv_bytes);
int  oc_dist_unique_id(void* id128);
int  oc_dist_attach_nccl(oc_model* m, const void* id128);
int  oc_dist_attach_host(oc_model* m, oc_exchange_fn fn, void* user);

/* ---- measurement ---- */
/* Per-kernel-class device timing with CUDA events on the model's stream.  Classes: */
enum { OC_TIMER_TENDENCY = 0, OC_TIMER_HALO = 1, OC_TIMER_POISSON_RHS = 2, OC_TIMER_FFT = 3, OC_TIMER_POISSON_MID = 4,
       OC_TIMER_PROJECTION = 5, OC_TIMER_AUX = 6, OC_TIMER_SUBSTEP = 7, OC_TIMER_COMM = 8, OC_TIMER_COUNT = 9 };
int  oc_timers_enable(oc_model* m, int enable);
int  oc_timers_reset(oc_model* m);
int  oc_timers_get(oc_model* m, double* ms /*[OC_TIMER_COUNT]*/, int64_t* launches /*[OC_TIMER_COUNT]*/);
/* CUDA-event stopwatch on the model's stream (bench.py times K steps with it); stop synchronises the stream. */
int  oc_stopwatch_start(oc_model* m);
int  oc_stopwatch_stop(oc_model* m, double* elapsed_ms);
/* page-locked host buffers for fast oc_upload_* / oc_download_* (cf. unified_array / device_copy_to!
 * ext/OceananigansCUDAExt.jl:89-100) */
int  oc_host_alloc(void** ptr, size_t nbytes);
int  oc_host_free(void* ptr);
int64_t oc_launch_count(oc_model* m);                                   /* kernels launched by this library so far */
int  oc_device_bytes(oc_model* m, int64_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* OCEANANIGANS_B200_H */
