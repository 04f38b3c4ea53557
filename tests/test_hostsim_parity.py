"""CPU tests (`-m "not gpu"`): the kernel sources, compiled as a host simulation (tests/hostsim/README.md), against the
oracle on the same seeded inputs.  This checks kernel logic and the C ABI plumbing without a GPU; the parity tests
proper are tests/test_gpu_parity.py."""
import os
import subprocess

import numpy as np
import pytest

from oceananigans_b200 import _lib
import parity_harness as ph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "liboc_hostsim.so")


@pytest.fixture(scope="module")
def hostsim():
    import __graft_entry__ as ge
    ge.build()                       # rebuilds tests/hostsim/liboc_hostsim.so (and the product library) only when a source is newer
    return _lib.Library(HOSTSIM)


@pytest.mark.parametrize("name,kw", ph.CASES[:-1], ids=[c[0] for c in ph.CASES[:-1]])
def test_hostsim_matches_oracle(hostsim, name, kw):
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", [ph.BENCH_INSTANCE_CASES[0], ph.BENCH_INSTANCE_CASES[3]], ids=["72x40x36 weno F64", "70x35x33 weno AB2"])
def test_hostsim_benchmark_kernel_instances_across_tiles(hostsim, name, kw):
    """the triply periodic 32×16-tile (two cells per thread) instances bench.py measures, across several tiles and z-chunks"""
    ph.check_case(kw, library=hostsim)


@pytest.mark.parametrize("name,kw", ph.UVW_CASES, ids=[c[0] for c in ph.UVW_CASES])
def test_hostsim_fused_momentum_kernel_matches_oracle(hostsim, name, kw):
    """UvwCenteredKernel: u, v and w tendencies + substep of Centered(2) models in one launch"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.ARRAY_DIFFUSIVITY_CASES, ids=[c[0] for c in ph.ARRAY_DIFFUSIVITY_CASES])
def test_hostsim_matches_oracle_with_array_valued_diffusivities(hostsim, name, kw):
    """ScalarDiffusivity(ν = array, κ = (T = array, S = number))  abstract_scalar_diffusivity_closure.jl:323-332"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.WENO_HI_CASES + ph.WENO_HI_F32_CASES, ids=[c[0] for c in ph.WENO_HI_CASES + ph.WENO_HI_F32_CASES])
def test_hostsim_matches_oracle_for_weno7_and_weno9(hostsim, name, kw):
    """WENO(order = 7 | 9): weno_interpolants.jl:81-90,175-185,303-307 (general tile kernel, AdvCoef::hi tables)"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.ADAPT_CASES, ids=[c[0] for c in ph.ADAPT_CASES])
def test_hostsim_matches_oracle_with_adapted_advection_order(hostsim, name, kw):
    """adapt_advection_order: FluxFormAdvection with the scheme lowered where N < buffer (adapt_advection_order.jl:18-96)"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.STRETCHED_CASES[:-1], ids=[c[0] for c in ph.STRETCHED_CASES[:-1]])
def test_hostsim_matches_oracle_on_stretched_grids(hostsim, name, kw):
    """vertically stretched grids — level-dependent metrics in every kernel + FourierTridiagonalPoissonSolver"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.SCHEME_CASES, ids=[c[0] for c in ph.SCHEME_CASES])
def test_hostsim_matches_oracle_for_the_other_advection_schemes(hostsim, name, kw):
    """Centered(4), UpwindBiased(1, 3, 5), WENO(3), advection = nothing (test/test_time_stepping.jl:261-267)"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


def test_hostsim_stretched_poisson_all_topologies(hostsim):
    """solve!(ϕ, ::FourierTridiagonalPoissonSolver, b): the sizes / faces of test/test_poisson_solvers_stretched_grids.jl:28-46"""
    import oceananigans_b200 as ob
    import oracle
    rng = np.random.default_rng(11)
    faces = {8: [1, 2, 4, 7, 11, 16, 22, 29, 37], 9: [1, 2, 4, 7, 11, 16, 22, 29, 37, 51], 4: [1, 2, 3, 4, 5]}
    for topo in ["PPB", "PBB", "BPB", "BBB", "FBB", "FPB", "BFB", "PFB"]:
        for N1, N2, Nz in [(8, 8, 8), (16, 8, 9), (8, 11, 8), (5, 8, 9), (7, 13, 8), (4, 5, 4)]:
            kw, size = {}, []
            if topo[0] != "F":
                kw["x"] = (0.0, 1.0); size.append(N1)
            if topo[1] != "F":
                kw["y"] = (0.0, 1.0); size.append(N2)
            size.append(Nz)
            z = [float(f) for f in faces[Nz]]
            grid = ob.RectilinearGrid(np.float64, size=tuple(size), z=z, topology=tuple(ph.TOPO[c] for c in topo), **kw)
            m = ob.NonhydrostaticModel(grid=grid, library=hostsim)
            om = oracle.OracleModel(oracle.Grid(np.float64, size=tuple(size), z=z, topology=tuple(topo), **kw))
            N = om.grid.N
            dzc = om.grid.dz_at("c", np.arange(1, Nz + 1))
            rhs = rng.standard_normal(N)
            rhs -= (rhs * dzc).sum() / (dzc.sum() * N[0] * N[1])          # compatible source: Σ Δz·b = 0
            a = ob.solve_poisson(m, rhs)
            b = om.solve_poisson_tridiagonal(rhs * dzc)
            assert np.abs(a - b).max() <= 1e-11 * max(np.abs(b).max(), 1.0), (topo, N)


def test_hostsim_poisson_all_topologies(hostsim):
    import oceananigans_b200 as ob
    import oracle
    rng = np.random.default_rng(7)
    for topo in ["PPP", "PPB", "PBP", "BPP", "PBB", "BBP", "BPB", "BBB"]:
        for N in [(8, 6, 7), (16, 16, 16), (1, 5, 4), (7, 1, 1)]:
            grid = ob.RectilinearGrid(np.float64, size=N, extent=(1, 2, 3), topology=tuple(ph.TOPO[c] for c in topo))
            m = ob.NonhydrostaticModel(grid=grid, library=hostsim)
            og = oracle.Grid(np.float64, size=N, extent=(1, 2, 3), topology=tuple(topo))
            om = oracle.OracleModel(og)
            rhs = rng.standard_normal(N)
            rhs -= rhs.mean()
            a, b = ob.solve_poisson(m, rhs), om.solve_poisson(rhs)
            assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1.0), (topo, N)


def test_staged_entry_points_match_fused_step(hostsim):
    """time_step! assembled from the staged C entry points (the reference's own sequence, runge_kutta_3.jl:93-170)
    equals the fused oc_time_step_rk3."""
    import oceananigans_b200 as ob
    kw = dict(N=(12, 10, 8), topo="PPB", scheme="weno", closure="scalar", bcs=True, library=hostsim)
    m1, om = ph.build_pair(**kw)
    m2, _ = ph.build_pair(**kw)
    ic = ph.initial_conditions(om)
    ob.set_(m1, **ic)
    ob.set_(m2, **ic)
    dt = 0.01
    ob.time_step_(m1, dt)
    g = [8 / 15, 5 / 12, 3 / 4]
    z = [0.0, -17 / 60, -5 / 12]
    ob.update_state_(m2, True)
    for s in (1, 2, 3):
        ob.compute_flux_bc_tendencies_(m2)
        ob.rk3_substep_(m2, dt, s)
        sdt = dt * (g[s - 1] + z[s - 1])
        ob.compute_pressure_correction_(m2, sdt)
        ob.make_pressure_correction_(m2, sdt)
        if s < 3:
            ob.cache_previous_tendencies_(m2)
        ob.update_state_(m2, True)
    for n in m1.fields:
        a, b = m1.fields[n].interior(), m2.fields[n].interior()
        assert ph.rel_linf(a, b) < 1e-13, n


def test_clock_matches_reference_semantics(hostsim):
    import oceananigans_b200 as ob
    m, om = ph.build_pair(N=(8, 8, 8), topo="PPB", scheme="centered", closure="none", buoy="none", library=hostsim)
    for _ in range(3):
        ob.time_step_(m, 0.3)
        om.time_step(0.3)
    assert m.clock.iteration == om.clock.iteration == 3
    assert m.clock.time == om.clock.time
    assert m.clock.last_Δt == om.clock.last_dt and m.clock.last_stage_Δt == om.clock.last_stage_dt


@pytest.mark.parametrize("ts", ["QuasiAdamsBashforth2", "RungeKutta3"])
def test_checkpoint_pickup_continues_bit_for_bit(hostsim, ts, tmp_path):
    """Checkpointer + set!(model, filepath) (src/OutputWriters/checkpointer.jl:161-262): a model picked up from the parent arrays,
    G⁻ and the clock continues exactly like the one that kept running (AB2 needs the restored G⁻)."""
    import oceananigans_b200 as ob
    kw = dict(N=(12, 10, 8), topo="PPB", scheme="weno", closure="amd", bcs=True, f=1e-2, ts=ts, library=hostsim)
    m1, om = ph.build_pair(**kw)
    ic = ph.initial_conditions(om)
    ob.set_(m1, **ic)
    dt = 0.004
    for _ in range(2):
        ob.time_step_(m1, dt)
    path = ob.Checkpointer(m1, prefix=str(tmp_path / "ckpt")).write()
    for _ in range(2):
        ob.time_step_(m1, dt)
    m2 = ph.build_product(**kw)
    ob.Checkpointer.pickup(m2, path)
    assert m2.clock.iteration == 2 and m2.clock.time == 2 * dt
    for _ in range(2):
        ob.time_step_(m2, dt)
    for n in m1.fields:
        assert np.array_equal(m1.fields[n].parent(), m2.fields[n].parent()), n
    assert np.array_equal(m1.pressures.pNHS.interior(), m2.pressures.pNHS.interior())
    assert m1.clock.time == m2.clock.time and m1.clock.iteration == m2.clock.iteration == 4


@pytest.mark.parametrize("ts", ["QuasiAdamsBashforth2", "RungeKutta3"])
def test_previous_tendency_survives_interleaved_state_updates(hostsim, ts, tmp_path):
    """G⁻ is cached by a pointer swap; update_state!(compute_tendencies=true), reads of Gⁿ / G⁻ and a pickup followed by
    update_state! (the reference's run!(pickup=true) -> initialize! -> update_state! sequence) between fused steps must not
    disturb it: in the reference they are harmless because cache_previous_tendencies! copies
    (quasi_adams_bashforth_2.jl:116-120).  Every variant must continue bit for bit like the plain run."""
    import oceananigans_b200 as ob
    kw = dict(N=(8, 8, 8), topo="PPB", scheme="centered", closure="scalar", ts=ts, library=hostsim)
    dt = 0.01

    def run(between, pickup_after=None):
        m = ph.build_product(**kw)
        om = ph.build_oracle(**{k: v for k, v in kw.items() if k != "library"})
        ob.set_(m, **ph.initial_conditions(om))
        for s in range(4):
            ob.time_step_(m, dt)
            between(m)
            if pickup_after == s:
                path = ob.Checkpointer(m, prefix=str(tmp_path / f"ck{s}")).write()
                m = ph.build_product(**kw)
                ob.Checkpointer.pickup(m, path)
                ob.update_state_(m, True)
        return m

    ref = run(lambda m: None)
    variants = {
        "update_state(compute_tendencies=true)": run(lambda m: ob.update_state_(m, True)),
        "update_state twice": run(lambda m: (ob.update_state_(m, True), ob.update_state_(m, True))),
        "compute_tendencies": run(lambda m: ob.compute_tendencies_(m)),
        "read Gn": run(lambda m: m.timestepper.Gn["u"].interior()),
        "read Gn, update_state": run(lambda m: (m.timestepper.Gm["v"].interior(), ob.update_state_(m, True))),
        "pickup + update_state": run(lambda m: None, pickup_after=1),
        "update_state, pickup + update_state": run(lambda m: ob.update_state_(m, True), pickup_after=2),
    }
    for what, m in variants.items():
        for n in ref.fields:
            assert np.array_equal(ref.fields[n].parent(), m.fields[n].parent()), (what, n)
        assert np.array_equal(ref.pressures.pNHS.interior(), m.pressures.pNHS.interior()), what


@pytest.mark.parametrize("name,kw", ph.SMAGORINSKY_CASES[:-1], ids=[c[0] for c in ph.SMAGORINSKY_CASES[:-1]])
def test_hostsim_matches_oracle_with_smagorinsky_closures(hostsim, name, kw):
    """Smagorinsky / SmagorinskyLilly (Smagorinskys/smagorinsky.jl:92-108, lilly_coefficient.jl:114-135)"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.AMD_CB_CASES[:-1], ids=[c[0] for c in ph.AMD_CB_CASES[:-1]])
def test_hostsim_matches_oracle_with_amd_buoyancy_modification(hostsim, name, kw):
    """AnisotropicMinimumDissipation(; Cb = 1): anisotropic_minimum_dissipation.jl:62-68, 168-172, 310-323 (AmdKernel<FT, STR, CB = true>)"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


def test_hostsim_amd_buoyancy_modification_changes_the_eddy_viscosity(hostsim):
    """The Cb term is active in the parity cases: νₑ differs from the Cb = nothing model in both directions of the max(0, ·) clip."""
    kw = dict(N=(12, 10, 8), topo="BBB", scheme="centered", buoy="tracer")
    nus = []
    for closure in ("amd", "amdcb"):
        m, om = ph.build_pair(library=hostsim, closure=closure, **kw)
        ic = ph.initial_conditions(om, tracer_noise=30.0)
        ph.ob.set_(m, **ic)
        nus.append(np.array(m.diffusivity_fields.nu_e.interior()))
    plain, cb = nus
    assert np.any((plain == 0) & (cb > 0)) and np.any((plain > 0) & (cb == 0)) and np.any((plain > 0) & (cb > 0) & (plain != cb))


@pytest.mark.parametrize("name,kw", ph.CORIOLIS_CASES, ids=[c[0] for c in ph.CORIOLIS_CASES])
def test_hostsim_matches_oracle_for_the_coriolis_family(hostsim, name, kw):
    """BetaPlane (beta_plane.jl:56-72), ConstantCartesianCoriolis (constant_cartesian_coriolis.jl:70-81)"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


def test_hostsim_nontraditional_beta_plane_rejects_what_it_does_not_cover(hostsim):
    """NonTraditionalBetaPlane needs y- and z-nodes: a Flat y or z is a loud error."""
    ob = ph.ob
    cor = ob.NonTraditionalBetaPlane(fz=0.7, fy=0.5, beta=1.0, gamma=-0.8, radius=5.0)
    flat = ob.RectilinearGrid(size=(8, 8), extent=(1, 1), topology=(ob.Periodic, ob.Flat, ob.Bounded))
    flat_y = ob.RectilinearGrid(size=(8, 8), extent=(1, 1), topology=(ob.Periodic, ob.Bounded, ob.Flat))
    for grid in (flat, flat_y):
        with pytest.raises((ob.OceananigansB200Error, NotImplementedError, ValueError)):
            ob.NonhydrostaticModel(grid=grid, coriolis=cor, library=hostsim)


@pytest.mark.parametrize("name,kw", ph.TILTED_CASES, ids=[c[0] for c in ph.TILTED_CASES])
def test_hostsim_matches_oracle_with_tilted_gravity(hostsim, name, kw):
    """BuoyancyForce(formulation; gravity_unit_vector)  buoyancy_force.jl:47-58, g_dot_b.jl:1-3"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.FLAT_CLOSURE_CASES, ids=[c[0] for c in ph.FLAT_CLOSURE_CASES])
def test_hostsim_matches_oracle_with_eddy_closures_on_two_dimensional_grids(hostsim, name, kw):
    """AnisotropicMinimumDissipation / Smagorinsky(-Lilly) with a Flat dimension"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.WALL_BC_CASES[:-1], ids=[c[0] for c in ph.WALL_BC_CASES[:-1]])
def test_hostsim_matches_oracle_with_lateral_wall_bcs(hostsim, name, kw):
    """Value / Gradient / Flux boundary conditions on west / east / south / north walls"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))


@pytest.mark.parametrize("name,kw", ph.ARRAY_BC_CASES, ids=[c[0] for c in ph.ARRAY_BC_CASES])
def test_hostsim_matches_oracle_with_array_valued_flux_bcs(hostsim, name, kw):
    """FluxBoundaryCondition(J::AbstractArray) on every Bounded side (oc_set_bc_array; compute_flux_bcs.jl:116-163)"""
    ph.check_case(kw, library=hostsim, steps=(1, 3))
