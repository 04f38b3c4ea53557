"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests.

Each test cites the reference test it restates (paths relative to /root/reference).  These run on CPU
(`-m "not gpu"`).  The Julia reference itself cannot run here, so this is the strongest pin available.
"""
import numpy as np
import pytest

from oracle import Grid, OracleModel
from oracle import advection as adv
from oracle import closures as clo
from oracle.grid import BC, Field, fill_halo_regions
from oracle.model import poisson_eigenvalues
from oracle.operators import Ctx, O, dC, ddF


# ----------------------------------------------------------------------------- coefficients (§8c-7)
def test_reconstruction_coefficients_docstring_values():
    # src/Advection/reconstruction_coefficients.jl:75-85 docstring
    c = adv.stencil_coefficients(np.float32, 1, 5)      # uniform_reconstruction_coefficients(Float32, Val(:left), 3)
    expect = np.array([-0.05, 0.45, 0.78333336, -0.21666667, 0.033333335], dtype=np.float32)
    assert np.array_equal(np.array(c, dtype=np.float32)[:4], expect[:4])
    assert abs(float(c[4]) - float(expect[4])) < 1e-7   # last = 1 - sum(others): 0.0333333f0 in the :119 docstring
    c64 = adv.stencil_coefficients(np.float64, 0, 2)
    assert c64 == (0.5, 0.5)                            # uniform_reconstruction_coefficients(Float64, Val(:symmetric), 1)


def test_centered4_and_weno5_coefficients():
    c4 = adv.centered_coefficients(np.float64, 2)
    assert np.allclose(c4, [-1 / 12, 7 / 12, 7 / 12, -1 / 12], rtol=0, atol=1e-15)
    # calc_reconstruction_stencil(Float32, 2, :symmetric, :x) docstring (:112-113)
    c4f = adv.centered_coefficients(np.float32, 2)
    assert c4f[1] == np.float32(0.5833333) and c4f[3] == np.float32(-0.083333336)
    assert abs(float(c4f[0]) - (-0.083333254)) < 1e-8
    w = adv.WENO(np.float64, 5)
    assert np.allclose(w.coeff_p[0], [1 / 3, 5 / 6, -1 / 6], atol=1e-15)
    assert np.allclose(w.coeff_p[1], [-1 / 6, 5 / 6, 1 / 3], atol=1e-15)
    assert np.allclose(w.coeff_p[2], [1 / 3, -7 / 6, 11 / 6], atol=1e-15)
    assert np.allclose(w.cstar, [0.3, 0.6, 0.1], atol=1e-16)
    assert w.buffer_scheme.buffer == 2 and w.buffer_scheme.buffer_scheme.kind == "upwind1"
    assert w.advecting_velocity_scheme.buffer == 2


def test_upwind_biased_stencils_and_polynomial_reproduction():
    # calc_reconstruction_stencil docstrings (src/Advection/reconstruction_coefficients.jl:103-120):
    #   (Float64, 1, :left) = 1.0 ψ[i-1];  (Float32, 1, :right) = 1.0 ψ[i];
    #   (Float32, 3, :left) = 0.0333333 ψ[i-3] - 0.21666667 ψ[i-2] + 0.78333336 ψ[i-1] + 0.45 ψ[i] - 0.05 ψ[i+1]
    u1 = adv.UpwindBiased(np.float64, 1)
    S = [np.array([3.0]), np.array([7.0])]                       # ψ[i-1], ψ[i]
    assert adv._upwind_value(u1, S, np.array([True]))[0] == 3.0 and adv._upwind_value(u1, S, np.array([False]))[0] == 7.0
    u5 = adv.UpwindBiased(np.float32, 5)
    expect = np.array([0.0333333, -0.21666667, 0.78333336, 0.45, -0.05], dtype=np.float32)    # ψ[i-3] … ψ[i+1]
    for n in range(5):
        S = [np.zeros(1, np.float32) for _ in range(6)]          # ψ[i-3] … ψ[i+2]
        S[n] = np.ones(1, np.float32)
        assert adv._upwind_value(u5, S, np.array([True]))[0] == expect[n]
    # the right-biased stencil is the mirror image (ψ[i+2] … ψ[i-2]) up to the rounding of "last = 1 - sum(others)"
    for n in range(5):
        S = [np.zeros(1, np.float32) for _ in range(6)]
        S[5 - n] = np.ones(1, np.float32)
        assert abs(adv._upwind_value(u5, S, np.array([False]))[0] - expect[n]) < 1e-7
    # order of accuracy: cell averages of a polynomial of degree order-1 give its exact face value
    for order in (3, 5):
        sch = adv.UpwindBiased(np.float64, order)
        B = sch.buffer
        for deg in range(order):
            avg = lambda m: ((m + 1.0) ** (deg + 1) - float(m) ** (deg + 1)) / (deg + 1)      # cell [m, m+1]
            S = [np.array([avg(m)]) for m in range(-B, B)]       # ψ[i-B] … ψ[i+B-1]; the face is x = 0
            exact = 1.0 if deg == 0 else 0.0
            for left in (True, False):
                assert abs(adv._upwind_value(sch, S, np.array([left]))[0] - exact) < 1e-12, (order, deg, left)
    c4 = adv.Centered(np.float64, 4)
    for deg in range(4):
        avg = lambda m: ((m + 1.0) ** (deg + 1) - float(m) ** (deg + 1)) / (deg + 1)
        S = [np.array([avg(m)]) for m in range(-2, 2)]
        assert abs(adv._centered_value(c4, S)[0] - (1.0 if deg == 0 else 0.0)) < 1e-12
    # buffer / fallback chains (upwind_biased_reconstruction.jl:57-74, centered_reconstruction.jl:24-45)
    assert u5.buffer == 3 and u5.buffer_scheme.buffer == 2 and u5.buffer_scheme.buffer_scheme.buffer == 1
    assert u5.advecting_velocity_scheme.buffer == 2 and u5.buffer_scheme.advecting_velocity_scheme.buffer == 1
    assert c4.buffer_scheme.buffer == 1


@pytest.mark.parametrize("name", ["upwind1", "centered", "upwind3", "centered4", "upwind5", "weno", "none"])
def test_time_stepping_works_with_advection_scheme(name):
    # test/test_time_stepping.jl:52-58,261-267 (+ incompressibility, which the reference checks separately)
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import parity_harness as ph
    sch = ph.oracle_scheme(name, np.float64)
    g = Grid(np.float64, size=(3, 3, 3), halo=(3, 3, 3), extent=(1, 2, 3), topology=("P", "P", "B"))
    m = OracleModel(g, advection=sch)
    rng = np.random.default_rng(0)
    m.set(u=rng.uniform(-1, 1, (3, 3, 3)), v=rng.uniform(-1, 1, (3, 3, 3)), w=rng.uniform(-1, 1, (3, 3, 4)))
    m.time_step(0.01)
    assert np.isfinite(m.u.interior).all() and np.abs(m.divergence()).max() < 1e-12


def test_halo_inflation_for_advection_schemes():
    # test/test_nonhydrostatic_models.jl:44-67: halos >= 2 for Centered(4) / UpwindBiased(3), >= 3 for WENO() / UpwindBiased(5)
    import oceananigans_b200 as ob
    for sch, need in ((adv.Centered(np.float64, 4), 2), (adv.UpwindBiased(np.float64, 3), 2), (adv.WENO(np.float64, 5), 3),
                      (adv.UpwindBiased(np.float64, 5), 3)):
        g = Grid(np.float64, size=(4, 4, 4), halo=(1, 1, 1), extent=(1, 2, 3), topology=("P", "P", "B"))
        assert OracleModel(g, advection=sch).grid.H == (need,) * 3
        g = Grid(np.float64, size=(4, 4, 4), halo=(1, 3, 4), extent=(1, 2, 3), topology=("P", "P", "B"))
        assert OracleModel(g, advection=sch).grid.H == (need, 3, 4)
    assert ob.Centered(order=4).buffer == 2 and ob.UpwindBiased(order=3).buffer == 2 and ob.UpwindBiased(order=5).buffer == 3
    with pytest.raises(ValueError):
        ob.UpwindBiased(order=4)


def test_weno_reproduces_polynomials():
    # WENO-5 is exact (to round-off of the non-linear weights) for smooth data: for a quadratic all candidate
    # stencils give the exact face value, so the result is exact regardless of the weights.
    g = Grid(np.float64, size=(16, 4, 4), extent=(16, 4, 4), topology=("P", "P", "P"))
    f = Field(g, "ccc")
    x = np.arange(-2, 20) - 0.5                          # cell centres incl. halos (Δ=1): x_i = i - 1/2
    f.data[...] = ((x ** 2)[:, None, None] + 1.0 / 12.0)  # cell average of x² over a unit cell = x² + 1/12
    ctx = Ctx(g, (4, 12), (1, 4), (1, 4))
    sch = adv.WENO(np.float64, 5)
    for left in (True, False):
        val = adv.biased_face(ctx, sch, ctx.field(f), 0, lambda o: np.full(ctx.shape, left))(O)
        xf = (np.arange(4, 13) - 1.0)                    # face i at x = i - 1
        assert np.allclose(val[:, 0, 0], xf ** 2, atol=1e-12)


# ----------------------------------------------------------------------------- halos (§8c-1)
@pytest.mark.parametrize("FT", [np.float32, np.float64])
@pytest.mark.parametrize("N", [(1, 1, 1), (5, 7, 9), (16, 16, 16)])
def test_halo_regions(FT, N):
    # test/test_halo_regions.jl:1-41
    rng = np.random.default_rng(0)
    g = Grid(FT, size=N, extent=(10, 20, 30), halo=(1, 1, 1), topology=("P", "P", "P"))
    f = Field(g, "ccc")
    f.set(rng.random(N))
    d = f.data
    assert (d[0] == 0).all() and (d[-1] == 0).all() and (d[:, 0] == 0).all() and (d[:, -1] == 0).all()
    assert (d[:, :, 0] == 0).all() and (d[:, :, -1] == 0).all()
    g = Grid(FT, size=N, extent=(100, 200, 300), halo=(1, 1, 1), topology=("P", "P", "B"))
    f = Field(g, "ccc")
    f.set(rng.random(N))
    fill_halo_regions(f)
    d = f.data
    Nx, Ny, Nz = N
    I = slice(1, Nx + 1)
    J = slice(1, Ny + 1)
    K = slice(1, Nz + 1)
    assert np.array_equal(d[0:1, J, K], d[Nx:Nx + 1, J, K])
    assert np.array_equal(d[I, 0:1, K], d[I, Ny:Ny + 1, K])
    assert np.array_equal(d[I, J, 0:1], d[I, J, 1:2])
    assert np.array_equal(d[I, J, Nz + 1:Nz + 2], d[I, J, Nz:Nz + 1])


def test_halo_fill_order_periodic_fills_corners():
    # boundary_condition_ordering.jl:113-128: Flux/Value/Gradient first, Periodic afterwards over the full parent
    # extent, so (periodic x) x (bounded z) corners hold the periodic image of the z-halo.
    g = Grid(np.float64, size=(4, 4, 4), extent=(1, 1, 1), topology=("P", "P", "B"))
    f = Field(g, "ccc")
    f.set(np.random.default_rng(1).random((4, 4, 4)))
    fill_halo_regions(f)
    H = 3
    assert np.array_equal(f.data[0:H, :, H - 1], f.data[4:4 + H, :, H - 1])       # west halo of the bottom halo plane
    assert (f.data[:, :, 0:H - 1] == 0).all()                                      # halo planes 2..H never written
    # value / gradient BCs (fill_halo_regions_value_gradient.jl:7-119)
    f2 = Field(g, "ccc", bcs={"bottom": BC("gradient", 0.5), "top": BC("value", 2.0)})
    f2.set(np.ones((4, 4, 4)))
    fill_halo_regions(f2)
    dz = 0.25
    assert np.allclose(f2.data[H:H + 4, H:H + 4, H - 1], 1.0 - 0.5 * dz)
    assert np.allclose(f2.data[H:H + 4, H:H + 4, H + 4], 1.0 + (2.0 - 1.0) / (dz / 2) * dz)


# ----------------------------------------------------------------------------- closures (§8c-3)
@pytest.mark.parametrize("FT", [np.float32, np.float64])
def test_constant_isotropic_diffusivity_fluxdiv(FT):
    # test/test_turbulence_closures.jl:36-66 — exact equality
    nu, kappa = FT(0.3), FT(0.7)
    g = Grid(FT, size=(3, 1, 4), extent=(3, 1, 4), topology=("P", "P", "B"))
    u, v, w = Field(g, "fcc"), Field(g, "cfc"), Field(g, "ccf")
    T = Field(g, "ccc")
    for k in range(4):
        u.interior[:, 0, k] = [0, -0.5, 0]
        v.interior[:, 0, k] = [0, -2, 0]
        w.interior[:, 0, k] = [0, -3, 0]
        T.interior[:, 0, k] = [0, -1, 0]
    for f in (u, v, w, T):
        fill_halo_regions(f)
    ctx = Ctx(g, (2, 2), (1, 1), (3, 3))
    U = (u, v, w)
    assert clo.div_q(ctx, kappa, T)[0, 0, 0] == -2 * kappa
    assert clo.div_tau(ctx, nu, U, 0)[0, 0, 0] == -2 * nu
    assert clo.div_tau(ctx, nu, U, 1)[0, 0, 0] == -4 * nu
    assert clo.div_tau(ctx, nu, U, 2)[0, 0, 0] == -6 * nu


# ----------------------------------------------------------------------------- Poisson (§8c-2)
TOPOS = [(a, b, c) for a in "PB" for b in "PB" for c in "PB"]


def _laplacian(g, phi_field):
    ctx = Ctx(g, (1, g.Nx), (1, g.Ny), (1, g.Nz))
    p = ctx.field(phi_field)
    tot = None
    for d in range(3):
        A = ctx.area(d, "f" if d == 2 else "c")
        t = dC(ctx, (lambda dd, A: (lambda o: A(o) * ddF(ctx, p, dd)(o)))(d, A), d)(O)
        tot = t if tot is None else tot + t
    return ctx.rvol("c")(O) * tot


def _divergence_free_poisson_solution(g, seed=0):
    # test/dependencies_for_poisson_solvers.jl:14-41,111-129
    rng = np.random.default_rng(seed)
    m = OracleModel(g)
    for f in m.U:
        f.set(rng.random(f.interior.shape))
        fill_halo_regions(f)
    R = m.divergence()
    m.compute_pressure_correction(1.0)
    lap = _laplacian(g, m.pNHS)
    # Julia's isapprox(A, B): norm(A-B) <= sqrt(eps) * max(norm(A), norm(B))
    tol = np.sqrt(np.finfo(g.FT).eps)
    return np.linalg.norm((lap - R).ravel()) <= tol * max(np.linalg.norm(lap.ravel()), np.linalg.norm(R.ravel()))


@pytest.mark.parametrize("topo", TOPOS)
@pytest.mark.parametrize("N", [7, 16])
def test_divergence_free_poisson_solution_square(topo, N):
    # test/test_poisson_solvers.jl:58-75
    for size in [(N, N, N), (1, N, N), (N, 1, N), (N, N, 1)]:
        assert _divergence_free_poisson_solution(Grid(np.float64, size=size, extent=(1, 1, 1), topology=topo))


@pytest.mark.parametrize("topo2", [("P", "P", "F"), ("P", "B", "F"), ("B", "B", "F"),
                                   ("P", "F", "B"), ("F", "P", "B"), ("F", "B", "B")])
def test_divergence_free_poisson_solution_2d(topo2):
    # test/test_poisson_solvers.jl:66-67 (two_dimensional_topologies)
    assert _divergence_free_poisson_solution(Grid(np.float64, size=(7, 16), extent=(1, 1), topology=topo2))


@pytest.mark.parametrize("topo", TOPOS)
def test_divergence_free_poisson_solution_rectangular(topo):
    # :78-85  even and prime sizes
    for Nx in (11, 16):
        for Nz in (11, 16):
            assert _divergence_free_poisson_solution(Grid(np.float64, size=(Nx, 16 if Nx == 11 else 11, Nz),
                                                          extent=(1, 1, 1), topology=topo))


def test_divergence_free_poisson_solution_float32():
    # :88-93
    assert _divergence_free_poisson_solution(Grid(np.float32, size=(16, 16, 16), extent=(1, 1, 1), topology=("P", "B", "B")))
    assert _divergence_free_poisson_solution(Grid(np.float32, size=(7, 11, 13), extent=(1, 1, 1), topology=("B", "B", "P")))


# ----------------------------------------------------------------------------- stretched grids (§8f-1)
def test_vertically_stretched_grid_spacings():
    # test/test_grids.jl:440-485 (test_rectilinear_grid_correct_spacings, the z part) + halo continuation of
    # generate_coordinate for a Bounded coordinate (src/Grids/grid_generation.jl:15-18,56-63)
    for FT in (np.float32, np.float64):
        N, S = 16, 3
        zf = lambda k: np.tanh(S * (2 * (k - 1) / N - 1)) / np.tanh(S)
        g = Grid(FT, size=(N, N, N), x=(0, N), y=(0, N), z=zf, topology=("P", "P", "B"))
        k = np.arange(1, N + 1)
        zc = (zf(k) + zf(k + 1)) / 2
        assert g.stretched and g._dzc.dtype == FT and g._dzf.dtype == FT
        assert np.allclose(g.nodes(2, "f"), zf(np.arange(1, N + 2)))
        assert np.allclose(g.nodes(2, "c"), zc)
        assert np.allclose(g.dz_at("c", k), zf(k + 1) - zf(k), rtol=1e-5 if FT is np.float32 else 1e-12)
        assert np.allclose(g.dz_at("f", k[1:]), zc[1:] - zc[:-1], rtol=1e-5 if FT is np.float32 else 1e-12)
        H = g.H[2]
        assert len(g._dzf) == N + 2 * H + 1 and len(g._dzc) == N + 2 * H          # test_grids.jl:652-653
        # halo cells continue with the first / last interior spacing; the wall face spacing equals the adjacent cell's
        rt = 1e-4 if FT is np.float32 else 1e-12      # (face positions are rounded to FT before differencing)
        assert np.allclose(g.dz_at("c", np.arange(1 - H, 1)), g.dz_at("c", 1), rtol=rt)
        assert np.allclose(g.dz_at("c", np.arange(N + 1, N + H + 1)), g.dz_at("c", N), rtol=rt)
        assert np.isclose(g.dz_at("f", 1), g.dz_at("c", 1), rtol=rt) and np.isclose(g.dz_at("f", N + 1), g.dz_at("c", N), rtol=rt)
    with pytest.raises(AssertionError):
        Grid(np.float64, size=(4, 4, 4), x=(0, 1), y=(0, 1), z=[0, 0.3, 0.2, 0.6, 1.0], topology=("P", "P", "B"))


def test_regularly_spaced_faces_reproduce_the_regular_grid():
    # test/test_grids.jl:638-663: a "stretched" grid built from regularly spaced faces has the regular grid's metrics;
    # here additionally: the FourierTridiagonal and the FFT-based (DCT) solvers then give the same model trajectory
    N = (8, 6, 8)
    res = []
    for z in ((-0.5, 0.0), [float(v) for v in np.linspace(-0.5, 0.0, 9)]):
        g = Grid(np.float64, size=N, x=(0, 1), y=(0, 1), z=z, topology=("P", "P", "B"))
        m = OracleModel(g, advection=adv.Centered(np.float64, 2), tracers=("T", "S"), buoyancy=clo.SeawaterBuoyancy(),
                        closure=clo.ScalarDiffusivity(1e-3, 1e-3), coriolis_f=1e-2)
        rng = np.random.default_rng(3)
        ic = {n: rng.uniform(-1, 1, m.fields[n].interior.shape) for n in ("u", "v", "w")}
        ic["T"] = 20 + 0.01 * rng.standard_normal(N)
        ic["S"] = 35 + 0.01 * rng.standard_normal(N)
        m.set(**ic)
        for _ in range(3):
            m.time_step(0.01)
        res.append(m)
    assert np.all(res[1].grid.dz_at("c", np.arange(1, 9)) == res[0].grid.dz)
    for n in res[0].fields:
        a, b = res[0].fields[n].interior, res[1].fields[n].interior
        assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max(), n
    assert np.abs(res[0].pNHS.interior - res[1].pNHS.interior).max() <= 1e-11 * np.abs(res[0].pNHS.interior).max()


FACES_EVEN = [1, 2, 4, 7, 11, 16, 22, 29, 37]
FACES_ODD = [1, 2, 4, 7, 11, 16, 22, 29, 37, 51]
VS_TOPOS = [("P", "P", "B"), ("P", "B", "B"), ("B", "P", "B"), ("B", "B", "B"), ("F", "B", "B"), ("F", "P", "B"),
            ("B", "F", "B"), ("P", "F", "B")]


def _stretched_poisson_solver_correct_answer(FT, topo, N1, N2, faces, seed=0):
    # stretched_poisson_solver_correct_answer, test/dependencies_for_poisson_solvers.jl:231-256 (stretched_axis = 3):
    # R = divergence of a random velocity; solve; ∇²ϕ ≈ R
    kw = {}
    size = []
    if topo[0] != "F":
        kw["x"] = (0, 1)
        size.append(N1)
    if topo[1] != "F":
        kw["y"] = (0, 1)
        size.append(N2)
    size.append(len(faces) - 1)
    g = Grid(FT, size=tuple(size), z=[float(f) for f in faces], topology=topo, **kw)
    rng = np.random.default_rng(seed)
    m = OracleModel(g)
    for f in m.U:
        f.set(rng.random(f.interior.shape))
        fill_halo_regions(f)
    R = m.divergence()
    m.compute_pressure_correction(1.0)          # Δzᶜᶜᶜ·R -> FourierTridiagonalPoissonSolver -> pNHS (+ halos)
    lap = _laplacian(g, m.pNHS)
    tol = np.sqrt(np.finfo(g.FT).eps)
    ok = np.linalg.norm((lap - R).ravel()) <= tol * max(np.linalg.norm(lap.ravel()), np.linalg.norm(R.ravel()))
    return ok and abs(float(m.pNHS.interior.mean())) <= 100 * np.finfo(g.FT).eps * max(1.0, np.abs(m.pNHS.interior).max())


@pytest.mark.parametrize("topo", VS_TOPOS)
def test_stretched_poisson_solver_correct_answer(topo):
    # test/test_poisson_solvers_stretched_grids.jl:12-52 with stretched_axis = 3 (the only stretched axis on this path)
    assert _stretched_poisson_solver_correct_answer(np.float64, topo, 4, 5, range(1, 5))
    assert _stretched_poisson_solver_correct_answer(np.float64, topo, 8, 8, range(1, 9))
    assert _stretched_poisson_solver_correct_answer(np.float64, topo, 7, 7, range(1, 8))
    assert _stretched_poisson_solver_correct_answer(np.float32, topo, 8, 8, range(1, 9))
    for faces in (FACES_EVEN, FACES_ODD):
        for N1, N2 in ((8, 8), (16, 8), (8, 16), (8, 11), (5, 8), (7, 13)):
            assert _stretched_poisson_solver_correct_answer(np.float64, topo, N1, N2, faces), (N1, N2)


def test_stretched_model_stays_incompressible_and_conserves_tracer():
    # test/test_time_stepping.jl:124-199 on a vertically stretched grid (the FourierTridiagonal path of time_step!)
    faces = [float(-0.5 + 0.5 * (f - 1) / 36.0) for f in FACES_EVEN]
    g = Grid(np.float64, size=(8, 6, 8), x=(0, 1), y=(0, 1), z=faces, topology=("P", "B", "B"))
    m = OracleModel(g, advection=adv.WENO(np.float64, 5), tracers=("T", "S"), buoyancy=clo.SeawaterBuoyancy(),
                    closure=clo.AnisotropicMinimumDissipation())
    rng = np.random.default_rng(5)
    ic = {n: rng.uniform(-1, 1, m.fields[n].interior.shape) for n in ("u", "v", "w")}
    ic["T"] = 20 + 0.01 * rng.standard_normal((8, 6, 8))
    ic["S"] = 35 + 0.01 * rng.standard_normal((8, 6, 8))
    m.set(**ic)
    dzc = g.dz_at("c", np.arange(1, 9))
    T0 = float((m.tracers["T"].interior * dzc).sum())
    for _ in range(5):
        m.time_step(0.005)
    assert np.abs(m.divergence()).max() < 5e-8
    assert abs(float((m.tracers["T"].interior * dzc).sum()) - T0) <= 1e-12 * abs(T0)      # volume-weighted: flux form conserves


def _analytic_error(N, topo, mode):
    # test/dependencies_for_poisson_solvers.jl:135-158
    g = Grid(np.float64, size=(N, N, N), x=(0, 2 * np.pi), y=(0, 2 * np.pi), z=(0, 2 * np.pi), topology=topo)
    m = OracleModel(g)
    xs = [g.nodes(d, "c") for d in range(3)]
    psi = lambda t, x: np.cos(mode * x / 2) if t == "B" else np.cos(mode * x)
    k2 = lambda t: (mode / 2) ** 2 if t == "B" else mode ** 2
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    Psi = psi(topo[0], X) * psi(topo[1], Y) * psi(topo[2], Z)
    f = -(k2(topo[0]) + k2(topo[1]) + k2(topo[2])) * Psi
    phi = m.solve_poisson(f)
    return np.mean(np.abs(phi - Psi))


@pytest.mark.parametrize("topo", TOPOS)
def test_poisson_solver_convergence(topo):
    # test/test_poisson_solvers.jl:100-106 : rate ≈ 2 (rtol 5e-3)
    e1, e2 = _analytic_error(64, topo, 1), _analytic_error(128, topo, 1)
    rate = np.log(e1 / e2) / np.log(128 / 64)
    assert abs(rate - 2) <= 5e-3 * 2
    e1, e2 = _analytic_error(67, topo, 2), _analytic_error(131, topo, 2)
    rate = np.log(e1 / e2) / np.log(131 / 67)
    assert abs(rate - 2) <= 5e-3 * 2


def test_poisson_eigenvalues():
    # src/Solvers/poisson_eigenvalues.jl:8-31
    lam = poisson_eigenvalues(8, 2.0, "P")
    assert lam[0] == 0 and np.isclose(lam[4], (2 / 0.25) ** 2)
    lam = poisson_eigenvalues(8, 2.0, "B")
    assert lam[0] == 0 and np.isclose(lam[1], (2 * np.sin(np.pi / 16) / 0.25) ** 2)
    assert (poisson_eigenvalues(1, 1.0, "F") == 0).all()


# ----------------------------------------------------------------------------- time stepping (§8c-4)
@pytest.mark.parametrize("FT", [np.float32, np.float64])
@pytest.mark.parametrize("ts", ["RungeKutta3", "QuasiAdamsBashforth2"])
@pytest.mark.parametrize("Nt", [1, 10])
def test_incompressible_in_time(FT, ts, Nt):
    # test/test_time_stepping.jl:124-158,432-460 (regular grid; Nt=100 is in the slow test below)
    g = Grid(FT, size=(32, 32, 32), x=(0, 1), y=(0, 1), z=(-1, 1), topology=("P", "P", "B"))
    m = OracleModel(g, timestepper=ts, buoyancy=clo.SeawaterBuoyancy(), tracers=("T", "S"))
    m.tracers["T"].interior[7:24, 7:24, 7:24] += FT(0.01)
    m.update_state()
    for _ in range(Nt):
        m.time_step(0.05)
    assert np.abs(m.divergence()).max() < 5e-8
    assert np.abs(m.w.interior).max() > 0


def test_ab2_first_step_is_euler_and_clock():
    # quasi_adams_bashforth_2.jl:88-96 ; clock.jl:128-143 ; runge_kutta_3.jl:107-161
    g = Grid(np.float64, size=(8, 8, 8), extent=(1, 1, 1), topology=("P", "P", "B"))
    m = OracleModel(g, timestepper="QuasiAdamsBashforth2")
    m.time_step(0.1)
    assert m.clock.iteration == 1 and m.clock.time == 0.1 and m.clock.last_dt == 0.1
    m = OracleModel(g, timestepper="RungeKutta3")
    m.time_step(0.3)
    assert m.clock.iteration == 1 and m.clock.stage == 1 and m.clock.last_dt == 0.3
    assert abs(m.clock.time - 0.3) < 1e-15


def test_tracer_conserved_in_channel():
    # test/test_time_stepping.jl:165-199 with an isotropic ScalarDiffusivity in place of the H/V pair (out of scope)
    g = Grid(np.float64, size=(16, 32, 16), extent=(160e3, 320e3, 1024), topology=("P", "B", "B"))
    m = OracleModel(g, closure=clo.ScalarDiffusivity(nu=1e-2, kappa=1e-2), buoyancy=clo.SeawaterBuoyancy(), tracers=("T", "S"))
    rng = np.random.default_rng(3)
    noise = rng.random((16, 32, 16))
    X, Y, Z = np.meshgrid(g.nodes(0, "c"), g.nodes(1, "c"), g.nodes(2, "c"), indexing="ij")
    m.set(T=10 + 1e-4 * Y + 5e-3 * Z + 1e-4 * noise)
    T0 = m.tracers["T"].interior.mean()
    for _ in range(10):
        m.time_step(600)
    assert abs(m.tracers["T"].interior.mean() - T0) <= 16 * 32 * 16 * np.finfo(np.float64).eps


# ----------------------------------------------------------------------------- dynamics (§8c-5)
def test_constant_stays_constant_under_diffusion():
    # test/test_dynamics.jl:17-32
    g = Grid(np.float64, size=(8, 8, 8), extent=(1, 1, 1), topology=("P", "P", "B"))
    m = OracleModel(g, closure=clo.ScalarDiffusivity(nu=1.0, kappa=1.0), tracers=("c",))
    m.set(c=np.full((8, 8, 8), np.pi), enforce_incompressibility=False)
    for _ in range(3):
        m.time_step(1e-3)
    assert np.allclose(m.tracers["c"].interior, np.pi, rtol=0, atol=1e-14)


@pytest.mark.parametrize("ts", ["RungeKutta3", "QuasiAdamsBashforth2"])
@pytest.mark.parametrize("dim", [0, 1, 2])
def test_diffusion_cosine(ts, dim):
    # test/test_dynamics.jl:65-87,497-530 : cos(2ξ) on [0, π/2], Bounded in the diffusing direction
    N, L = 128, np.pi / 2
    size = [1, 1, 1]
    size[dim] = N
    topo = ["P", "P", "P"]
    topo[dim] = "B"
    ext = [(0, 1)] * 3
    ext[dim] = (0, L)
    g = Grid(np.float64, size=tuple(size), x=ext[0], y=ext[1], z=ext[2], topology=tuple(topo))
    names = [n for n, d in (("u", 0), ("v", 1), ("w", 2)) if d != dim] + ["c"]
    for name in names:
        m = OracleModel(g, closure=clo.ScalarDiffusivity(nu=1.0, kappa=1.0), tracers=("c",), timestepper=ts)
        xi = g.nodes(dim, "c")
        shape = [1, 1, 1]
        shape[dim] = N
        f = m.fields[name]
        f.interior[...] = np.cos(2 * xi).reshape(shape)
        m.update_state()
        dt = 1e-6 * float(g.L[2]) ** 2
        for _ in range(5):
            m.time_step(dt)
        exact = np.exp(-4 * m.clock.time) * np.cos(2 * xi).reshape(shape)
        assert np.allclose(f.interior, exact, atol=1e-6, rtol=1e-6), (name, dim)


def test_taylor_green_vortex():
    # test/test_dynamics.jl:216-261,703-711 : AB2, N=64, 10 steps, max relative error < 5e-6
    N = 64
    g = Grid(np.float64, size=(N, N, 2), extent=(1, 1, 1), topology=("P", "P", "B"))
    m = OracleModel(g, closure=clo.ScalarDiffusivity(nu=1.0, kappa=0.0), timestepper="QuasiAdamsBashforth2")
    u = lambda x, y, z, t: -np.sin(2 * np.pi * y) * np.exp(-4 * np.pi ** 2 * t) + 0 * x
    v = lambda x, y, z, t: np.sin(2 * np.pi * x) * np.exp(-4 * np.pi ** 2 * t) + 0 * y
    m.set(u=lambda x, y, z: u(x, y, z, 0), v=lambda x, y, z: v(x, y, z, 0))
    dt = (1 / (10 * np.pi)) * (1 / N) ** 2
    for _ in range(10):
        m.time_step(dt)
    t = m.clock.time
    XF, YC, ZC = np.meshgrid(g.nodes(0, "f"), g.nodes(1, "c"), g.nodes(2, "c"), indexing="ij")
    XC, YF, _ = np.meshgrid(g.nodes(0, "c"), g.nodes(1, "f"), g.nodes(2, "c"), indexing="ij")
    with np.errstate(divide="ignore", invalid="ignore"):
        ue, ve = u(XF, YC, ZC, t), v(XC, YF, ZC, t)
        mu, mv = np.abs(ue) > 1e-12, np.abs(ve) > 1e-12   # nodes where the exact solution is ~0 give 0/0 in the reference too
        uerr = np.abs((m.u.interior - ue) / ue)[mu].max()
        verr = np.abs((m.v.interior - ve) / ve)[mv].max()
    assert uerr < 5e-6 and verr < 5e-6


@pytest.mark.parametrize("flat_y", [False, True])
def test_internal_wave(flat_y):
    # test/test_internal_wave_dynamics.jl:4-87 ; test/test_dynamics.jl:634-701
    L = 2 * np.pi
    nu = 1e-9
    z0, delta, a0, mm, kk, f, N2 = -L / 3, L / 20, 1e-3, 16, 1, 0.2, 1.0
    sigma = np.sqrt((N2 * kk ** 2 + f ** 2 * mm ** 2) / (kk ** 2 + mm ** 2))
    dt = 0.01 / sigma
    cg = mm * sigma / (kk ** 2 + mm ** 2) * (f ** 2 / sigma ** 2 - 1)
    Uc = a0 * kk * sigma / (sigma ** 2 - f ** 2)
    Vc = a0 * kk * f / (sigma ** 2 - f ** 2)
    Wc = a0 * mm * sigma / (sigma ** 2 - N2)
    Bc = a0 * mm * N2 / (sigma ** 2 - N2)
    a = lambda x, z, t: np.exp(-(z - cg * t - z0) ** 2 / (2 * delta) ** 2)
    u = lambda x, z, t: a(x, z, t) * Uc * np.cos(kk * x + mm * z - sigma * t)
    v = lambda x, z, t: a(x, z, t) * Vc * np.sin(kk * x + mm * z - sigma * t)
    w = lambda x, z, t: a(x, z, t) * Wc * np.cos(kk * x + mm * z - sigma * t)
    b = lambda x, z, t: a(x, z, t) * Bc * np.sin(kk * x + mm * z - sigma * t) + N2 * z
    if flat_y:
        g = Grid(np.float64, size=(128, 128), x=(0, L), z=(-L, 0), topology=("P", "F", "B"))
        wrap = lambda fn: (lambda x, z: fn(x, z, 0))
    else:
        g = Grid(np.float64, size=(128, 1, 128), x=(0, L), y=(0, L), z=(-L, 0), topology=("P", "P", "B"))
        wrap = lambda fn: (lambda x, y, z: fn(x, z, 0))
    m = OracleModel(g, closure=clo.ScalarDiffusivity(nu=nu, kappa=nu), buoyancy=clo.BuoyancyTracer(),
                    tracers=("b",), coriolis_f=f)
    m.set(u=wrap(u), v=wrap(v), w=wrap(w), b=wrap(b))
    for _ in range(10):
        m.time_step(dt)
    X, _, Z = np.meshgrid(g.nodes(0, "f"), np.zeros(1), g.nodes(2, "c"), indexing="ij")
    ue = u(X, Z, m.clock.time)
    rel = np.mean((m.u.interior - ue) ** 2) / np.mean(ue ** 2)
    assert rel < 1e-4


def test_flux_bc_budget():
    # test/test_boundary_conditions_integration.jl:28-52,309-360 : <c> ≈ flux * t / L  for a top/bottom flux
    g = Grid(np.float64, size=(4, 4, 16), extent=(1, 1, 0.75), topology=("P", "P", "B"))
    flux = np.pi
    for side, sign in (("bottom", 1.0), ("top", -1.0)):
        m = OracleModel(g, tracers=("c",), closure=clo.ScalarDiffusivity(nu=1.0, kappa=1.0),
                        boundary_conditions={"c": {side: BC("flux", flux)}})
        dt = 1e-6
        for _ in range(10):
            m.time_step(dt)
        mean = m.tracers["c"].interior.mean()
        assert np.isclose(mean, sign * flux * m.clock.time / 0.75, rtol=1e-10)


def test_weno_beta_conditioning():
    """DESIGN.md §2 numerical note.  The reference's expanded smoothness indicators (weno_interpolants.jl:204-216) cancel
    catastrophically for fields with a large mean (S ≈ 35 ± 0.01): evaluated in Float64 (the oracle does exactly that) one
    reconstruction carries ~1e-11 relative round-off, while the algebraically identical difference form used by the CUDA
    kernel (13/4 (second difference)² + 3/4 (one-sided difference)²) is accurate to ~1e-14.  This bounds how closely ANY two
    Float64 evaluations of the reference formula can agree on T and S, and explains the ~1e-12 GPU–oracle gap there."""
    from oracle import advection as adv
    rng = np.random.default_rng(0)
    n = 20000
    S = [35 + 0.01 * rng.standard_normal(n) for _ in range(6)]
    sch = adv.WENO(np.float64, 5)
    val = adv._weno_value(sch, S, np.ones(n, bool))
    L = np.longdouble
    q = [s.astype(L) for s in S[:5]]

    def beta(a, b, c, C):
        return a * (C[0] * a + C[1] * b + C[2] * c) + b * (C[3] * b + C[4] * c) + c * c * C[5]

    b0 = beta(q[2], q[3], q[4], (10, -31, 11, 25, -19, 4))
    b1 = beta(q[1], q[2], q[3], (4, -13, 5, 13, -13, 4))
    b2 = beta(q[0], q[1], q[2], (4, -19, 11, 25, -31, 10))
    eps = L(np.float32(1e-8))
    tau = abs(b0 - b2)
    a = [L(3) / 10 * (1 + (tau / (b0 + eps)) ** 2), L(3) / 5 * (1 + (tau / (b1 + eps)) ** 2), L(1) / 10 * (1 + (tau / (b2 + eps)) ** 2)]
    p0 = q[2] / 3 + q[3] * 5 / 6 - q[4] / 6
    p1 = -q[1] / 6 + q[2] * 5 / 6 + q[3] / 3
    p2 = q[0] / 3 - q[1] * 7 / 6 + q[2] * 11 / 6
    truth = (a[0] * p0 + a[1] * p1 + a[2] * p2) / (a[0] + a[1] + a[2])
    err_expanded = float(np.abs(val - truth).max() / 35)
    # the kernel's arrangement (csrc/oc_march.h: weno5_value_c), in Float64
    q = S[:5]
    e1, e2, e3, e4 = q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3]
    d0, d1, d2 = e4 - e3, e3 - e2, e2 - e1
    g0, g1, g2 = e4 - 3 * e3, e2 + e3, 3 * e2 - e1
    es = float(np.float32(1e-8)) * 4 / 3
    B0, B1, B2 = 13 / 3 * d0 * d0 + g0 * g0 + es, 13 / 3 * d1 * d1 + g1 * g1 + es, 13 / 3 * d2 * d2 + g2 * g2 + es
    t2 = (B0 - B2) ** 2
    w0, w1, w2 = (B0 * B0 + t2) * (B1 * B2) ** 2, (B1 * B1 + t2) * (B0 * B2) ** 2, (B2 * B2 + t2) * (B0 * B1) ** 2
    den = 6 * w2 + 36 * w1 + 18 * w0
    num = w2 * (5 * e2 - 2 * e1) + w1 * (12 * e3 + 6 * e2) + w0 * (12 * e3 - 3 * e4)
    mine = q[2] + num / den
    err_diff = float(np.abs(mine - truth).max() / 35)
    assert err_diff < 1e-13                 # the kernel's form is accurate
    assert 1e-13 < err_expanded < 1e-9      # the reference's form loses ~5 digits more on such data


# ----------------------------------------------------------------------------- Smagorinsky / SmagorinskyLilly (§8f item 3)
# The reference holds no known-answer numbers for this closure (its regression files are downloaded at test time), so the
# restatement is pinned by the closed forms of its own docstrings (smagorinsky.jl:52-67, lilly_coefficient.jl:13-16,114-126):
# νₑ = (Cˢ Δᶠ)² √(2Σ²) ς,  ς = √(1 − Cb N²⁺ / Σ²),  Δᶠ = ∛(Δx Δy Δz),  κₑ = νₑ / Pr — evaluated for flows with uniform strain.
def _smag_fields(g, ufun=None, vfun=None, wfun=None, bfun=None):
    u, v, w, b = Field(g, "fcc"), Field(g, "cfc"), Field(g, "ccf"), Field(g, "ccc", None, "b")
    for f, fun in ((u, ufun), (v, vfun), (w, wfun), (b, bfun)):
        if fun is not None:
            # nodes of every non-Flat dimension, including the halos, so that the strain is uniform up to the walls
            X = [g.x0[d] + (np.arange(f.data.shape[d]) - g.H[d] + (0.5 if f.loc[d] == "c" else 0.0)) * float(g.D[d]) for d in range(3)]
            f.data[...] = fun(X[0][:, None, None], X[1][None, :, None], X[2][None, None, :])
    return (u, v, w), b


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_smagorinsky_uniform_strain_closed_forms(FT):
    g = Grid(FT, size=(6, 5, 4), extent=(3.0, 2.0, 1.0), topology=("B", "B", "B"), halo=(2, 2, 2))
    C, s, a = 0.16, 0.7, -0.4
    Df2 = float(np.cbrt(float(g.D[0]) * float(g.D[1]) * float(g.D[2]))) ** 2
    ctx = Ctx(g, (1, g.Nx), (1, g.Ny), (1, g.Nz))
    nu, kap = Field(g, "ccc"), {"b": Field(g, "ccc")}
    tol = 50 * np.finfo(FT).eps
    # simple shear u = s z: Σ13 = s/2, Σ² = s²/2, νₑ = (CΔ)² |s|
    U, b = _smag_fields(g, ufun=lambda x, y, z: s * z + 0 * x + 0 * y)
    clo.compute_smagorinsky(ctx, clo.Smagorinsky(C, Pr=2.0), U, {"b": b}, None, nu, kap)
    assert np.allclose(nu.interior, C ** 2 * Df2 * abs(s), rtol=tol, atol=0)
    assert np.allclose(kap["b"].interior, C ** 2 * Df2 * abs(s) / 2.0, rtol=tol, atol=0)
    # plane strain u = a x, v = -a y: Σ² = 2a², νₑ = (CΔ)² 2|a|
    U, b = _smag_fields(g, ufun=lambda x, y, z: a * x + 0 * y + 0 * z, vfun=lambda x, y, z: -a * y + 0 * x + 0 * z)
    clo.compute_smagorinsky(ctx, clo.Smagorinsky(C), U, {"b": b}, None, nu, kap)
    assert np.allclose(nu.interior, C ** 2 * Df2 * 2 * abs(a), rtol=tol, atol=0)
    # horizontal shear v = s x and w = s y: Σ12 = Σ23 = s/2, Σ² = s², νₑ = (CΔ)² √2 |s|
    U, b = _smag_fields(g, vfun=lambda x, y, z: s * x + 0 * y + 0 * z, wfun=lambda x, y, z: s * y + 0 * x + 0 * z)
    clo.compute_smagorinsky(ctx, clo.Smagorinsky(C), U, {"b": b}, None, nu, kap)
    assert np.allclose(nu.interior, C ** 2 * Df2 * np.sqrt(2.0) * abs(s), rtol=tol, atol=0)
    # fluid at rest: νₑ = 0 exactly (√0), also through the Lilly branch (Σ² == 0 -> ς = 0)
    U, b = _smag_fields(g, bfun=lambda x, y, z: 0.3 * z + 0 * x + 0 * y)
    clo.compute_smagorinsky(ctx, clo.SmagorinskyLilly(C, 1.0), U, {"b": b}, clo.BuoyancyTracer(), nu, kap)
    assert np.all(nu.interior == 0)


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_amd_buoyancy_modification_closed_forms(FT):
    """AnisotropicMinimumDissipation(; Cb)  anisotropic_minimum_dissipation.jl:62-68: νₑ = max(0, −Cν δ² (r − Cb ζ) / q) with
    Cb ζ = Cb_norm_wᵢ_bᵢᶜᶜᶜ / Δᶠz (:168-172, :310-323).  The reference has no numbers for the (unvalidated, :137) modification; the
    restatement is pinned by closed forms for uniform gradients:
      w = a x, b = s x:  q = (Δᶠx/Δᶠz)² a², r = 0, ζ = a s Δᶠx²/Δᶠz²   ->  νₑ = Cν δ² Cb s / a      (same with y)
      w = a z, b = N² z:  q = a², r = a³, ζ = a N²                       ->  νₑ = −Cν δ² (a − Cb N² / a)"""
    g = Grid(FT, size=(6, 5, 4), extent=(3.0, 2.0, 1.0), topology=("B", "B", "B"), halo=(2, 2, 2))
    Cnu, a, s = 1.0 / 3.0, 0.4, 0.7
    Df = [2 * float(g.D[d]) for d in range(3)]
    d2 = 3.0 / sum(1.0 / x ** 2 for x in Df)
    ctx = Ctx(g, (1, g.Nx), (1, g.Ny), (1, g.Nz))
    nu, kap = Field(g, "ccc"), {"b": Field(g, "ccc")}
    tol = 100 * np.finfo(FT).eps
    bt = clo.BuoyancyTracer()
    for wfun, bfun in ((lambda x, y, z: a * x + 0 * y + 0 * z, lambda x, y, z: s * x + 0 * y + 0 * z),
                       (lambda x, y, z: a * y + 0 * x + 0 * z, lambda x, y, z: s * y + 0 * x + 0 * z)):
        U, b = _smag_fields(g, wfun=wfun, bfun=bfun)
        for Cb in (1.0, 2.5):
            clo.compute_amd(ctx, clo.AnisotropicMinimumDissipation(Cb=Cb), U, {"b": b}, nu, kap, bt)
            assert np.allclose(nu.interior, Cnu * d2 * Cb * s / a, rtol=tol, atol=0), Cb
        # an unstable alignment (Cb ζ < 0) is clipped; Cb = nothing and "no buoyancy" leave r = 0 -> νₑ = 0
        clo.compute_amd(ctx, clo.AnisotropicMinimumDissipation(Cb=-1.0), U, {"b": b}, nu, kap, bt)
        assert np.all(nu.interior == 0)
        clo.compute_amd(ctx, clo.AnisotropicMinimumDissipation(), U, {"b": b}, nu, kap, bt)
        assert np.all(nu.interior == 0)
        clo.compute_amd(ctx, clo.AnisotropicMinimumDissipation(Cb=1.0), U, {"b": b}, nu, kap, None)
        assert np.all(nu.interior == 0)
    for aa, N2, Cb in ((-0.4, 0.05, 1.0), (-0.4, 0.05, 0.0), (-0.4, -0.2, 1.0), (0.4, 0.3, 1.0), (0.4, 0.05, 1.0)):
        U, b = _smag_fields(g, wfun=lambda x, y, z: aa * z + 0 * x + 0 * y, bfun=lambda x, y, z: N2 * z + 0 * x + 0 * y)
        clo.compute_amd(ctx, clo.AnisotropicMinimumDissipation(Cb=Cb), U, {"b": b}, nu, kap, bt)
        want = max(0.0, -Cnu * d2 * (aa - Cb * N2 / aa))
        assert np.allclose(nu.interior, want, rtol=tol, atol=10 * np.finfo(FT).eps * Cnu * d2), (aa, N2, Cb)
    # SeawaterBuoyancy: b = g (α T − β S)  (seawater_buoyancy.jl:203-207, linear_equation_of_state.jl:72-80)
    sw = clo.SeawaterBuoyancy()
    U, T = _smag_fields(g, wfun=lambda x, y, z: a * x + 0 * y + 0 * z, bfun=lambda x, y, z: s * x + 0 * y + 0 * z)
    _, S = _smag_fields(g, bfun=lambda x, y, z: -2 * s * x + 0 * y + 0 * z)
    kap2 = {"T": Field(g, "ccc"), "S": Field(g, "ccc")}
    clo.compute_amd(ctx, clo.AnisotropicMinimumDissipation(Cb=1.0), U, {"T": T, "S": S}, nu, kap2, sw)
    sb = float(sw.g) * (float(sw.alpha) * s + float(sw.beta) * 2 * s)
    assert np.allclose(nu.interior, Cnu * d2 * sb / a, rtol=tol, atol=0)


def test_smagorinsky_lilly_stability_function():
    FT = np.float64
    g = Grid(FT, size=(4, 4, 6), extent=(2.0, 2.0, 3.0), topology=("B", "B", "B"), halo=(2, 2, 2))
    C, s = 0.23, 0.5
    Df2 = float(np.cbrt(float(g.D[0]) * float(g.D[1]) * float(g.D[2]))) ** 2
    ctx = Ctx(g, (1, g.Nx), (1, g.Ny), (1, g.Nz))
    nu, kap = Field(g, "ccc"), {"b": Field(g, "ccc")}
    S2 = s * s / 2
    for N2, Cb in ((0.05, 1.0), (0.05, 2.0), (-0.3, 1.0), (1.0, 1.0), (0.05, 0.0)):
        U, b = _smag_fields(g, ufun=lambda x, y, z: s * z + 0 * x + 0 * y, bfun=lambda x, y, z: N2 * z + 0 * x + 0 * y)
        clo.compute_smagorinsky(ctx, clo.SmagorinskyLilly(C, Cb), U, {"b": b}, clo.BuoyancyTracer(), nu, kap)
        sig = np.sqrt(1 - min(1.0, Cb * max(0.0, N2) / S2))
        assert np.allclose(nu.interior, sig * C ** 2 * Df2 * abs(s), rtol=1e-13, atol=0), (N2, Cb)
    # the three branches of stability(): lilly_coefficient.jl:122-126
    f = clo.lilly_stability
    assert f(FT, FT(-1.0), FT(2.0), FT(1.0)) == 1.0            # unstable stratification: no reduction
    assert f(FT, FT(4.0), FT(2.0), FT(1.0)) == 0.0             # Ri-number cut-off
    assert f(FT, FT(1.0), FT(0.0), FT(1.0)) == 0.0             # Σ² == 0
    assert f(FT, FT(1.0), FT(2.0), FT(1.0)) == np.sqrt(0.5)


def test_smagorinsky_model_steps_and_stays_incompressible():
    # test/test_time_stepping.jl:243-256,384-400: every closure of the list time-steps; test/test_time_stepping.jl:124-158 incompressibility
    for closure, buoy in ((clo.Smagorinsky(), None), (clo.SmagorinskyLilly(0.23, 1.0, 1.0), clo.SeawaterBuoyancy())):
        g = Grid(np.float64, size=(8, 8, 8), extent=(1, 1, 1), topology=("P", "P", "B"), halo=(1, 1, 1))
        m = OracleModel(g, closure=closure, tracers=("T", "S"), buoyancy=buoy)
        assert m.grid.H == (2, 2, 2)                               # required halo 2 (Smagorinsky{…} <: AbstractScalarDiffusivity{…, 2})
        rng = np.random.default_rng(3)
        m.set(u=rng.uniform(-1, 1, m.u.interior.shape), v=rng.uniform(-1, 1, m.v.interior.shape),
              T=20 + rng.standard_normal(m.tracers["T"].interior.shape), S=35 + 0 * m.tracers["S"].interior)
        T0 = m.tracers["T"].interior.mean()
        for _ in range(3):
            m.time_step(1e-3)
        assert np.isfinite(m.u.interior).all() and m.nu_e.interior.min() >= 0 and m.nu_e.interior.max() > 0
        ctx = Ctx(m.grid, (1, 8), (1, 8), (1, 8))
        from oracle.operators import div_ccc
        assert np.abs(div_ccc(ctx, m.u, m.v, m.w)).max() < 1e-12
        assert abs(m.tracers["T"].interior.mean() - T0) < 1e-13 * 20


# ----------------------------------------------------------------------------- operators (test/test_operators.jl)
def _phi2_field(g, phi):
    """a ccc field holding ϕ² in its interior (the point function f(i, j, k, grid, ϕ) = ϕ[i, j, k]^2 of the reference test)"""
    f = Field(g, "ccc")
    f.interior[...] = phi ** 2
    return f


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_function_differences_derivatives_interpolations(FT):
    # test/test_operators.jl:52-86 (test_function_differentiation: extent 3 => Δ = 1, ∂ == δ), :125-152 (test_function_interpolation),
    # :8-50 (differences, on the non-immersed grid) — exact equality at (i, j, k) = (2, 2, 2)
    from oracle.operators import dC, dF, ddC, iC, iF
    g = Grid(FT, size=(3, 3, 3), extent=(3, 3, 3), topology=("P", "P", "B"))
    phi = np.random.default_rng(5).random((3, 3, 3)).astype(FT)
    p2 = phi ** 2
    ctx = Ctx(g, (2, 2), (2, 2), (2, 2))
    q = ctx.field(_phi2_field(g, phi))
    at = lambda i, j, k: p2[i - 1, j - 1, k - 1]
    for d in range(3):
        lo, hi = [2, 2, 2], [2, 2, 2]
        lo[d], hi[d] = 1, 3
        assert dC(ctx, q, d)(O)[0, 0, 0] == at(*hi) - at(2, 2, 2)              # δᶜ: ϕ²[i+1] − ϕ²[i]
        assert dF(ctx, q, d)(O)[0, 0, 0] == at(2, 2, 2) - at(*lo)              # δᶠ: ϕ²[i] − ϕ²[i−1]
        assert ddC(ctx, q, d)(O)[0, 0, 0] == at(*hi) - at(2, 2, 2)             # ∂ᶜ with Δ = 1
        assert ddF(ctx, q, d)(O)[0, 0, 0] == at(2, 2, 2) - at(*lo)
        assert iC(ctx, q, d)(O)[0, 0, 0] == (at(*hi) + at(2, 2, 2)) / 2        # ℑᶜ
        assert iF(ctx, q, d)(O)[0, 0, 0] == (at(2, 2, 2) + at(*lo)) / 2        # ℑᶠ


def test_derivatives_on_a_stretched_z():
    # test/test_operators.jl:88-123 (test_function_differentiation, stretched part), restricted to the z direction (the only one this
    # path stretches): faces (0, 1, 3, 6) => Δzᶜ = (1, 2, 3), centres (0.5, 2, 4.5) => Δzᶠ(2) = 1.5;  ∂z = δz / Δz at (2, 2, 2)
    from oracle.operators import ddC
    FT = np.float64
    g = Grid(FT, size=(3, 3, 3), x=(0, 3), y=(0, 3), z=[0.0, 1.0, 3.0, 6.0], topology=("B", "B", "B"))
    phi = np.random.default_rng(6).random((3, 3, 3))
    p2 = phi ** 2
    ctx = Ctx(g, (2, 2), (2, 2), (2, 2))
    q = ctx.field(_phi2_field(g, phi))
    assert ddC(ctx, q, 2)(O)[0, 0, 0] == 1 / 2.0 * (p2[1, 1, 2] - p2[1, 1, 1])          # 1/dc(2) (ϕ²[k+1] − ϕ²[k]), dc(2) = 3 − 1
    assert ddF(ctx, q, 2)(O)[0, 0, 0] == 1 / 1.5 * (p2[1, 1, 1] - p2[1, 1, 0])          # 1/df(2) (ϕ²[k] − ϕ²[k−1]), df(2) = 2 − 0.5
    assert ddC(ctx, q, 0)(O)[0, 0, 0] == p2[2, 1, 1] - p2[1, 1, 1]                      # regular x, Δ = 1


def test_grid_lengths_areas_volumes():
    # test/test_operators.jl:154-222: RectilinearGrid(size=(1,1,1), extent=(π, 2π, 3π)): Δ = π, 2π, 3π; Ax = 6π², Ay = 3π², Az = 2π², V = 6π³
    FT = np.float64
    g = Grid(FT, size=(1, 1, 1), extent=(np.pi, 2 * np.pi, 3 * np.pi), topology=("P", "P", "B"))
    pi = FT(np.pi)
    assert g.D == (pi, FT(2 * np.pi), FT(3 * np.pi))
    ctx = Ctx(g, (1, 1), (1, 1), (1, 1))
    for zl in "cf":
        assert ctx.area(0, zl)(O) == FT(2 * np.pi) * FT(3 * np.pi) and np.isclose(ctx.area(0, zl)(O), 6 * np.pi ** 2, rtol=4e-16)
        assert ctx.area(1, zl)(O) == pi * FT(3 * np.pi) and np.isclose(ctx.area(1, zl)(O), 3 * np.pi ** 2, rtol=4e-16)
        assert ctx.area(2, zl)(O) == pi * FT(2 * np.pi) and np.isclose(ctx.area(2, zl)(O), 2 * np.pi ** 2, rtol=4e-16)
        assert np.isclose(ctx.vol(zl)(O), 6 * np.pi ** 3, rtol=4e-16)
        assert ctx.dz(zl)(O) == FT(3 * np.pi)


def test_flat_dimension_operators():
    # test/test_operators.jl:262-304: on a grid with a Flat dimension, differences along it vanish and the other directions act as in 3-D
    from oracle.operators import dC, dF, iC, iF
    FT = np.float64
    g = Grid(FT, size=(4, 5), extent=(1, 1), topology=("F", "P", "B"))
    rng = np.random.default_rng(7)
    f = Field(g, "ccc")
    f.interior[...] = rng.random(f.interior.shape)
    fill_halo_regions(f)
    ctx = Ctx(g, (1, 1), (2, 3), (2, 4))
    q = ctx.field(f)
    assert np.all(dC(ctx, q, 0)(O) == 0) and np.all(dF(ctx, q, 0)(O) == 0)
    assert np.all(iC(ctx, q, 0)(O) == q(O)) and np.all(iF(ctx, q, 0)(O) == q(O))         # Flat interpolation = identity
    a = f.interior
    assert np.array_equal(dC(ctx, q, 1)(O)[0], a[0, 2:4, 1:4] - a[0, 1:3, 1:4])
    assert np.array_equal(dF(ctx, q, 2)(O)[0], a[0, 1:3, 1:4] - a[0, 1:3, 0:3])


# ----------------------------------------------------------------------------- Coriolis family (test/test_coriolis.jl, test/test_dynamics.jl)
def _coriolis_constructor_checks(mod, exc):
    """test/test_coriolis.jl:17-51 (values) and :104-119 (argument errors), for the oracle (mod = oracle.closures) or the product API"""
    pi, e = np.pi, np.e
    assert np.isclose(mod.FPlane(f=pi).f, pi)
    assert np.isclose(mod.FPlane(rotation_rate=2, latitude=30).f, 2.0)
    c = mod.ConstantCartesianCoriolis(f=1, rotation_axis=[0, np.cos(np.pi / 4), np.sin(np.pi / 4)])
    assert np.isclose(c.fy, np.cos(np.pi / 4)) and np.isclose(c.fz, np.sin(np.pi / 4)) and c.fx == 0
    c = mod.ConstantCartesianCoriolis(f=10, rotation_axis=[np.sqrt(1 / 3)] * 3)
    assert np.allclose([c.fx, c.fy, c.fz], 10 * np.sqrt(1 / 3))
    b = mod.BetaPlane(f0=pi, beta=2 * pi)
    assert np.isclose(b.f0, pi) and np.isclose(b.beta, 2 * pi)
    b = mod.BetaPlane(latitude=70, radius=2 * pi, rotation_rate=3 * pi)
    assert np.isclose(b.f0, 6 * pi * np.sin(np.deg2rad(70))) and np.isclose(b.beta, 6 * pi * np.cos(np.deg2rad(70)) / (2 * pi))
    # test/test_coriolis.jl:53-68 (values) and :121-133 (argument errors)
    n = mod.NonTraditionalBetaPlane(fz=pi, fy=e, beta=1 / 7, gamma=5)
    assert np.isclose(n.fz, pi) and np.isclose(n.fy, e) and np.isclose(n.beta, 1 / 7) and np.isclose(n.gamma, 5)
    n = mod.NonTraditionalBetaPlane(rotation_rate=pi, latitude=17, radius=e)
    s17, c17 = np.sin(np.deg2rad(17)), np.cos(np.deg2rad(17))
    assert np.isclose(n.fz, 2 * pi * s17) and np.isclose(n.fy, 2 * pi * c17)
    assert np.isclose(n.beta, 2 * pi * c17 / e) and np.isclose(n.gamma, -4 * pi * s17 / e) and n.R == e
    for kwargs in ({}, dict(rotation_rate=7e-5), dict(fz=1, latitude=40), dict(fz=1, rotation_rate=7e-5, latitude=40), dict(fy=1, latitude=40),
                   dict(fy=1, rotation_rate=7e-5, latitude=40), dict(fz=1, fy=2, latitude=40), dict(fz=1, fy=2, rotation_rate=7e-5, latitude=40),
                   dict(fz=1, fy=2, beta=3, latitude=40), dict(fz=1, fy=2, beta=3, rotation_rate=7e-5, latitude=40),
                   dict(fz=1, fy=2, beta=3, gamma=4, latitude=40), dict(fz=1, fy=2, beta=3, gamma=4, rotation_rate=7e-5, latitude=40)):
        with pytest.raises(exc):
            mod.NonTraditionalBetaPlane(**kwargs)
    for bad in (lambda: mod.FPlane(), lambda: mod.FPlane(rotation_rate=7e-5), lambda: mod.FPlane(f=1, latitude=40),
                lambda: mod.FPlane(f=1, rotation_rate=7e-5, latitude=40),
                lambda: mod.ConstantCartesianCoriolis(rotation_axis=[0, 1, 1]), lambda: mod.ConstantCartesianCoriolis(f=1, latitude=45),
                lambda: mod.ConstantCartesianCoriolis(fx=1, latitude=45), lambda: mod.ConstantCartesianCoriolis(fx=1, f=1),
                lambda: mod.ConstantCartesianCoriolis(f=1, rotation_axis=[0, 1, 1]),
                lambda: mod.BetaPlane(), lambda: mod.BetaPlane(f0=1), lambda: mod.BetaPlane(beta=1),
                lambda: mod.BetaPlane(f0=1e-4, beta=1e-11, latitude=70)):
        with pytest.raises(exc):
            bad()


def test_coriolis_constructors():
    _coriolis_constructor_checks(clo, ValueError)


def test_nontraditional_beta_plane_terms_for_uniform_flows():
    """non_traditional_beta_plane.jl:79-96 evaluated for uniform flows (interpolations of constants are exact): with
    2Ωʸ = fy (1 − z/R) + γ y and 2Ωᶻ = fz (1 + 2z/R) + β y,
      u = 1:  y_f_cross_U = 2Ωᶻ(y_f, z_c),  z_f_cross_U = −2Ωʸ(y_c, z_f),  x_f_cross_U = 0
      v = 1:  x_f_cross_U = −2Ωᶻ(y_c, z_c);   w = 1:  x_f_cross_U = +2Ωʸ(y_c, z_c)"""
    FT = np.float64
    g = Grid(FT, size=(6, 5, 4), extent=(3.0, 2.0, 2.0), topology=("P", "P", "P"), halo=(2, 2, 2))
    cor = clo.NonTraditionalBetaPlane(fz=0.7, fy=-0.5, beta=2.0, gamma=1.5, radius=3.0)
    ctx = Ctx(g, (1, g.Nx), (1, g.Ny), (1, g.Nz))
    yc, yf = g.nodes(1, "c")[None, :, None], g.nodes(1, "f")[None, :g.Ny, None]
    zc, zf = g.nodes(2, "c")[None, None, :], g.nodes(2, "f")[None, None, :g.Nz]
    Oy = lambda y, z: cor.fy * (1 - z / cor.R) + cor.gamma * y
    Oz = lambda y, z: cor.fz * (1 + 2 * z / cor.R) + cor.beta * y
    def fields(which):
        U = (Field(g, "fcc"), Field(g, "cfc"), Field(g, "ccf"))
        U[which].data[...] = 1.0
        return U
    one = np.ones((g.Nx, 1, 1))
    U = fields(0)
    assert np.allclose(clo.coriolis_cross(ctx, cor, U, 1), one * Oz(yf, zc), rtol=1e-14, atol=0)
    assert np.allclose(clo.coriolis_cross(ctx, cor, U, 2), -one * Oy(yc, zf), rtol=1e-14, atol=0)
    assert np.all(clo.coriolis_cross(ctx, cor, U, 0) == 0)
    assert np.allclose(clo.coriolis_cross(ctx, cor, fields(1), 0), -one * Oz(yc, zc), rtol=1e-14, atol=0)
    assert np.allclose(clo.coriolis_cross(ctx, cor, fields(2), 0), one * Oy(yc, zc), rtol=1e-14, atol=0)
    # the coefficients vary over the domain (the z/R and β y, γ y factors are active)
    assert np.ptp(Oz(yf, zc)) > 0.5 and np.ptp(Oy(yc, zf)) > 0.5


def test_inertial_oscillations_with_rotation_about_different_axes():
    # test/test_dynamics.jl:357-397: a uniform flow under FPlane(f = 1) (u = 1) and under ConstantCartesianCoriolis(f = 1, axis = x̂) (v = 1)
    # for half an inertial period: w_z == 0, u_x == 0, |U| ≈ 1, u_z ≈ v_x, v_z ≈ w_x.  (The reference uses a (Flat, Flat, Flat) grid and
    # Δt = 1e-3; a 2×2×2 periodic box with uniform fields is the same ODE — every interpolation of a uniform field is the identity —
    # and Δt = 1e-2 keeps the oracle fast: RK3 damps the amplitude by (fΔt)⁴/24 per step, 1.3e-7 in total.)
    f0, dt = 1.0, 1e-2
    nsteps = int(round(np.pi / f0 / dt))
    out = {}
    for axis, cor, ic in (("x", clo.ConstantCartesianCoriolis(f=f0, rotation_axis=(1, 0, 0)), "v"), ("z", clo.FPlane(f=f0), "u")):
        g = Grid(np.float64, size=(2, 2, 2), extent=(1, 1, 1), topology=("P", "P", "P"))
        m = OracleModel(g, coriolis=cor)
        m.set(**{ic: np.ones(m.fields[ic].interior.shape)})
        for _ in range(nsteps):
            m.time_step(dt)
        for n in "uvw":
            a = m.fields[n].interior
            assert np.ptp(a) < 1e-14                       # the flow stays uniform
        out[axis] = tuple(float(m.fields[n].interior[0, 0, 0]) for n in "uvw")
    (u_x, v_x, w_x), (u_z, v_z, w_z) = out["x"], out["z"]
    assert w_z == 0 and u_x == 0
    assert np.isclose(np.hypot(v_x, w_x), 1.0, rtol=1e-6) and np.isclose(np.hypot(u_z, v_z), 1.0, rtol=1e-6)
    assert np.isclose(u_z, v_x, rtol=1e-12, atol=1e-12) and np.isclose(v_z, w_x, rtol=1e-12, atol=1e-12)
    # half an inertial period reverses the flow: (u, v)(t) = (cos f t, −sin f t)
    t = nsteps * dt
    assert np.isclose(u_z, np.cos(f0 * t), atol=1e-6) and np.isclose(v_z, -np.sin(f0 * t), atol=1e-6)


def test_beta_plane_reduces_to_f_plane_and_varies_linearly_in_y():
    # beta_plane.jl:56-72: x_f_cross_U = −(f₀ + β y) ℑxy v with y = ynode(fcc).  With v = 1: −x_f_cross_U = f₀ + β y_c(j) exactly.
    FT = np.float64
    g = Grid(FT, size=(4, 6, 2), x=(0, 1), y=(-3.0, 3.0), z=(-1, 0), topology=("P", "P", "P"))
    u, v, w = Field(g, "fcc"), Field(g, "cfc"), Field(g, "ccf")
    v.data[...] = 1.0
    u.data[...] = 2.0
    ctx = Ctx(g, (1, 4), (1, 6), (1, 2))
    bp = clo.BetaPlane(f0=0.5, beta=0.25)
    yc, yf = g.nodes(1, "c"), g.nodes(1, "f")
    x = clo.coriolis_cross(ctx, bp, (u, v, w), 0)
    y = clo.coriolis_cross(ctx, bp, (u, v, w), 1)
    assert np.allclose(-x[0, :, 0], 0.5 + 0.25 * yc, rtol=1e-15) and np.allclose(y[0, :, 0], 2.0 * (0.5 + 0.25 * yf), rtol=1e-15)
    assert np.all(clo.coriolis_cross(ctx, bp, (u, v, w), 2) == 0)
    fp = clo.coriolis_cross(ctx, clo.FPlane(f=0.5), (u, v, w), 0)
    assert np.array_equal(clo.coriolis_cross(ctx, clo.BetaPlane(f0=0.5, beta=0.0), (u, v, w), 0), fp)


# ---------------------------------------------------------------------------------------------------------------------
# Reference-held pins restated in round 2 (VERDICT r01 "missing" item 3)
# ---------------------------------------------------------------------------------------------------------------------
def gaussian_advection_setup(N=128):
    """passive_tracer_advection_test  test/test_dynamics.jl:177-208 — the numbers of the reference test"""
    Ld, U, V = 1.0, 0.5, 0.8
    delta, x0, y0 = Ld / 15, Ld / 2, Ld / 2
    dt = 0.05 * Ld / N / np.sqrt(U ** 2 + V ** 2)
    T = lambda x, y, z, t: np.exp(-((x - U * t - x0) ** 2 + (y - V * t - y0) ** 2) / (2 * delta ** 2))
    return Ld, U, V, dt, T


def gaussian_relative_error(T_num, nodes, T, time):
    """relative_error(u_num, u, time)  test/test_dynamics.jl:10-15 : mean((num - exact)²) / mean(exact²)"""
    X, Y, Z = np.meshgrid(*nodes, indexing="ij")
    ans = T(X, Y, Z, time)
    return float(np.mean((T_num - ans) ** 2) / np.mean(ans ** 2))


@pytest.mark.parametrize("ts", ["QuasiAdamsBashforth2"])          # the reference runs AB2 only (test_dynamics.jl:628)
def test_passive_tracer_advection(ts):
    """A Gaussian advected diagonally by a uniform flow for 100 steps on the reference's 128 × 128 × 2 grid with its defaults
    (Centered(2), (Periodic, Periodic, Bounded), κ = ν = 1e-12, SeawaterBuoyancy, T and S): relative error < 1e-4."""
    N, Nt = 128, 100
    Ld, U, V, dt, T = gaussian_advection_setup(N)
    g = Grid(np.float64, size=(N, N, 2), extent=(Ld, Ld, Ld), topology=("P", "P", "B"))
    m = OracleModel(g, closure=clo.ScalarDiffusivity(nu=1e-12, kappa=1e-12), buoyancy=clo.SeawaterBuoyancy(), tracers=("T", "S"),
                    timestepper=ts)
    m.set(u=lambda x, y, z: U + 0 * x, v=lambda x, y, z: V + 0 * x, T=lambda x, y, z: T(x, y, z, 0.0))
    for _ in range(Nt):
        m.time_step(dt)
    nodes = [g.nodes(0, "c"), g.nodes(1, "c"), g.nodes(2, "c")]
    err = gaussian_relative_error(m.tracers["T"].interior, nodes, T, m.clock.time)
    assert err < 1e-4, err
    assert np.abs(m.tracers["S"].interior).max() == 0.0                       # S was never set


# (topology, extents) and the (field, side, L) list of test/test_boundary_conditions_integration.jl:309-360
FLUX_BUDGET_MATRIX = []
for _topo, _names, _sides in ((("P", "B", "B"), ("u", "c"), (("north", 0.4), ("south", 0.4), ("top", 0.5), ("bottom", 0.5))),
                              (("B", "P", "B"), ("v", "c"), (("east", 0.3), ("west", 0.3), ("top", 0.5), ("bottom", 0.5))),
                              (("B", "B", "P"), ("w", "c"), (("east", 0.3), ("west", 0.3), ("north", 0.4), ("south", 0.4)))):
    for _n in _names:
        for _s, _L in _sides:
            FLUX_BUDGET_MATRIX.append((_topo, _n, _s, _L))


@pytest.mark.parametrize("topo,name,side,Lside", FLUX_BUDGET_MATRIX, ids=["".join(t) + f"-{n}-{s}" for t, n, s, _ in FLUX_BUDGET_MATRIX])
def test_nonhydrostatic_flux_budget_matrix(topo, name, side, Lside):
    """test_nonhydrostatic_flux_budget (test/test_boundary_conditions_integration.jl:28-52): every field × side of the reference's
    matrix on its 2 × 2 × 2 grid, flux = ±π, one time step of Δt = 1: mean(ϕ) ≈ flux t / L."""
    g = Grid(np.float64, size=(2, 2, 2), x=(0, 0.3), y=(0, 0.4), z=(0, 0.5), topology=topo)
    flux = np.pi
    direction = 1 if side in ("west", "south", "bottom") else -1
    m = OracleModel(g, tracers=("c",), boundary_conditions={name: {side: BC("flux", flux * direction)}})
    m.fields[name].set(0)
    m.time_step(1.0)
    mean = float(m.fields[name].interior.mean())
    if name in ("u", "v", "w") and topo["uvw".index(name)] == "B":
        pytest.skip("not in the reference matrix")       # (the matrix only lists tangential velocities; kept for safety)
    assert np.isclose(mean, flux * m.clock.time / Lside, rtol=1e-8), (mean, flux * m.clock.time / Lside)


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_fluxes_with_diffusivity_boundary_conditions_are_correct(FT):
    """test/test_boundary_conditions_integration.jl:54-103: AMD with a Value boundary condition κ₀ on κₑ at the bottom and a Gradient
    condition bz on b there: the bottom diffusive flux is −κ₀ bz whatever the closure computes inside, so the mean of b changes by
    flux · t / Lz  (the reference's own numbers: −3.14159265e-5 after 10 steps; atol 1e-6)."""
    Lz = 1.0
    k0 = FT(np.exp(-3))
    bz = FT(np.pi)
    flux = -k0 * bz
    g = Grid(FT, size=(16, 16, 16), extent=(1, 1, Lz), topology=("P", "P", "B"))
    bcs = {"b": {"bottom": BC("gradient", bz)}, "kappa_e": {"b": {"bottom": BC("value", k0)}}}
    m = OracleModel(g, timestepper="QuasiAdamsBashforth2", tracers=("b",), buoyancy=clo.BuoyancyTracer(),
                    closure=clo.AnisotropicMinimumDissipation(), boundary_conditions=bcs)
    m.set(b=lambda x, y, z: z * float(bz))
    mean0 = float(m.tracers["b"].interior.astype(np.float64).mean())
    dt = 1e-6 * Lz ** 2 / float(k0)
    for n in range(10):
        m.time_step(dt, euler=(n == 0))
    mean1 = float(m.tracers["b"].interior.astype(np.float64).mean())
    assert abs((mean1 - mean0) - float(flux) * m.clock.time / Lz) <= 1e-6
    if FT == np.float64:
        assert np.isclose(mean1 - mean0, -3.141592656086267e-5, rtol=1e-6)      # the Float64 value quoted in the reference test


@pytest.mark.parametrize("order", [5, 7, 9])
def test_weno_high_order_tables(order):
    """WENO(order = 5 | 7 | 9) tables restated from weno_interpolants.jl:81-90 (C★), :117-118 (coeff_p), :175-185 (smoothness coefficients):
    (i) the C★-weighted candidates ARE the UpwindBiased(order) reconstruction (what "optimal weights" means: Balsara & Shu), which pins
    C★ and the ordering of coeff_p against the independently restated stencil coefficients; (ii) the smoothness forms of mirrored
    sub-stencils are mirror images (a typo in one of the 115 decimals breaks it) and vanish on constants; (iii) smooth data: every
    candidate reproduces polynomials of degree < buffer exactly, so the result does whatever the weights are."""
    FT = np.float64
    w = adv.WENO(FT, order)
    B = w.buffer
    # (i) Σ_r C★_r p_r = UpwindBiased(order) left reconstruction
    comb = np.zeros(2 * B - 1)
    for r in range(B):
        for j in range(B):
            comb[B - 1 - r + j] += float(w.cstar[r]) * float(w.coeff_p[r][j])        # S_r = (q[B-1-r] … q[2B-2-r])
    left = adv.stencil_coefficients(FT, B - 2, order)
    up = np.array([float(left[order - 1 - n]) for n in range(order)])                # q_n = ψ[i-B+n] -> coeff_left[order-1-n]
    assert np.allclose(comb, up, atol=1e-14)
    assert abs(sum(float(c) for c in w.cstar) - 1) < 1e-15
    # (ii) mirror symmetry and constants
    def quad(C):
        Q, c = np.zeros((B, B)), 0
        for s_ in range(B):
            for i in range(s_, B):
                Q[s_, i] += C[c + i - s_]
            c += B - s_
        return Q
    for r in range(B):
        Q, Qm = quad([float(x) for x in w.smooth[r]]), quad([float(x) for x in w.smooth[B - 1 - r]])
        S, Sm = Q + Q.T, Qm + Qm.T
        assert np.allclose(S, Sm[::-1, ::-1], atol=1e-12)
        assert abs(S.sum()) < 1e-9                                                   # β of a constant field is zero
    # (iii) polynomial of degree B - 1 through the conditional interpolation
    g = Grid(FT, size=(24, 4, 4), extent=(24, 4, 4), topology=("P", "P", "P"), halo=(B, B, B))
    f = Field(g, "ccc")
    x = np.arange(-B, 24 + B) + 0.5
    deg = B - 1
    # cell averages of x^deg over unit cells: ((x+½)^(deg+1) − (x−½)^(deg+1)) / (deg+1)
    f.data[...] = (((x + 0.5) ** (deg + 1) - (x - 0.5) ** (deg + 1)) / (deg + 1))[:, None, None]
    ctx = Ctx(g, (8, 16), (1, 4), (1, 4))
    for left_bias in (True, False):
        val = adv.biased_face(ctx, w, ctx.field(f), 0, lambda o: np.full(ctx.shape, left_bias))(O)
        xf = np.arange(8, 17) - 1.0
        assert np.allclose(val[:, 0, 0], xf ** deg, rtol=1e-10, atol=1e-8)
