"""Known-answer tests the REFERENCE holds for this path, restated through the C ABI (missing item 3):

  * passive_tracer_advection_test             test/test_dynamics.jl:177-208   (Gaussian advected for 100 steps, rel. error < 1e-4)
  * test_nonhydrostatic_flux_budget           test/test_boundary_conditions_integration.jl:28-52, 309-360 (every field × side)
  * fluxes_with_diffusivity_boundary_conditions_are_correct   :54-103  (incl. the Float64 number quoted in the reference's comment)

The oracle versions are in tests/test_oracle_known_answers.py.  Here the same checks drive the product: the host-simulation build of the
kernel sources on CPU, the CUDA library under `-m gpu`."""
import numpy as np
import pytest

import oceananigans_b200 as ob
from test_oracle_known_answers import FLUX_BUDGET_MATRIX, gaussian_advection_setup, gaussian_relative_error

_TOPO = {"P": ob.Periodic, "B": ob.Bounded}


def _kw(library):
    return {} if library is None else {"library": library}


def passive_tracer_advection(library, ts="QuasiAdamsBashforth2", N=128, Nt=100):
    Ld, U, V, dt, T = gaussian_advection_setup(N)
    grid = ob.RectilinearGrid(np.float64, size=(N, N, 2), extent=(Ld, Ld, Ld))
    model = ob.NonhydrostaticModel(grid=grid, closure=ob.ScalarDiffusivity(nu=1e-12, kappa=1e-12), timestepper=ts,
                                   buoyancy=ob.SeawaterBuoyancy(), tracers=("T", "S"), **_kw(library))
    ob.set_(model, u=lambda x, y, z: U + 0 * x, v=lambda x, y, z: V + 0 * x, T=lambda x, y, z: T(x, y, z, 0.0))
    for _ in range(Nt):
        ob.time_step_(model, dt)
    nodes = [grid.nodes(d, ob.Center) for d in range(3)]
    err = gaussian_relative_error(model.tracers.T.interior(), nodes, T, model.clock.time)
    assert err < 1e-4, err
    return err


def flux_budget(library, topo, name, side, Lside):
    grid = ob.RectilinearGrid(np.float64, size=(2, 2, 2), x=(0, 0.3), y=(0, 0.4), z=(0, 0.5), topology=tuple(_TOPO[c] for c in topo))
    flux = np.pi
    direction = 1 if side in ("west", "south", "bottom") else -1
    bcs = {name: ob.FieldBoundaryConditions(**{side: ob.FluxBoundaryCondition(flux * direction)})}
    model = ob.NonhydrostaticModel(grid=grid, boundary_conditions=bcs, tracers=("c",), **_kw(library))
    sim = ob.Simulation(model, Δt=1.0, stop_iteration=1)
    ob.run_(sim)
    mean = float(model.fields[name].interior().mean())
    assert np.isclose(mean, flux * model.clock.time / Lside, rtol=1e-8), (mean, flux * model.clock.time / Lside)


def fluxes_with_diffusivity_boundary_conditions(library, FT):
    Lz = 1.0
    k0, bz = FT(np.exp(-3)), FT(np.pi)
    flux = -k0 * bz
    grid = ob.RectilinearGrid(FT, size=(16, 16, 16), extent=(1, 1, Lz))
    bcs = {"b": ob.FieldBoundaryConditions(bottom=ob.GradientBoundaryCondition(float(bz))),
           "κₑ": {"b": ob.FieldBoundaryConditions(bottom=ob.ValueBoundaryCondition(float(k0)))}}
    model = ob.NonhydrostaticModel(grid=grid, timestepper="QuasiAdamsBashforth2", tracers=("b",), buoyancy=ob.BuoyancyTracer(),
                                   closure=ob.AnisotropicMinimumDissipation(), boundary_conditions=bcs, **_kw(library))
    ob.set_(model, b=lambda x, y, z: z * float(bz))
    mean0 = float(model.tracers.b.interior().astype(np.float64).mean())
    dt = 1e-6 * Lz ** 2 / float(k0)
    for n in range(10):
        ob.time_step_(model, dt, euler=(n == 0))
    mean1 = float(model.tracers.b.interior().astype(np.float64).mean())
    assert abs((mean1 - mean0) - float(flux) * model.clock.time / Lz) <= 1e-6
    if FT == np.float64:
        assert np.isclose(mean1 - mean0, -3.141592656086267e-5, rtol=1e-6)      # the Float64 value quoted at :97-99 of the reference test


# ---- CPU: host simulation of the kernel sources ----------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hostsim():
    from test_host_api import _hostsim
    return _hostsim()


def test_passive_tracer_advection_hostsim(hostsim):
    passive_tracer_advection(hostsim, Nt=100)


@pytest.mark.parametrize("topo,name,side,Lside", FLUX_BUDGET_MATRIX, ids=["".join(t) + f"-{n}-{s}" for t, n, s, _ in FLUX_BUDGET_MATRIX])
def test_nonhydrostatic_flux_budget_matrix_hostsim(hostsim, topo, name, side, Lside):
    flux_budget(hostsim, topo, name, side, Lside)


@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_fluxes_with_diffusivity_boundary_conditions_hostsim(hostsim, FT):
    fluxes_with_diffusivity_boundary_conditions(hostsim, FT)


def test_diffusivity_bc_argument_errors(hostsim):
    grid = ob.RectilinearGrid(np.float64, size=(8, 8, 8), extent=(1, 1, 1))
    kbc = {"κₑ": {"b": ob.FieldBoundaryConditions(bottom=ob.ValueBoundaryCondition(0.1))}}
    with pytest.raises(ValueError):       # no closure with diffusivity fields
        ob.NonhydrostaticModel(grid=grid, tracers=("b",), boundary_conditions=kbc, library=hostsim)
    m = ob.NonhydrostaticModel(grid=grid, tracers=("b",), closure=ob.AnisotropicMinimumDissipation(), library=hostsim)
    from oceananigans_b200 import _lib as L
    assert hostsim.oc_set_diffusivity_bc(m._h, L.OC_FIELD_KAPPA_E0, 0, L.OC_BC_VALUE, 1.0) != 0       # west of a Periodic dimension
    assert hostsim.oc_set_diffusivity_bc(m._h, 0, 4, L.OC_BC_VALUE, 1.0) != 0                         # not a diffusivity field
    assert hostsim.oc_set_diffusivity_bc(m._h, L.OC_FIELD_NU_E, 4, L.OC_BC_VALUE, 1.0) == 0


# ---- GPU: the CUDA library -------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cuda():
    from oceananigans_b200 import _lib
    lib = _lib.load()
    assert lib.path.endswith("liboceananigans_b200.so")
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("ts", ["QuasiAdamsBashforth2", "RungeKutta3"])
def test_passive_tracer_advection_cuda(cuda, ts):
    passive_tracer_advection(cuda, ts=ts)


@pytest.mark.gpu
def test_nonhydrostatic_flux_budget_matrix_cuda(cuda):
    for topo, name, side, Lside in FLUX_BUDGET_MATRIX:
        flux_budget(cuda, topo, name, side, Lside)


@pytest.mark.gpu
@pytest.mark.parametrize("FT", [np.float64, np.float32])
def test_fluxes_with_diffusivity_boundary_conditions_cuda(cuda, FT):
    fluxes_with_diffusivity_boundary_conditions(cuda, FT)
