"""GPU parity tests (`-m gpu`) of the SURVEY §8f item 3 widening added after the last GPU session of round 1: the Smagorinsky /
SmagorinskyLilly closures.  Same harness and tolerances as tests/test_gpu_parity.py (through the C ABI, against the CPU oracle,
1 and 10 steps, interior and parent arrays).  The kernel logic of these cases is also exercised on CPU by the host-simulation
build (tests/test_hostsim_parity.py)."""
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


@pytest.mark.parametrize("name,kw", ph.SMAGORINSKY_CASES, ids=[c[0] for c in ph.SMAGORINSKY_CASES])
def test_cuda_matches_oracle_with_smagorinsky_closures(ob, name, kw):
    """Smagorinsky(coefficient, Pr) and SmagorinskyLilly(C, Cb, Pr): Smagorinskys/smagorinsky.jl:92-108, lilly_coefficient.jl:114-135"""
    ph.check_case(kw, library=None, steps=(1, 10))


@pytest.mark.parametrize("name,kw", ph.AMD_CB_CASES, ids=[c[0] for c in ph.AMD_CB_CASES])
def test_cuda_matches_oracle_with_amd_buoyancy_modification(ob, name, kw):
    """AnisotropicMinimumDissipation(; Cb = 1): anisotropic_minimum_dissipation.jl:62-68, 168-172, 310-323 (AmdKernel<FT, STR, CB = true>)"""
    ph.check_case(kw, library=None, steps=(1, 10))


def test_wizard_known_answers_and_diffusion_timescale_cuda(ob):
    """TimeStepWizard known answers of test/test_simulations.jl:14-76 (advective and diffusive CFL) and cell_diffusion_timescale of the
    eddy-viscosity closures, with maximum(νₑ) / maximum(κₑ) reduced on the device (oc_field_maximum_abs)."""
    import test_diagnostics as td
    td._wizard_known_answers(None, 4)
    td._diffusion_timescale_of_eddy_closures(None)


@pytest.mark.parametrize("FT", ["f64", "f32"])
def test_host_api_mirrors_the_reference_constructor_and_set(ob, FT):
    """test/test_nonhydrostatic_models.jl:22-69,101-195 through the CUDA library (tests/test_host_api.py holds the checks)"""
    import numpy as np
    import test_host_api as th
    if FT == "f64":
        th.model_construction(None)
        th.halo_adjustment(None)
    th.setting_model_fields(None, np.float64 if FT == "f64" else np.float32)


@pytest.mark.parametrize("name,kw", ph.CORIOLIS_CASES, ids=[c[0] for c in ph.CORIOLIS_CASES])
def test_cuda_matches_oracle_for_the_coriolis_family(ob, name, kw):
    """BetaPlane (beta_plane.jl:56-72) and ConstantCartesianCoriolis (constant_cartesian_coriolis.jl:70-81)"""
    ph.check_case(kw, library=None, steps=(1, 10))


def test_coriolis_constructors_and_inertial_oscillation_cuda(ob):
    """test/test_coriolis.jl:17-51,104-119 and the inertial oscillations of test/test_dynamics.jl:357-397 through the CUDA library"""
    import test_host_api as th
    th.coriolis_constructors_and_inertial_oscillation(None, size=(16, 12, 8))


@pytest.mark.parametrize("name,kw", ph.TILTED_CASES, ids=[c[0] for c in ph.TILTED_CASES])
def test_cuda_matches_oracle_with_tilted_gravity(ob, name, kw):
    """BuoyancyForce(formulation; gravity_unit_vector)  buoyancy_force.jl:47-58, g_dot_b.jl:1-3"""
    ph.check_case(kw, library=None, steps=(1, 10))


def test_stratified_fluid_remains_at_rest_with_tilted_gravity_cuda(ob):
    """test/test_dynamics.jl:263-353 through the CUDA library"""
    import test_host_api as th
    th.stratified_fluid_remains_at_rest_with_tilted_gravity(th._product_maker(None))


@pytest.mark.parametrize("name,kw", ph.WALL_BC_CASES, ids=[c[0] for c in ph.WALL_BC_CASES])
def test_cuda_matches_oracle_with_lateral_wall_bcs(ob, name, kw):
    """Value / Gradient / Flux boundary conditions on west / east / south / north walls"""
    ph.check_case(kw, library=None, steps=(1, 10))


@pytest.mark.parametrize("name,kw", ph.ARRAY_BC_CASES, ids=[c[0] for c in ph.ARRAY_BC_CASES])
def test_cuda_matches_oracle_with_array_valued_flux_bcs(ob, name, kw):
    """FluxBoundaryCondition(J::AbstractArray) on every Bounded side (oc_set_bc_array; compute_flux_bcs.jl:116-163)"""
    ph.check_case(kw, library=None, steps=(1, 10))


def test_asynchronous_output_cuda(ob):
    """snapshots taken by oc_output_begin are unaffected by the time steps issued before oc_output_wait"""
    import test_host_api as th
    th.asynchronous_output(None)


def _widening_goldens():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg
    return sorted(mg.GOLDEN_CASES_WIDENING)


@pytest.mark.parametrize("name", _widening_goldens())
def test_cuda_matches_widening_golden(ob, name):
    """frozen vectors of the widening features (tests/golden/make_golden.py: GOLDEN_CASES_WIDENING)"""
    import test_golden as tg
    tg._product_vs_golden(name, None)


@pytest.mark.parametrize("name,kw", ph.ADAPT_CASES, ids=[c[0] for c in ph.ADAPT_CASES])
def test_cuda_matches_oracle_with_adapted_advection_order(ob, name, kw):
    """adapt_advection_order (src/Advection/adapt_advection_order.jl:18-96): the scheme lowered per direction where N < buffer"""
    ph.check_case(kw, library=None, steps=(1, 10))


@pytest.mark.parametrize("name,kw", ph.ARRAY_DIFFUSIVITY_CASES, ids=[c[0] for c in ph.ARRAY_DIFFUSIVITY_CASES])
def test_cuda_matches_oracle_with_array_valued_diffusivities(ob, name, kw):
    """ScalarDiffusivity(ν = array, κ = (T = array, S = number)): abstract_scalar_diffusivity_closure.jl:323-332"""
    ph.check_case(kw, library=None, steps=(1, 10))


@pytest.mark.parametrize("name,kw", ph.WENO_HI_CASES, ids=[c[0] for c in ph.WENO_HI_CASES])
def test_cuda_matches_oracle_for_weno7_and_weno9(ob, name, kw):
    """WENO(order = 7 | 9): weno_interpolants.jl:81-90,175-185,303-307 — general tile kernel, halos of 4 / 5, order-reduction chains near walls"""
    ph.check_case(kw, library=None, steps=(1, 10))


# (kept last in the last GPU test file: added after the round's final GPU run — verified on the host simulation only so far)
@pytest.mark.parametrize("name,kw", ph.FLAT_CLOSURE_CASES, ids=[c[0] for c in ph.FLAT_CLOSURE_CASES])
def test_cuda_matches_oracle_with_eddy_closures_on_two_dimensional_grids(ob, name, kw):
    """AnisotropicMinimumDissipation / Smagorinsky(-Lilly) with a Flat dimension"""
    ph.check_case(kw, library=None, steps=(1, 10))
