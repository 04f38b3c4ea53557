"""On-device step diagnostics: cell_advection_timescale, max|u,v,w|, hasnan and the TimeStepWizard logic
(src/Advection/cell_advection_timescale.jl:13-34, src/Simulations/time_step_wizard.jl:65-116, src/Diagnostics/nan_checker.jl)."""
import math
import os

import numpy as np
import pytest

import parity_harness as ph


def _check(library):
    import oceananigans_b200 as ob
    for topo, N in (("PPB", (16, 12, 8)), ("PPP", (40, 9, 7)), ("PPF", (12, 10, 1))):
        m, om = ph.build_pair(N=N, topo=topo, scheme="weno", closure="none", buoy="none", library=library)
        ic = ph.initial_conditions(om)
        ob.set_(m, **ic)
        om.set(**ic)
        d = ob.step_diagnostics(m)
        g = om.grid
        u, v, w = (m.fields[n].interior().astype(np.float64) for n in "uvw")      # the device state itself
        nz = N[2]
        inv = np.abs(u) / g.D[0] + np.abs(v) / g.D[1] + (0 if topo[2] == "F" else np.abs(w[:, :, :nz]) / g.D[2])
        assert abs(d["cell_advection_timescale"] - (1 / inv).min()) <= 1e-14 * (1 / inv).min()
        assert d["max_abs_u"] == np.abs(u).max() and d["max_abs_v"] == np.abs(v).max()
        assert d["max_abs_w"] == np.abs(w[:, :, :nz]).max()
        assert d["has_nan"] is False and ob.hasnan(m) is False
        # wizard: new Δt = clamp(min(max_change Δt, max(min_change Δt, cfl τ)))
        wiz = ob.TimeStepWizard(cfl=0.5, max_change=1.1, min_change=0.5)
        tau = d["cell_advection_timescale"]
        assert wiz.new_time_step(1e-9, m) == pytest.approx(1.1e-9)
        assert wiz.new_time_step(10.0, m) == pytest.approx(5.0)
        assert wiz.new_time_step(0.5 * tau, m) == pytest.approx(0.5 * tau)
        bad = ic["u"].copy()
        bad[1, 2, 0] = np.nan
        m.velocities.u.set(bad)
        assert ob.hasnan(m) is True
    with pytest.raises(ValueError):
        ob.TimeStepWizard(max_change=0.9)


def test_diagnostics_hostsim():
    from oceananigans_b200 import _lib
    import __graft_entry__ as ge
    if not os.path.exists(ge.HOSTSIM):
        ge.build()
    _check(_lib.Library(ge.HOSTSIM))


@pytest.mark.gpu
def test_diagnostics_cuda():
    from oceananigans_b200 import _lib
    _lib.load()
    _check(None)


def _wizard_known_answers(library, n):
    """wall_time_step_wizard_tests of the reference (test/test_simulations.jl:14-76), through the host API: the same numbers
    (CFL = 0.45, u₀ = 7, Δt = 2.5, ν = 1) on an n×n×n grid of extent 1 (the reference uses n = 1)."""
    import oceananigans_b200 as ob
    kw = {} if library is None else {"library": library}
    grid = ob.RectilinearGrid(np.float64, size=(n, n, n), extent=(1, 1, 1))
    dx = grid.dx
    model = ob.NonhydrostaticModel(grid=grid, **kw)
    CFL, u0, dt = 0.45, 7.0 * dx, 2.5           # u₀ scaled with Δx so that n > 1 reproduces the reference's numbers (Δx = 1 there)
    u = np.zeros((n, n, n)); u[0, 0, 0] = u0
    model.velocities.u.set(u)
    assert model.velocities.u.maximum_abs() == u0 and model.velocities.w.maximum_abs() == 0.0
    dt = ob.TimeStepWizard(cfl=CFL, max_change=math.inf, min_change=0).new_time_step(dt, model)
    assert dt == pytest.approx(CFL * dx / u0, rel=1e-14)
    assert ob.TimeStepWizard(cfl=CFL, max_change=math.inf, min_change=0.75).new_time_step(1.0, model) == pytest.approx(0.75)
    dt = ob.TimeStepWizard(cfl=CFL, max_change=math.inf, min_change=0, min_Δt=1.99).new_time_step(dt, model)
    assert dt == pytest.approx(1.99)
    u[0, 0, 0] = u0 / 100
    model.velocities.u.set(u)
    assert ob.TimeStepWizard(cfl=CFL, max_change=1.1, min_change=0).new_time_step(1.0, model) == pytest.approx(1.1)
    dt = ob.TimeStepWizard(cfl=CFL, max_change=math.inf, min_change=0, max_Δt=3.99).new_time_step(dt, model)
    assert dt == pytest.approx(3.99)
    # diffusive CFL: Δt = diff_CFL Δx² / ν
    model = ob.NonhydrostaticModel(grid=grid, closure=ob.ScalarDiffusivity(nu=1.0), **kw)
    dt = ob.TimeStepWizard(cfl=math.inf, diffusive_cfl=0.45, max_change=math.inf, min_change=0).new_time_step(dt, model)
    assert dt == pytest.approx(max(0.45 * dx ** 2 / 1.0, 0.0), rel=1e-14)
    assert ob.cell_diffusion_timescale(ob.NonhydrostaticModel(grid=grid, **kw)) == math.inf
    # "stretched" grid with z = k -> k faces (spacing 1): the advective answer only depends on Δx
    if n == 1:
        gs = ob.RectilinearGrid(np.float64, size=(1, 1, 2), x=(0, 1), y=(0, 1), z=lambda k: float(k), halo=(1, 1, 1))
        model = ob.NonhydrostaticModel(grid=gs, **kw)
        model.velocities.u.set(np.full((1, 1, 2), u0))
        dt = ob.TimeStepWizard(cfl=CFL, max_change=math.inf, min_change=0).new_time_step(2.5, model)
        assert dt == pytest.approx(CFL * gs.dx / u0, rel=1e-14)


def _diffusion_timescale_of_eddy_closures(library):
    """cell_diffusion_timescale for Smagorinsky / AMD (turbulence_closure_diagnostics.jl:57-69): Δ² / (max νₑ · max(1, 1/min Pr)) and
    min(Δ²/max νₑ, Δ²/max κₑ), with the maxima reduced on the device."""
    import oceananigans_b200 as ob
    for closure, kind in (("lilly", "smag"), ("amd", "amd")):
        m, om = ph.build_pair(N=(16, 12, 8), topo="PPB", scheme="weno", closure=closure, library=library)
        ic = ph.initial_conditions(om)
        ob.set_(m, **ic)
        om.set(**ic)
        d2 = min(float(om.grid.D[d]) for d in range(3)) ** 2
        nu = m.diffusivity_fields.nu_e.interior().max()
        assert m.diffusivity_fields.nu_e.maximum_abs() == nu
        if kind == "smag":
            want = d2 / (nu * max(1.0, 1.0 / 1.0))            # Pr = (1, 1.5) in the harness: min Pr = 1
            want = min(want, d2 / 1.05e-6, d2 / 1.46e-7)      # the ScalarDiffusivity of the tuple
        else:
            want = min([d2 / nu] + [d2 / f.interior().max() for f in m.diffusivity_fields.kappa_e.values()])
        assert ob.cell_diffusion_timescale(m) == pytest.approx(want, rel=1e-14)
        # NaN propagates like Julia's maximum
        bad = ic["T"].copy(); bad[3, 2, 1] = np.nan
        m.tracers["T"].set(bad)
        assert math.isnan(m.tracers["T"].maximum_abs())


def test_wizard_known_answers_hostsim():
    from oceananigans_b200 import _lib
    import __graft_entry__ as ge
    if not os.path.exists(ge.HOSTSIM):
        ge.build()
    lib = _lib.Library(ge.HOSTSIM)
    _wizard_known_answers(lib, 1)
    _wizard_known_answers(lib, 4)
    _diffusion_timescale_of_eddy_closures(lib)
