"""On-device step diagnostics: cell_advection_timescale, max|u,v,w|, hasnan and the TimeStepWizard logic
(src/Advection/cell_advection_timescale.jl:13-34, src/Simulations/time_step_wizard.jl:65-116, src/Diagnostics/nan_checker.jl)."""
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
