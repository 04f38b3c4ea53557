"""oracle/nhm_step.c (C99 + OpenMP twin, used as the multi-core CPU baseline) against the NumPy oracle: two independent
restatements of the reference algorithm must agree to round-off on the triply periodic benchmark physics."""
import numpy as np
import pytest

import parity_harness as ph


@pytest.mark.parametrize("N,scheme,tracers", [((16, 8, 8), "weno", True), ((12, 10, 8), "weno", True), ((16, 8, 8), "centered", False)])
def test_c_twin_matches_numpy_oracle(N, scheme, tracers):
    from oracle.c_twin import CTwin
    kw = dict(N=N, topo="PPP", scheme=scheme) if tracers else dict(N=N, topo="PPP", scheme=scheme, closure="none", buoy="none")
    om = ph.build_oracle(**kw)
    ic = ph.initial_conditions(om)
    om.set(**ic)
    ct = CTwin(N, ph.EXTENT, weno=scheme == "weno", tracers=tracers, nu=1e-3 if tracers else 0.0, kappa=2e-3 if tracers else 0.0)
    ct.set(**ic)
    dt = 0.1 * float(min(om.grid.D))
    for s in range(3):
        if s:
            om.time_step(dt)
            ct.time_step(dt)
        for n in om.fields:
            assert ph.rel_linf(ct.get(n), om.fields[n].interior) <= 2e-12, (s, n)
        assert ph.rel_linf(ct.get("p"), om.pNHS.interior) <= 1e-10, (s, "p")
