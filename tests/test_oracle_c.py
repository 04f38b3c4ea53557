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


def test_reference_smoothness_indicators_carry_the_round_off():
    """Why T and S agree with the reference's formulas only to ~1e-12 … 1e-8 (growing with the number of cells: the worst cell counts):
    the same model evaluated (a) in the reference's order in Float64, (b) in Float64 with the WENO smoothness indicators in difference
    form — the form the CUDA kernel uses (oc_march.h: weno5_value_c) —, (c) in the reference's order in x87 extended precision (64-bit
    mantissa).  (b) agrees with (c) to 1e-14; (a) does not: the difference is the reference formula's own cancellation (β is a small
    difference of terms ~ 30 S²), not an error of either restatement.  u, v, w agree to 1e-14 in all three."""
    from oracle.c_twin import CTwin
    N = (24, 24, 24)
    rng = np.random.default_rng(1234)
    ic = {n: rng.uniform(-1, 1, N) for n in ("u", "v", "w")}
    ic["T"] = 20 + 0.01 * rng.standard_normal(N)
    ic["S"] = 35 + 0.01 * rng.standard_normal(N)
    kw = dict(nu=1e-3, kappa=2e-3)
    ref, dif, ext = CTwin(N, ph.EXTENT, **kw), CTwin(N, ph.EXTENT, beta_difference_form=True, **kw), CTwin(N, ph.EXTENT, extended=True, **kw)
    dt = 0.1 * min(ph.EXTENT[d] / N[d] for d in range(3))
    for m in (ref, dif, ext):
        m.set(**ic)
        for _ in range(2):
            m.time_step(dt)
    for n in ("u", "v", "w"):
        assert ph.rel_linf(ref.get(n), ext.get(n)) < 1e-14 and ph.rel_linf(dif.get(n), ext.get(n)) < 1e-14, n
    for n in ("T", "S"):
        e_ref, e_dif = ph.rel_linf(ref.get(n), ext.get(n)), ph.rel_linf(dif.get(n), ext.get(n))
        assert e_dif < 1e-14, (n, e_dif)
        assert e_ref > 20 * e_dif and e_ref < 1e-10, (n, e_ref, e_dif)
