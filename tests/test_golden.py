"""Frozen vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).

CPU: the oracle still reproduces them bit-for-bit-ish (guards the checker against drift) and the host simulation of the kernels
matches them.  GPU (-m gpu): the CUDA library, through the C ABI, matches them to the north_star tolerance."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
import parity_harness as ph  # noqa: E402

CASES = sorted(mg.GOLDEN_CASES)
WIDENING = sorted(mg.GOLDEN_CASES_WIDENING)          # CPU checks here; their GPU check is in tests/test_widening_gpu.py
ALL = {**mg.GOLDEN_CASES, **mg.GOLDEN_CASES_WIDENING}


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def _fields(z, s):
    return {k[len(f"s{s}_"):]: z[k] for k in z.files if k.startswith(f"s{s}_")}


def _ic(z):
    return {k[3:]: z[k] for k in z.files if k.startswith("ic_")}


@pytest.mark.parametrize("name", CASES + WIDENING)
def test_oracle_reproduces_golden(name):
    z = _load(name)
    kw = ALL[name]
    om = ph.build_oracle(**kw)
    om.set(**_ic(z))
    tol = 1e-13 if kw.get("FT", np.float64) is np.float64 else 1e-5
    for s in range(1, max(mg.STEPS) + 1):
        om.time_step(float(z["dt"]))
        if s in mg.STEPS:
            for n, ref in _fields(z, s).items():
                got = om.pNHS.interior if n == "p" else om.fields[n].interior
                assert ph.rel_linf(got, ref) <= tol, (name, s, n)


def _product_vs_golden(name, library):
    import oceananigans_b200 as ob
    z = _load(name)
    kw = ALL[name]
    FT = kw.get("FT", np.float64)
    m = ph.build_product(library=library, **kw)
    ob.set_(m, **_ic(z))
    for s in range(1, max(mg.STEPS) + 1):
        ob.time_step_(m, float(z["dt"]))
        if s in mg.STEPS:
            for n, ref in _fields(z, s).items():
                got = m.pressures.pNHS.interior() if n == "p" else m.fields[n].interior()
                assert ph.rel_linf(got, ref) <= ph.TOL[FT], (name, s, n)


@pytest.mark.parametrize("name", CASES + WIDENING)
def test_hostsim_matches_golden(name):
    from oceananigans_b200 import _lib
    import __graft_entry__ as ge
    if not os.path.exists(ge.HOSTSIM):
        ge.build()
    _product_vs_golden(name, _lib.Library(ge.HOSTSIM))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden(name):
    from oceananigans_b200 import _lib
    _lib.load()
    _product_vs_golden(name, None)
