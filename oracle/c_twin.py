"""ctypes front-end of oracle/nhm_step.c — the C99 + OpenMP twin of the NumPy oracle for the triply periodic benchmark physics.

TEST INFRASTRUCTURE ONLY.  Built by `__graft_entry__.build()` (gcc -O3 -fopenmp) into oracle/_build/; used by tests/test_oracle_c.py
(cross-check against the NumPy oracle) and by bench.py's CPU legs (multi-core CPU baseline).  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "nhm_step.c")
LIB = os.path.join(HERE, "_build", "libnhm_step.so")
LIB_EXTENDED = os.path.join(HERE, "_build", "libnhm_step_ld.so")      # the same formulas in x87 extended precision (-DNHC_LONG_DOUBLE)


def build(force=False, extended=False):
    lib = LIB_EXTENDED if extended else LIB
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.run(["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-std=gnu99", "-shared", "-fPIC"] +
                       (["-DNHC_LONG_DOUBLE"] if extended else []) + [SRC, "-o", lib, "-lm"], check=True)
    return lib


class CTwin:
    """NonhydrostaticModel on a (Periodic, Periodic, Periodic) regular grid: Centered(2) | WENO(5), optional T,S + SeawaterBuoyancy,
    ScalarDiffusivity(ν, κ), RungeKutta3."""
    NAMES = ("u", "v", "w", "T", "S", "p")

    def __init__(self, size, extent, weno=True, tracers=True, nu=0.0, kappa=0.0, g=9.80665, alpha=1.67e-4, beta=7.8e-4, extended=False, beta_difference_form=False):
        self.lib = C.CDLL(build(extended=extended))
        self.lib.nhc_create.restype = C.c_void_p
        self.lib.nhc_create.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int] * 2 + [C.c_double] * 5
        for fn in (self.lib.nhc_set, self.lib.nhc_get):
            fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.lib.nhc_finalize.argtypes = [C.c_void_p]
        self.lib.nhc_step.argtypes = [C.c_void_p, C.c_double]
        self.lib.nhc_destroy.argtypes = [C.c_void_p]
        self.N = tuple(int(n) for n in size)
        self.ntr = 2 if tracers else 0
        self.h = self.lib.nhc_create(*self.N, *[float(x) for x in extent], int(weno), self.ntr, nu, kappa, g, alpha, beta)
        # the WENO smoothness indicators in difference form (exact identity, well conditioned) instead of the reference's expanded form:
        # tests use it to attribute T / S differences to the reference's own Float64 round-off (see nhm_step.c: weno5)
        self.lib.nhc_set_beta_difference_form.argtypes = [C.c_void_p, C.c_int]
        self.lib.nhc_set_beta_difference_form(self.h, int(beta_difference_form))

    def set(self, **fields):
        for n, a in fields.items():
            a = np.asfortranarray(np.asarray(a, dtype=np.float64).reshape(self.N))
            self.lib.nhc_set(self.h, self.NAMES.index(n), a.ctypes.data_as(C.c_void_p))
        self.lib.nhc_finalize(self.h)

    def time_step(self, dt):
        self.lib.nhc_step(self.h, float(dt))

    def get(self, name):
        a = np.empty(self.N, dtype=np.float64, order="F")
        self.lib.nhc_get(self.h, self.NAMES.index(name), a.ctypes.data_as(C.c_void_p))
        return a

    def __del__(self):
        try:
            self.lib.nhc_destroy(self.h)
        except Exception:
            pass
