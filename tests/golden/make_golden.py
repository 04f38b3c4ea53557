"""Generates tests/golden/*.npz from the CPU oracle (oracle/), the restatement of the reference algorithm.

    python tests/golden/make_golden.py

The Julia reference cannot run in the build image (no julia binary) and its own regression files are downloaded at
test time (test/data_dependencies.jl:17-38), so these vectors freeze the ORACLE's output: they guard the oracle against
drift (tests/test_golden.py, CPU) and give the CUDA path (-m gpu) and the host simulation a fixed target.
Each file holds the seeded initial condition (after set!'s projection) and u, v, w, p, tracers after 1 and 3 steps.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN_CASES = {
    "ppp_weno_ts_f64": dict(N=(12, 10, 8), topo="PPP", scheme="weno"),
    "ppb_weno_amd_fplane_bcs_f64": dict(N=(12, 10, 8), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True),
    "ppp_centered_f64": dict(N=(12, 10, 8), topo="PPP", scheme="centered", closure="none", buoy="none"),
    "ppf_weno_2d_f64": dict(N=(12, 10, 1), topo="PPF", scheme="weno", closure="none", buoy="none"),
    "ppb_weno_ab2_f64": dict(N=(12, 10, 8), topo="PPB", scheme="weno", ts="QuasiAdamsBashforth2"),
    "ppp_weno_ts_f32": dict(N=(12, 10, 8), topo="PPP", scheme="weno", FT=np.float32),
    # vertically stretched grids (FourierTridiagonalPoissonSolver)
    "stretched_ppb_weno_amd_fplane_bcs_f64": dict(N=(12, 10, 8), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, stretch="smooth"),
    "stretched_bbb_centered_scalar_f64": dict(N=(12, 10, 8), topo="BBB", scheme="centered", stretch="facr"),
}
# the round-1 widening: Smagorinsky-Lilly, the Coriolis family, tilted gravity, array-valued flux BCs.  Kept in a
# separate table because their GPU checks live in tests/test_widening_gpu.py (they run after the GPU-verified suites).
GOLDEN_CASES_WIDENING = {
    "ppb_weno_smagorinsky_lilly_fplane_bcs_f64": dict(N=(12, 10, 8), topo="PPB", scheme="weno", closure="lilly", f=1e-2, bcs=True),
    "bbb_centered_cartesian_coriolis_tilted_gravity_f64": dict(N=(12, 10, 8), topo="BBB", scheme="centered", buoy="tracer",
                                                               f=("cartesian", 0.3, -0.5, 0.7), tilt=(0.48, -0.6, -0.64)),
    "ppb_weno_betaplane_array_bcs_f64": dict(N=(12, 10, 8), topo="PPB", scheme="weno", f=("beta", 0.3, 2.0), bcs="array"),
    # AMD with the buoyancy modification (Cb = 1) under a NonTraditionalBetaPlane, on a stretched grid (ABI v4 features)
    "stretched_ppb_weno_amd_cb_nontraditional_betaplane_bcs_f64": dict(N=(12, 10, 8), topo="PPB", scheme="weno", closure="amdcb", bcs=True,
                                                                       f=("ntbeta", 0.7, -0.5, 2.0, 1.5, 3.0), stretch="smooth"),
}
STEPS = (1, 3)




def main(only=None):
    import oracle
    from oracle import advection as adv, closures as clo
    import parity_harness as ph
    for name, kw in {**GOLDEN_CASES, **GOLDEN_CASES_WIDENING}.items():
        if only and name not in only:
            continue
        om = ph.build_oracle(**kw)
        ic = ph.initial_conditions(om)
        om.set(**ic)
        dmin = [float(om.grid.D[d]) for d in range(3) if not om.grid.flat(d) and om.grid.D[d] is not None]
        if om.grid.stretched:
            dmin.append(float(np.min(om.grid.dz_at("c", np.arange(1, om.grid.Nz + 1)))))
        dt = 0.1 * min(dmin)
        out = {"dt": np.float64(dt)}
        for n, a in ic.items():
            out["ic_" + n] = a
        for s in range(1, max(STEPS) + 1):
            om.time_step(dt)
            if s in STEPS:
                for n in om.fields:
                    out[f"s{s}_{n}"] = om.fields[n].interior.copy()
                out[f"s{s}_p"] = om.pNHS.interior.copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name, {k: v.shape for k, v in out.items() if k.startswith("s1_")})


if __name__ == "__main__":
    main(only=sys.argv[1:])      # no arguments: regenerate every file
