#!/usr/bin/env python
"""bench.py — cell-updates/s per full time step of the NonhydrostaticModel hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c3f32|c4|c4s|c4l|c4cb|c1|c3u5|c2c4] [--impl ours|reference]

One "step" is one full `time_step!(model, Δt)` (RK3: three stages, each tendency+substep, halo fills, FFT
pressure solve and projection) over one synthetic, seeded initial state.

  value        whole-job cell-updates/s with the state resident in HBM (CUDA events on the library's stream)
  e2e          the same metric through the host API with HOST buffers: every step uploads the prognostic
               fields from pinned host memory (set!), steps, and downloads u, v, w, tracers and p (interior)
  roofline     the dominant kernel class (fused tendency+substep launches): algorithmic bytes / CUDA-event time
               against MEASURED_PEAKS.json's HBM copy bandwidth; `step` = whole-step algorithmic bytes
                / step time against the same peak
  cpu_baseline the oracle (NumPy port of the reference algorithm; oracle/) timed on this host on a bounded sample

`--impl reference` times the CPU restatement of the reference (the Julia reference cannot run here: no julia
binary) on the same physics at a bounded size.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cell-updates/sec per full RK3 step"
# measured DRAM traffic per MarchKernel launch (bytes) and FP64 instructions executed per cell: derived from the ncu capture of the SHIPPED
# kernels by the session that made the capture (profiles/ncu_march_traffic.json names the capture files) — not constants of this file
def _ncu_facts():
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_march_traffic.json")) as fh:
            d = json.load(fh)
    except Exception:
        return {}, {}, {}
    return ({k: v["dram_bytes_per_launch_mean"] for k, v in d.items()}, {k: v["fp64_instr_per_cell"]["mean"] for k, v in d.items()},
            {k: v["source"] for k, v in d.items()})


NCU_TRAFFIC, NCU_FP64_PER_CELL, NCU_SOURCE = _ncu_facts()

# name -> description of the BASELINE.json configuration
WORKLOADS = {
    "c1": dict(N=(128, 128, 1), topo="PPF", FT="f64", adv="weno", tracers=(), buoy=None, closure=None, F=3, b=0, a=0,
               label="C1 128^2 (P,P,Flat) WENO-5 F64"),
    "c2": dict(N=(256, 256, 256), topo="PPP", FT="f64", adv="centered", tracers=(), buoy=None, closure=None, F=3, b=0, a=0,
               label="C2 256^3 (P,P,P) Centered-2 F64 no tracers"),
    "c3": dict(N=(512, 512, 512), topo="PPP", FT="f64", adv="weno", tracers=("T", "S"), buoy="seawater", closure="scalar",
               F=5, b=1, a=0, label="C3 512^3 (P,P,P) WENO-5 T,S SeawaterBuoyancy ScalarDiffusivity F64"),
    "c3f32": dict(N=(512, 512, 512), topo="PPP", FT="f32", adv="weno", tracers=("T", "S"), buoy="seawater", closure="scalar",
                  F=5, b=1, a=0, label="C3 512^3 (P,P,P) WENO-5 T,S SeawaterBuoyancy ScalarDiffusivity F32"),
    "c4": dict(N=(512, 512, 256), topo="PPB", FT="f64", adv="weno", tracers=("T", "S"), buoy="seawater", closure="amd",
               F=5, b=1, a=1, label="C4 512^2x256 (P,P,B) WENO-5 AMD FPlane flux BCs F64"),
    # the C3 physics with the linear fifth-order upwind scheme / C2 with Centered(4) — general tile kernel
    "c3u5": dict(N=(512, 512, 512), topo="PPP", FT="f64", adv="upwind5", tracers=("T", "S"), buoy="seawater", closure="scalar",
                 F=5, b=1, a=0, label="C3 physics with UpwindBiased(order=5): 512^3 (P,P,P) T,S SeawaterBuoyancy ScalarDiffusivity F64"),
    "c2c4": dict(N=(256, 256, 256), topo="PPP", FT="f64", adv="centered4", tracers=(), buoy=None, closure=None, F=3, b=0, a=0,
                 label="C2 physics with Centered(order=4): 256^3 (P,P,P) F64 no tracers"),
    # the C4 physics on a vertically stretched (surface-refined) grid — FourierTridiagonalPoissonSolver
    "c4s": dict(N=(512, 512, 256), topo="PPB", FT="f64", adv="weno", tracers=("T", "S"), buoy="seawater", closure="amd",
                F=5, b=1, a=1, stretched=True,
                label="C4s 512^2x256 (P,P,B) stretched z, FourierTridiagonal solver, WENO-5 AMD FPlane flux BCs F64"),
    # the C4 physics with the closure of test/test_nonhydrostatic_regression.jl:68 instead of AMD
    "c4l": dict(N=(512, 512, 256), topo="PPB", FT="f64", adv="weno", tracers=("T", "S"), buoy="seawater", closure="lilly",
                F=5, b=1, a=1,
                label="C4 physics with (SmagorinskyLilly(C=0.23, Cb=1, Pr=1), ScalarDiffusivity): 512^2x256 (P,P,B) WENO-5 FPlane flux BCs F64"),
    # the C4 physics with AMD's buoyancy modification (AnisotropicMinimumDissipation(; Cb = 1), anisotropic_minimum_dissipation.jl:62-68)
    "c4cb": dict(N=(512, 512, 256), topo="PPB", FT="f64", adv="weno", tracers=("T", "S"), buoy="seawater", closure="amdcb",
                 F=5, b=1, a=1, label="C4 physics with AnisotropicMinimumDissipation(Cb=1): 512^2x256 (P,P,B) WENO-5 FPlane flux BCs F64"),
}
LES = ("amd", "lilly", "amdcb")       # the C4 family: Δ = 1 m, LES initial condition, Δt = 1 s, FPlane, surface flux BCs


def stretched_faces(Nz, Lz):
    """Surface-refined faces on [-Lz, 0] (spacing ratio 4:1 bottom:top), cf. examples/ocean_wind_mixing_and_convection.jl"""
    s = np.arange(Nz + 1, dtype=np.float64) / Nz
    return -Lz + Lz * (s + 0.6 * np.sin(np.pi * s) / np.pi)


def reals_per_cell_step(w):
    """10F + 3b + 9b + 3a(1+Nt) + 3a(4+2Nt) + 57 (2-D: Poisson 7 + projection 5 per stage)."""
    F, b, a = w["F"], w["b"], w["a"]
    Nt = F - 3
    if w["N"][2] == 1:
        return 10 * F + 3 * (7 + 5)
    return 10 * F + 3 * b + 9 * b + 3 * a * (1 + Nt) + 3 * a * (4 + 2 * Nt) + 57


def tendency_reals_per_cell_step(w):
    """K1 of SURVEY.md §8d summed over the 3 stages: read F + b + a(1+Nt) (+F G⁻ for s>1), write F (+F Gⁿ for s<3)."""
    F, b, a = w["F"], w["b"], w["a"]
    Nt = F - 3
    return 3 * (2 * F + b + a * (1 + Nt)) + 2 * F + 2 * F


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------- model construction
PX = 1      # --px: ranks along x (Partition(PX, world / PX)); 1 = slabs in y, what the driver's scaling run uses


def global_size(w, world):
    """Weak scaling: 2^27 cells per GPU for the 512^3 workloads; 8 GPUs = BASELINE's 1024^3 (slab-decomposed in y)."""
    if world == 1:
        return tuple(w["N"])
    if w["topo"] != "PPP" or w["N"] != (512, 512, 512):
        raise SystemExit("multi-GPU runs: the triply periodic 512^3-per-GPU workloads (c3, c3f32)")
    return {2: (1024, 512, 512), 4: (1024, 1024, 512), 8: (1024, 1024, 1024)}[world]


def build_model(w, device, rank=0, world=1):
    import oceananigans_b200 as ob
    FT = np.float64 if w["FT"] == "f64" else np.float32
    topo = {"P": ob.Periodic, "B": ob.Bounded, "F": ob.Flat}
    nonflat = [d for d in range(3) if w["topo"][d] != "F"]
    gsize = global_size(w, world)
    size = tuple(gsize[d] for d in nonflat)
    if w.get("closure") in LES:
        extent = tuple(float(w["N"][d]) for d in nonflat)                      # Δ = 1 m
    elif w["topo"] == "PPF":
        extent = (2 * np.pi, 2 * np.pi)
    else:
        extent = tuple(1.0 for _ in nonflat)
    arch = ob.B200(device) if world == 1 else ob.Distributed(ob.B200(device), partition=ob.Partition(PX, world // PX), rank=rank, nranks=world)
    if world > 1:
        extent = tuple(gsize[d] / 512.0 for d in nonflat)          # same Δ as the single-GPU workload
    if w.get("stretched"):
        grid = ob.RectilinearGrid(arch, FT, size=size, x=(0.0, extent[0]), y=(0.0, extent[1]),
                                  z=[float(v) for v in stretched_faces(w["N"][2], extent[2])], topology=tuple(topo[c] for c in w["topo"]))
    else:
        grid = ob.RectilinearGrid(arch, FT, size=size, extent=extent, topology=tuple(topo[c] for c in w["topo"]))
    adv = {"weno": ob.WENO, "centered": ob.Centered, "upwind5": lambda: ob.UpwindBiased(order=5),
           "centered4": lambda: ob.Centered(order=4)}[w["adv"]]()
    kw = dict(grid=grid, advection=adv, tracers=w["tracers"])
    if w["buoy"] == "seawater":
        if w["closure"] in LES:
            kw["buoyancy"] = ob.SeawaterBuoyancy(equation_of_state=ob.LinearEquationOfState(thermal_expansion=2e-4, haline_contraction=8e-4))
        else:
            kw["buoyancy"] = ob.SeawaterBuoyancy()
    if w["closure"] == "scalar":
        kw["closure"] = ob.ScalarDiffusivity(nu=1e-5, kappa=1e-5)
    elif w["closure"] in LES:
        kw["closure"] = ob.AnisotropicMinimumDissipation() if w["closure"] == "amd" else \
            ob.AnisotropicMinimumDissipation(Cb=1.0) if w["closure"] == "amdcb" else \
            (ob.SmagorinskyLilly(C=0.23, Cb=1.0, Pr=1.0), ob.ScalarDiffusivity(nu=1.05e-6, kappa=1.46e-7))
        kw["coriolis"] = ob.FPlane(f=1e-4)
        kw["boundary_conditions"] = {      # test/regression_tests/ocean_large_eddy_simulation_regression_test.jl:19-37
            "u": ob.FieldBoundaryConditions(top=ob.FluxBoundaryCondition(-2e-5)),
            "T": ob.FieldBoundaryConditions(top=ob.FluxBoundaryCondition(5e-5), bottom=ob.GradientBoundaryCondition(0.005)),
            "S": ob.FieldBoundaryConditions(top=ob.FluxBoundaryCondition(5e-8)),
        }
    model = ob.NonhydrostaticModel(**kw)
    return ob, model


def synthetic_ic(w, model, seed=1234):
    """rng(1234); u,v,w ~ U(-1,1); T = 20 + 0.01 N(0,1); S = 35 + 0.01 N(0,1) (C4: LES-like profile)."""
    rng = np.random.default_rng(seed)
    FT = model.grid.FT
    ic = {}
    for n in ("u", "v", "w"):
        shape = tuple(model.fields[n].info().interior_size)
        if w["closure"] in LES:
            ic[n] = (1e-3 * rng.standard_normal(shape, dtype=np.float32)).astype(FT)
        else:
            ic[n] = rng.uniform(-1, 1, shape).astype(FT) if FT is np.float64 else (2 * rng.random(shape, dtype=np.float32) - 1)
    for n in w["tracers"]:
        shape = tuple(model.fields[n].info().interior_size)
        base = {"T": 20.0, "S": 35.0}.get(n, 0.0)
        ic[n] = (base + 0.01 * rng.standard_normal(shape, dtype=np.float32)).astype(FT)
    return ic


def default_dt(w):
    if w["closure"] in LES:
        return 1.0
    if w["topo"] == "PPF":
        return 0.01
    return 0.1 / w["N"][0]


# ----------------------------------------------------------------------------------------------- correctness next to the number
def parity_check(w, local, rank, world, steps=2, library=None):
    """Before anything is timed: the benchmark's physics on a small global grid, 64 x (16 N) x 64, decomposed over the N ranks exactly like
    the timed problem (slab-y, NCCL halo exchange, transposed distributed FFT), against the single-domain CPU restatement of the reference
    (oracle/nhm_step.c — the checker, never the thing measured) after `steps` RK3 steps.  Relative L-inf over u, v, w, p, T, S (normalised
    by the global maximum), max over ranks.  `worst` is against the twin in the reference's order of operations; `worst_difference_form`
    against the twin with the WENO smoothness indicators in the well-conditioned difference form the kernel uses (the reference's expanded
    form carries its own ~1e-12 … 1e-8 round-off on T, S: tests/test_oracle_c.py).  Triply periodic workloads only (the twin's scope)."""
    if w["topo"] != "PPP" or w["adv"] not in ("weno", "centered") or w["closure"] in LES:
        return None
    import oceananigans_b200 as ob
    from oracle.c_twin import CTwin
    FT = np.float64 if w["FT"] == "f64" else np.float32
    N = (64, 16 * world, 64)
    extent = tuple(n / 64.0 for n in N)
    arch = ob.B200(local) if world == 1 else ob.Distributed(ob.B200(local), partition=ob.Partition(PX, world // PX), rank=rank, nranks=world)
    grid = ob.RectilinearGrid(arch, FT, size=N, extent=extent, topology=(ob.Periodic, ob.Periodic, ob.Periodic))
    kw = dict(grid=grid, advection=ob.WENO() if w["adv"] == "weno" else ob.Centered(), tracers=w["tracers"])
    if w["buoy"] == "seawater":
        kw["buoyancy"] = ob.SeawaterBuoyancy()
    nu = 1e-5 if w["closure"] == "scalar" else 0.0
    if w["closure"] == "scalar":
        kw["closure"] = ob.ScalarDiffusivity(nu=nu, kappa=nu)
    if library is not None:          # tests/test_bench_contract.py: the host simulation of the kernel sources (CPU-only test of this function)
        kw["library"] = library
    model = ob.NonhydrostaticModel(**kw)
    rng = np.random.default_rng(4321)
    ic = {n: rng.uniform(-1, 1, N) for n in ("u", "v", "w")}
    for n in w["tracers"]:
        ic[n] = {"T": 20.0, "S": 35.0}.get(n, 0.0) + 0.01 * rng.standard_normal(N)
    ic = {n: a.astype(FT).astype(np.float64) for n, a in ic.items()}
    if world > 1:
        (i0, ni), (j0, nj) = arch.local_range(0, N[0]), arch.local_range(1, N[1])
    else:
        (i0, ni), (j0, nj) = (0, N[0]), (0, N[1])
    sl = (slice(i0, i0 + ni), slice(j0, j0 + nj), slice(None))
    ob.set_(model, **{n: a[sl] for n, a in ic.items()})
    twins = {"worst": CTwin(N, extent, weno=w["adv"] == "weno", tracers=bool(w["tracers"]), nu=nu, kappa=nu)}
    if w["adv"] == "weno" and w["tracers"]:
        twins["worst_difference_form"] = CTwin(N, extent, weno=True, tracers=True, nu=nu, kappa=nu, beta_difference_form=True)
    dt = 0.1 / 64
    for ct in twins.values():
        ct.set(**ic)
    for _ in range(steps):
        ob.time_step_(model, dt)
        for ct in twins.values():
            ct.time_step(dt)
    out = {}
    for key, ct in twins.items():
        worst = 0.0
        for n in tuple(ic) + ("p",):
            ref = ct.get(n)
            got = (model.pressures.pNHS if n == "p" else model.fields[n]).interior().astype(np.float64)
            worst = max(worst, float(np.abs(got - ref[sl]).max() / np.abs(ref).max()))
        out[key] = worst
    del model
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([out.get("worst", 0.0), out.get("worst_difference_form", 0.0)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = {k: float(t[i].item()) for i, k in enumerate(("worst", "worst_difference_form")) if k in out}
    tol = 1e-11 if FT is np.float64 else 1e-4
    out.update({"tol": tol, "ok": bool(out["worst"] <= tol), "grid": list(N), "steps": steps, "ranks": world,
                "against": "oracle/nhm_step.c (C99 restatement of the reference algorithm) on the undecomposed grid, u, v, w, p and tracers"})
    return out


def time_workload(name, local, steps=5, warmup=3):
    """One more BASELINE configuration, timed on the device like the main line (state resident in HBM, CUDA events on the library's
    stream), so that the driver's record carries configs 2-4 too."""
    w = WORKLOADS[name]
    ob, model = build_model(w, local)
    lib, h = model._lib, model._h
    ob.set_(model, **synthetic_ic(w, model))
    dt = default_dt(w)
    for _ in range(warmup):
        ob.time_step_(model, dt)
    model.sync()
    ms = C.c_double()
    lib.check(lib.oc_stopwatch_start(h))
    for _ in range(steps):
        ob.time_step_(model, dt)
    lib.check(lib.oc_stopwatch_stop(h, C.byref(ms)))
    model.timers(enable=True)
    model.timers(reset=True)
    for _ in range(steps):
        ob.time_step_(model, dt)
    timers = model.timers()
    model.timers(enable=False)
    if not np.isfinite(float(np.abs(model.velocities.u.interior()).max())):
        raise SystemExit(f"{name}: state became non-finite")
    cells = int(np.prod(w["N"]))
    itemsize = 8 if w["FT"] == "f64" else 4
    ms_per_step = ms.value / steps
    peak, _ = peaks()
    step_gbs = cells * reals_per_cell_step(w) * itemsize / (ms_per_step * 1e-3) / 1e9
    del model
    return {"workload": w["label"], "dtype": w["FT"], "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "value": cells / (ms_per_step * 1e-3), "unit": "cell-updates/s",
            "step_roofline": {"achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                              "reals_per_cell_step": reals_per_cell_step(w)},
            "kernel_ms_per_step": {k: v[0] / steps for k, v in timers.items() if v[1]}}


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    if world > 1:            # torch is plumbing for the multi-rank runs only (rendezvous, NCCL id broadcast, max over ranks)
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOADS[args.workload]
    parity = None if args.no_parity_check else parity_check(w, local, rank, world)
    ob, model = build_model(w, local, rank, world)
    lib, h = model._lib, model._h
    ic = synthetic_ic(w, model, seed=1234 + rank)
    ob.set_(model, **ic)
    names = ("u", "v", "w") + tuple(w["tracers"])
    dt = default_dt(w)
    gsz = global_size(w, world)
    cells = int(np.prod(gsz)) // world                      # cells per GPU
    itemsize = 8 if w["FT"] == "f64" else 4

    def barrier():
        if world > 1:
            dist.barrier()
        model.sync()

    # warm-up
    for _ in range(args.warmup):
        ob.time_step_(model, dt)
    barrier()
    launches0 = model.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = C.c_double()
    barrier()
    lib.check(lib.oc_stopwatch_start(h))
    for _ in range(args.steps):
        ob.time_step_(model, dt)
    lib.check(lib.oc_stopwatch_stop(h, C.byref(ms)))
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    elapsed_ms = ms.value
    launches = model.launch_count() - launches0
    # per-kernel-class breakdown: a separate short pass with the per-launch CUDA-event timers on (they are off in the timed region: ~2
    # event records per launch, and a model with timers on does not replay captured graphs)
    timer_steps = max(1, min(args.steps, 5))
    model.timers(enable=True)
    model.timers(reset=True)
    for _ in range(timer_steps):
        ob.time_step_(model, dt)
    timers = model.timers()
    model.timers(enable=False)
    timers = {k: (v[0] * args.steps / timer_steps, v[1] * args.steps / timer_steps) for k, v in timers.items()}     # scaled to `steps`
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * cells / (ms_per_step * 1e-3)

    # sanity: the state must still be finite (a NaN run is not a measurement)
    umax = float(np.abs(model.velocities.u.interior()).max())
    if not np.isfinite(umax):
        raise SystemExit("state became non-finite during the benchmark")

    # ---- e2e: host buffers in, host buffers out, every step -------------------------------------------
    # The call sequence a host-driven user makes through the C ABI: per step, set! every prognostic field from page-locked host arrays,
    # time_step!, fetch u, v, w, tracers and p into page-locked host arrays.  Two variants are timed:
    #   pipelined (the `e2e` value): oc_upload_begin / oc_output_begin — every copy is asynchronous and stream-ordered, two staging sets,
    #       so the H2D copy of step n+1's inputs and the D2H copy of step n-1's results overlap step n's kernels (PCIe is full duplex);
    #       the host waits only for the tickets it is about to reuse.  All copies of all steps lie inside the timed region (the loop ends
    #       with a drain of every ticket).
    #   serial (`e2e.serial`): oc_upload_interior / oc_download_interior, each of which blocks the host — round 1's figure.
    e2e = None
    if not args.no_e2e:
        bufs = {}
        for n in names + ("pNHS",):
            f = model.pressures.pNHS if n == "pNHS" else model.fields[n]
            shape = tuple(f.info().interior_size)
            nbytes = int(np.prod(shape)) * itemsize
            ptrs = []
            for _ in range(2):               # two host buffers per field: the results of step n-1 are still being written while step n runs
                p = C.c_void_p()
                lib.check(lib.oc_host_alloc(C.byref(p), nbytes))
                ptrs.append(p)
            bufs[n] = (ptrs, nbytes, f)
        for n in names:                      # host copy of the current state = every step's input
            ptrs, nbytes, f = bufs[n]
            lib.check(lib.oc_download_interior(h, f.id, ptrs[0], nbytes))
            C.memmove(ptrs[1], ptrs[0], nbytes)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        h2d = sum(bufs[n][1] for n in names)
        d2h = sum(bufs[n][1] for n in names + ("pNHS",))
        lo3, tk = (C.c_int * 3)(0, 0, 0), C.c_int()

        def wait_all(tickets):
            for t in tickets:
                lib.check(lib.oc_output_wait(h, t))

        def pipelined(nsteps):
            ins, outs = [[], []], [[], []]                  # tickets in flight per staging set
            for s in range(nsteps):
                a = s % 2
                wait_all(ins[a]); ins[a] = []               # set a's staging buffers / host inputs (step s-2) are free again
                for n in names:                             # set!(model; u=…, v=…, …) from pinned host memory
                    ptrs, nbytes, f = bufs[n]
                    lib.check(lib.oc_upload_begin(h, f.id, ptrs[a], nbytes, C.byref(tk)))
                    ins[a].append(tk.value)
                lib.check(lib.oc_set_finalize(h, 0))
                ob.time_step_(model, dt)
                wait_all(outs[a]); outs[a] = []             # the host arrays of set a hold step s-2's results: consumed, reusable
                for n in names + ("pNHS",):                 # Array(interior(field))
                    ptrs, nbytes, f = bufs[n]
                    sz = (C.c_int * 3)(*f.info().interior_size)
                    lib.check(lib.oc_output_begin(h, f.id, lo3, sz, ptrs[a], nbytes, C.byref(tk)))
                    outs[a].append(tk.value)
            for a in (0, 1):
                wait_all(ins[a]); wait_all(outs[a])

        def serial(nsteps):
            for _ in range(nsteps):
                for n in names:
                    ptrs, nbytes, f = bufs[n]
                    lib.check(lib.oc_upload_interior(h, f.id, ptrs[0], nbytes))
                lib.check(lib.oc_set_finalize(h, 0))
                ob.time_step_(model, dt)
                for n in names + ("pNHS",):
                    ptrs, nbytes, f = bufs[n]
                    lib.check(lib.oc_download_interior(h, f.id, ptrs[0], nbytes))

        def timed(fn, nsteps):
            barrier()
            t0 = time.perf_counter()
            fn(nsteps)
            barrier()
            sec = (time.perf_counter() - t0) / nsteps
            if world > 1:
                t = torch.tensor([sec], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
            return sec

        pipelined(2)                                        # allocates the staging buffers (not part of any step)
        pipe_steps = max(e2e_steps, 20)                     # the pipeline's fill and drain are inside the timed region: amortise them
        e2e_s = timed(pipelined, pipe_steps)
        serial_s = timed(serial, e2e_steps)
        e2e = {"value": world * cells / e2e_s, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s * 1e3, "steps": pipe_steps,
               "what": "per step: set! u,v,w,tracers from pinned host arrays (oc_upload_begin), time_step!, fetch u,v,w,tracers,pNHS into pinned "
                       "host arrays (oc_output_begin); copies are stream-ordered and overlap the neighbouring steps' kernels (two staging sets); "
                       "timed from the first upload to the last completed download",
               "pcie_gbs": (h2d + d2h) / e2e_s / 1e9,
               "serial": {"value": world * cells / serial_s, "ms_per_step": serial_s * 1e3, "steps": e2e_steps,
                          "what": "the same with the blocking oc_upload_interior / oc_download_interior (round 1's e2e)"}}
        for n in bufs:
            for p in bufs[n][0]:
                lib.oc_host_free(p)

    device_bytes = model.device_bytes()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # the other single-GPU BASELINE configurations, device-timed like the main line (sequentially, after the main model is gone)
    others = None
    if world == 1 and not args.no_other_workloads and args.workload == "c3":
        del model
        others = {n: time_workload(n, local, steps=200 if n == "c1" else 5, warmup=6 if n == "c1" else 3) for n in ("c2", "c3f32", "c4", "c1")}

    peak, peak_src = peaks()
    tend_ms, tend_n = timers["tendency"]
    tend_bytes_step = cells * tendency_reals_per_cell_step(w) * itemsize
    launches_per_step = tend_n / args.steps if args.steps else 0
    bytes_per_launch = tend_bytes_step / launches_per_step if launches_per_step else 0
    avg_launch_ms = tend_ms / tend_n if tend_n else float("nan")
    achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if tend_n else 0.0
    step_bytes = cells * reals_per_cell_step(w) * itemsize
    step_gbs = step_bytes / (ms_per_step * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": w["FT"], "data": "synthetic (seeded rng 1234, SURVEY.md 8d)",
        "config": {"workload": w["label"] if world == 1 else w["label"].replace("512^3", "x".join(map(str, gsz)) + (" (slab-y over %d GPUs)" % world if PX == 1 else " (%d x %d pencils)" % (PX, world // PX))),
                   "grid": list(gsz), "topology": w["topo"], "timestepper": "RungeKutta3", "dt": dt,
                   "cells_per_gpu": cells, "l2_policy": "working set (>25 GB at 512^3) far exceeds the 126 MB L2; no flush needed"
                   if cells >= 256 ** 3 else "working set may fit L2 (launch-latency configuration)"},
        "roofline": {"bound": "hbm", "kernel": "TendencyKernel (fused tendency + RK3 substep, one launch per prognostic field)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(args.workload if world == 1 else None),
                     "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 5 launches of one stage: " + NCU_SOURCE[args.workload]) if (world == 1 and args.workload in NCU_TRAFFIC) else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "avg_launch_ms": avg_launch_ms, "launches_per_step": launches_per_step,
                     "step": {"algorithmic_bytes": step_bytes, "achieved": step_gbs, "frac": step_gbs / peak,
                              "reals_per_cell_step": reals_per_cell_step(w)}},
        # WENO-5 in Float64 is bound by the FP64 pipe, not by HBM: FP64 warp-instructions issued per second by the
        # MarchKernel launches against B200's 64 FP64 lanes / clk / SM (148 SMs at the sampled SM clock)
        "fp64_pipe": ({"fp64_instr_per_cell_per_launch": NCU_FP64_PER_CELL[args.workload],
                       "achieved_lane_ops_per_s": NCU_FP64_PER_CELL[args.workload] * cells / (avg_launch_ms * 1e-3),
                       "peak_lane_ops_per_s": 64 * 148 * (clocks.get("sm_mhz") or 1965.0) * 1e6,
                       "frac": NCU_FP64_PER_CELL[args.workload] * cells / (avg_launch_ms * 1e-3) / (64 * 148 * (clocks.get("sm_mhz") or 1965.0) * 1e6),
                       "source": "instruction count from the ncu source page (profiles/ncu_march_traffic.json); time from this run's CUDA events"}
                      if (world == 1 and args.workload in NCU_FP64_PER_CELL and tend_n) else None),
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timers.items()},
        "kernel_launches_per_step": {k: v[1] / args.steps for k, v in timers.items()},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": e2e,
        "device_bytes": device_bytes,
        "parity_check": parity,
        "other_workloads": others,
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(w, args.cpu_size)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- CPU legs
def oracle_model(w, N):
    import oracle
    from oracle import advection as adv, closures as clo
    from oracle.grid import BC
    FT = np.float64 if w["FT"] == "f64" else np.float32
    nonflat = [d for d in range(3) if w["topo"][d] != "F"]
    size = tuple(N[d] for d in nonflat)
    extent = tuple(float(N[d]) for d in nonflat) if w["closure"] in LES else tuple(1.0 for _ in nonflat)
    if w.get("stretched"):
        og = oracle.Grid(FT, size=size, x=(0.0, extent[0]), y=(0.0, extent[1]),
                         z=[float(v) for v in stretched_faces(N[2], extent[2])], topology=tuple(w["topo"]))
    else:
        og = oracle.Grid(FT, size=size, extent=extent, topology=tuple(w["topo"]))
    kw = dict(advection={"weno": lambda: adv.WENO(FT, 5), "centered": lambda: adv.Centered(FT, 2),
                         "upwind5": lambda: adv.UpwindBiased(FT, 5), "centered4": lambda: adv.Centered(FT, 4)}[w["adv"]](),
              tracers=w["tracers"])
    if w["buoy"] == "seawater":
        kw["buoyancy"] = clo.SeawaterBuoyancy()
    if w["closure"] == "scalar":
        kw["closure"] = clo.ScalarDiffusivity(1e-5, 1e-5)
    elif w["closure"] in LES:
        kw["closure"] = clo.AnisotropicMinimumDissipation() if w["closure"] == "amd" else \
            clo.AnisotropicMinimumDissipation(Cb=1.0) if w["closure"] == "amdcb" else \
            (clo.SmagorinskyLilly(0.23, 1.0, 1.0), clo.ScalarDiffusivity(1.05e-6, 1.46e-7))
        kw["coriolis_f"] = 1e-4
        kw["boundary_conditions"] = {"u": {"top": BC("flux", -2e-5)}, "T": {"top": BC("flux", 5e-5), "bottom": BC("gradient", 0.005)},
                                     "S": {"top": BC("flux", 5e-8)}}
    om = oracle.OracleModel(og, **kw)
    rng = np.random.default_rng(1234)
    if w["closure"] in LES:          # the LES workloads step with Δt = 1 s at Δ = 1 m: small velocities, like synthetic_ic
        ic = {n: 1e-3 * rng.standard_normal(om.fields[n].interior.shape) for n in ("u", "v", "w")}
    else:
        ic = {n: rng.uniform(-1, 1, om.fields[n].interior.shape) for n in ("u", "v", "w")}
    for n in w["tracers"]:
        ic[n] = {"T": 20.0, "S": 35.0}.get(n, 0.0) + 0.01 * rng.standard_normal(om.fields[n].interior.shape)
    om.set(**ic)
    return om


def cpu_sample_size(w, n):
    return tuple(1 if w["N"][d] == 1 else min(w["N"][d], n) for d in range(3))


def c_twin_model(w, n):
    """The C99 + OpenMP twin of the oracle (oracle/nhm_step.c) — triply periodic physics only (c2, c3)."""
    from oracle.c_twin import CTwin
    N = (n, n, n)
    ct = CTwin(N, (1.0, 1.0, 1.0), weno=w["adv"] == "weno", tracers=bool(w["tracers"]),
               nu=1e-5 if w["closure"] == "scalar" else 0.0, kappa=1e-5 if w["closure"] == "scalar" else 0.0)
    rng = np.random.default_rng(1234)
    ic = {name: rng.uniform(-1, 1, N) for name in ("u", "v", "w")}
    if w["tracers"]:
        ic["T"] = 20.0 + 0.01 * rng.standard_normal(N)
        ic["S"] = 35.0 + 0.01 * rng.standard_normal(N)
    ct.set(**ic)
    return ct, N


def cpu_run(w, n_numpy, max_steps, budget_s, warmup=1):
    """Time the CPU restatement of the reference algorithm on a bounded sample of the workload: the multi-threaded C twin for the
    triply periodic workloads (Float64 arithmetic), the NumPy oracle otherwise.  Returns (cells/s, s/step, steps, cores, sample)."""
    if w["topo"] == "PPP" and w["closure"] not in LES and w["adv"] in ("weno", "centered"):
        n = 128
        model, N = c_twin_model(w, n)
        dt = 0.1 / n
        cores = int(os.environ.get("OMP_NUM_THREADS", 0)) or (os.cpu_count() or 1)
        what = f"C99+OpenMP twin of the oracle (oracle/nhm_step.c), {cores} threads, Float64"
    else:
        N = cpu_sample_size(w, n_numpy)
        model = oracle_model(w, N)
        dt = default_dt(w) * w["N"][0] / N[0] if w["closure"] not in LES else 1.0
        cores = 1
        what = "NumPy oracle, 1 thread"
    for _ in range(warmup):
        model.time_step(dt)
    t0 = time.perf_counter()
    steps = 0
    while True:
        model.time_step(dt)
        steps += 1
        if time.perf_counter() - t0 > budget_s or steps >= max_steps:
            break
    s = (time.perf_counter() - t0) / steps
    sample = (f"{steps} RK3 step(s) of the same physics at {N[0]}x{N[1]}x{N[2]} ({s:.2f} s/step), {what}; "
              "the Julia reference cannot run here (no julia binary)")
    return int(np.prod(N)) / s, s, steps, cores, sample, N


def cpu_baseline(w, n):
    v, s, steps, cores, sample, N = cpu_run(w, n, max_steps=40, budget_s=10.0)
    return {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's algorithm (oracle/) on all the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)          # torchrun pins it to 1: the reference arm may use every core
    w = WORKLOADS[args.workload]
    v, s, steps, cores, sample, N = cpu_run(w, args.cpu_size, max_steps=max(1, args.steps), budget_s=60.0, warmup=args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (seeded rng 1234)", "config": {"workload": w["label"], "sample_grid": list(N)},
        "cpu_baseline": {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--px", type=int, default=1, help="ranks along x: Partition(px, gpus / px); default 1 = slabs in y")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
    ap.add_argument("--cpu-size", type=int, default=48, help="edge of the bounded CPU sample")
    args = ap.parse_args()
    global PX
    PX = max(1, args.px)
    if args.gpus % PX != 0:
        raise SystemExit("--px must divide --gpus")
    # The contract is ONE JSON line on stdout.  Libraries write banners to file descriptor 1 behind Python's back ("NCCL version …" at
    # communicator creation), so everything but the result line is routed to stderr: fd 1 points to fd 2 while the benchmark runs,
    # and print() is given the real stdout for the JSON line only.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    builtins_print = print

    def emit(*a, **k):
        builtins_print(*a, file=real_stdout, **k)
        real_stdout.flush()

    globals()["print"] = emit
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()


if __name__ == "__main__":
    main()
