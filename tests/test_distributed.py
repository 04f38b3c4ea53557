"""The N > 1 path on CPU: world_size-2 (and 4) gloo processes, each running the host simulation of the kernels on its y-slab
with the same decomposition / pack / exchange / transposed-FFT logic the CUDA library uses with NCCL.  The slabs must agree
with the single-domain oracle to the north_star tolerance (cf. test/test_distributed_poisson_solvers.jl:70-89 and
test/test_distributed_models.jl:334-407,518-536 of the reference, which also run 4 ranks on one node)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_ranks(R, case, timeout=600, backend="gloo"):
    import __graft_entry__ as ge
    if backend == "gloo" and not os.path.exists(ge.HOSTSIM):
        ge.build()
    port = _free_port()
    procs = []
    for r in range(R):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(R), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1",
                   OC_DIST_BACKEND=backend)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), json.dumps(case)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=timeout) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    return json.loads(outs[0][0].strip().splitlines()[-1])


CPU_MODEL_CASES = [
    (2, dict(N=(16, 12, 8), topo="PPP", scheme="weno", steps=2)),
    (2, dict(N=(16, 12, 8), topo="PPP", scheme="centered", f=1e-2, steps=2)),
    (4, dict(N=(12, 16, 8), topo="PPP", scheme="weno", steps=1)),
    (2, dict(N=(16, 12, 8), topo="PPP", scheme="weno", steps=2, f32=True)),
    # the LES topology: Bounded z (DCT twiddles around the transposed stage), AMD, FPlane, flux / gradient / value BCs
    (2, dict(N=(16, 12, 8), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, steps=2)),
    (2, dict(N=(12, 8, 6), topo="PPB", scheme="centered", steps=2, ts="QuasiAdamsBashforth2")),
    # the round-1 widening on slabs: Smagorinsky(-Lilly) eddy viscosities need their own halo exchange; BetaPlane needs the rank's y offset
    (2, dict(N=(16, 12, 8), topo="PPB", scheme="weno", closure="lilly", f=("beta", 0.3, 2.0), bcs=True, steps=2)),
    (4, dict(N=(12, 16, 8), topo="PPP", scheme="upwind3", closure="smag", f=("beta", 0.3, 2.0), steps=1)),
    # the three-component Coriolis forms on slabs: w and v halos across the slab boundary; NonTraditionalBetaPlane needs the rank's y offset
    (2, dict(N=(16, 12, 8), topo="PPB", scheme="weno", f=("cartesian", 0.3, -0.5, 0.7), bcs=True, steps=2)),
    (2, dict(N=(16, 12, 8), topo="PPP", scheme="centered", f=("ntbeta", 0.7, -0.5, 2.0, 1.5, 3.0), steps=2)),
    (4, dict(N=(12, 16, 8), topo="PPB", scheme="weno", closure="amd", f=("ntbeta", 0.7, -0.5, 2.0, 1.5, 3.0), steps=1)),
    # Bounded x (whole on every slab: a complex (z, x) stage with the DCT's twiddles) and Bounded y (walls on the outer ranks only — the
    # RightConnected / LeftConnected local grids of distributed_grids.jl:75-126 — an open chain of halo messages, the y DCT in the
    # transposed layout), with non-default boundary conditions on every wall; 3 and 4 ranks have fully connected slabs in between
    (2, dict(N=(16, 12, 8), topo="BPB", scheme="weno", bcs="walls", steps=2)),
    (2, dict(N=(16, 12, 8), topo="PBB", scheme="weno", closure="amd", f=1e-2, bcs="walls", steps=2)),
    (2, dict(N=(16, 12, 8), topo="BBB", scheme="weno", closure="lilly", f=("beta", 0.3, 2.0), bcs="walls", steps=2)),
    (3, dict(N=(12, 18, 9), topo="BBB", scheme="weno", bcs="walls", steps=2)),
    (4, dict(N=(12, 16, 8), topo="BBB", scheme="centered", closure="amd", f=("cartesian", 0.3, -0.5, 0.7), bcs="walls", steps=2, ts="QuasiAdamsBashforth2")),
    (4, dict(N=(12, 16, 8), topo="PBB", scheme="upwind3", closure="smag", f=("ntbeta", 0.7, -0.5, 2.0, 1.5, 3.0), bcs="walls", steps=1)),
    (2, dict(N=(16, 12, 8), topo="PBP", scheme="weno", bcs="walls", f=1e-2, steps=2)),
    (2, dict(N=(16, 12, 8), topo="BPB", scheme="weno", bcs="walls", steps=2, f32=True)),
    # pencils, Partition(px, R / px) (distributed_architectures.jl:242-302): halo exchange along x then y (corners through the two hops),
    # the two-transpose solve (z <-> y among a column of ranks, y <-> x among a row), ϕ's neighbour row AND column in the projection;
    # px = R is the slab decomposition in x
    (4, dict(N=(16, 12, 8), topo="PPP", scheme="weno", steps=2, px=2)),
    (4, dict(N=(16, 12, 8), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, steps=2, px=2)),
    (4, dict(N=(16, 12, 8), topo="BBB", scheme="weno", closure="lilly", f=("beta", 0.3, 2.0), bcs="walls", steps=2, px=2)),
    (2, dict(N=(16, 12, 8), topo="BPB", scheme="weno", bcs="walls", steps=2, px=2)),
    (6, dict(N=(18, 12, 6), topo="PBB", scheme="centered", closure="amd", f=("cartesian", 0.3, -0.5, 0.7), bcs="walls", steps=2, ts="QuasiAdamsBashforth2", px=3)),
    (6, dict(N=(12, 18, 9), topo="BBB", scheme="upwind3", closure="smag", f=("ntbeta", 0.7, -0.5, 2.0, 1.5, 3.0), bcs="walls", steps=1, px=2)),
    (4, dict(N=(16, 12, 8), topo="BPP", scheme="weno", steps=2, px=2, f32=True)),
    # slabs tall enough for the interior / boundary-strip split of the tendency launches (three or more tile rows of 8 or 16): on the host
    # simulation the deferred exchange is synchronous but the launches are partitioned exactly as on the GPU
    (2, dict(N=(16, 64, 8), topo="PBB", scheme="weno", bcs="walls", steps=2)),
    (2, dict(N=(16, 96, 8), topo="PPP", scheme="weno", steps=2)),
    # two-dimensional models: (x, y) on slabs in x, (y, z) on slabs in y (a Flat dimension cannot be partitioned)
    (2, dict(N=(16, 12, 1), topo="PPF", scheme="weno", closure="none", buoy="none", steps=2, px=2)),
    (4, dict(N=(16, 12, 1), topo="BPF", scheme="centered", closure="scalar", buoy="none", steps=2, px=4)),
    (2, dict(N=(1, 12, 8), topo="FPB", scheme="weno", buoy="tracer", f=0.2, steps=2)),
    (4, dict(N=(1, 16, 8), topo="FBB", scheme="upwind3", buoy="tracer", steps=2)),
    # vertically stretched grids: the distributed Fourier-tridiagonal solve (distributed_fft_tridiagonal_solver.jl:262-293) on slabs and pencils
    (2, dict(N=(16, 12, 10), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, stretch="smooth", steps=2)),
    (4, dict(N=(12, 16, 8), topo="BBB", scheme="centered", closure="both", bcs="walls", stretch="facr", steps=2)),
    (2, dict(N=(16, 12, 10), topo="PPB", scheme="weno", closure="lilly", bcs=True, stretch="smooth", steps=2, ts="QuasiAdamsBashforth2")),
    (4, dict(N=(16, 12, 10), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, stretch="smooth", steps=2, px=2)),
    # tilted gravity (BuoyancyForce(…; gravity_unit_vector)) on slabs and pencils
    (2, dict(N=(16, 12, 8), topo="PPB", scheme="centered", buoy="tracer", f=1e-2, bcs=True, tilt=(0.6, 0.0, -0.8), tracer_noise=1.0, steps=2)),
    (4, dict(N=(16, 12, 8), topo="PPB", scheme="weno", tilt=(0.0, -0.8660254037844386, -0.5), steps=2, px=2)),
    (2, dict(N=(16, 12, 8), topo="BBB", scheme="weno", buoy="tracer", closure="amd", tilt=(0.6, 0.0, -0.8), tracer_noise=1.0, bcs="walls", steps=2)),
    # array-valued Flux / Value / Gradient boundary conditions: every rank loads its share of each wall (a collective call)
    (2, dict(N=(16, 12, 8), topo="PPB", scheme="weno", bcs="array", steps=2)),
    (2, dict(N=(12, 10, 8), topo="BBB", scheme="centered", closure="amd", bcs="array", ts="QuasiAdamsBashforth2", steps=2)),
    (4, dict(N=(16, 12, 8), topo="PBB", scheme="upwind3", bcs="array", steps=2, px=2)),
    (4, dict(N=(16, 16, 8), topo="BPB", scheme="weno", bcs="array", steps=2)),
    # array-valued ν / κ
    (2, dict(N=(16, 12, 8), topo="PPB", scheme="centered", closure="arrays+const", bcs=True, ts="QuasiAdamsBashforth2", steps=2)),
    (4, dict(N=(16, 12, 8), topo="PBB", scheme="weno", closure="arrays", steps=2, px=2)),
    # the rest of the advection family next to one-sided walls (an outer slab of a Bounded dimension lowers the order on ONE side only)
    (2, dict(N=(16, 12, 10), topo="PPB", scheme="weno7", closure="amd", f=1e-2, bcs=True, steps=1)),
    (4, dict(N=(16, 20, 10), topo="BBB", scheme="weno9", steps=1, px=2)),
    (3, dict(N=(16, 24, 9), topo="PBB", scheme="weno7", closure="amd", bcs="walls", steps=1)),
    (2, dict(N=(16, 12, 8), topo="PBB", scheme="upwind5", steps=2)),
    (2, dict(N=(16, 12, 8), topo="BBP", scheme="centered4", steps=2, px=2)),
    (2, dict(N=(16, 12, 8), topo="PBB", scheme="weno3", steps=2)),
]

POISSON_CASES = [(R, dict(N=N, topo=topo, poisson=True)) for R, N in ((2, (16, 12, 8)), (4, (10, 16, 12)), (3, (9, 15, 6)))
                 for topo in ("PPP", "PPB", "PBB", "BBB", "BPP", "PBP")]
STRETCHED_POISSON_CASES = [
    (2, dict(N=(16, 12, 8), topo="PPB", poisson=True, stretch="smooth")),
    (4, dict(N=(10, 16, 12), topo="BBB", poisson=True, stretch="facr")),
    (3, dict(N=(9, 15, 6), topo="PBB", poisson=True, stretch="smooth")),
    (4, dict(N=(16, 12, 8), topo="PPB", poisson=True, stretch="smooth", px=2)),
    (6, dict(N=(12, 12, 6), topo="BBB", poisson=True, stretch="facr", px=3)),
    (2, dict(N=(16, 12, 8), topo="PBB", poisson=True, stretch="smooth", px=2)),
]
# the reference's stretched matrix (test/test_distributed_poisson_solvers.jl:150-163: (Bounded, Bounded, Bounded), z given by its faces),
# minus the sizes whose local extent is below this library's internal halo of 3 ((4, 44, 8) on (4,1,1), (44, 4, 8) on (1,4,1))
STRETCHED_POISSON_CASES += [(4, dict(N=N, topo="BBB", poisson=True, stretch="smooth", px=px))
                            for px, N in ((4, (44, 44, 8)), (4, (16, 44, 8)), (1, (44, 44, 8)), (1, (16, 44, 8)), (2, (22, 8, 8)), (2, (8, 22, 8)))]
PENCIL_POISSON_CASES = [(R, dict(N=N, topo=topo, poisson=True, px=px))
                        for R, px, N in ((4, 2, (16, 12, 8)), (2, 2, (16, 12, 8)), (6, 3, (12, 12, 6)), (6, 2, (12, 18, 9)))
                        for topo in ("PPP", "PBB", "BBB", "BPP")]
# local extents beyond one 32 x 32 tile of the transposition kernels (TransposeKernel, PencilYXKernel), partial tiles included
PENCIL_POISSON_CASES += [(4, dict(N=(80, 72, 4), topo=topo, poisson=True, px=2)) for topo in ("PPP", "BBB")]

# the reference's own matrix (test/test_distributed_poisson_solvers.jl:128-148): sizes x process grids (4,1,1), (1,4,1), (2,2,1) x four
# topologies, 3-D
REFERENCE_POISSON_MATRIX = [(4, dict(N=N, topo=topo, poisson=True, px=px))
                            for topo in ("PPP", "PPB", "PBB", "BBB")
                            for px, N in ((4, (44, 44, 8)), (4, (16, 44, 8)), (1, (44, 44, 8)), (1, (44, 16, 8)), (1, (16, 44, 8)),
                                          (2, (22, 44, 8)), (2, (44, 22, 8)))]

# test/test_distributed_models.jl:334-407 (H in 1:3; (4,1,1), (1,4,1), (2,2,1) process grids; here 16 x 16 x 8 so that every local size
# holds the library's internal halo of 3)
HALO_FILL_CASES = [(4, dict(N=(16, 16, 8), topo="PPP", halo_fill=H, px=px)) for H in (1, 2, 3) for px in (4, 1, 2)]

_BATCH = {}


def cpu_result(R, case):
    """Every CPU case with the same number of ranks shares ONE launch of R gloo processes (starting the processes and importing torch
    costs more than the cases themselves): the first request for a rank count runs them all."""
    if R not in _BATCH:
        todo = [c for r, c in CPU_MODEL_CASES + POISSON_CASES + PENCIL_POISSON_CASES + STRETCHED_POISSON_CASES + REFERENCE_POISSON_MATRIX + HALO_FILL_CASES if r == R]
        res = run_ranks(R, todo, timeout=1800)
        _BATCH[R] = {json.dumps(c, sort_keys=True): x for c, x in zip(todo, res)}
    return _BATCH[R][json.dumps(case, sort_keys=True)]


@pytest.mark.parametrize("R,case", CPU_MODEL_CASES)
def test_slab_decomposition_matches_single_domain_oracle(R, case):
    res = cpu_result(R, case)
    tol = 1e-4 if case.get("f32") else 1e-11
    assert res["ranks"] == R and res["worst"] <= tol, res


@pytest.mark.parametrize("R,case", POISSON_CASES, ids=[f"{r}-{c['topo']}" for r, c in POISSON_CASES])
def test_distributed_poisson_solver_matches_single_domain_solve(R, case):
    """The distributed solver alone on every mix of Periodic and Bounded dimensions, even and odd sizes, 2 / 3 / 4 ranks
    (test/test_distributed_poisson_solvers.jl:70-89,128-148 runs (4,1,1), (1,4,1), (2,2,1) partitions x 4 topologies)"""
    res = cpu_result(R, case)
    assert res["ranks"] == R and res["worst"] <= 1e-13, res


@pytest.mark.parametrize("R,case", STRETCHED_POISSON_CASES, ids=[f"{r}-px{c.get('px', 1)}-{c['topo']}-{c['stretch']}-{'x'.join(map(str, c['N']))}" for r, c in STRETCHED_POISSON_CASES])
def test_distributed_fourier_tridiagonal_solver_matches_single_domain_solve(R, case):
    """DistributedFourierTridiagonalPoissonSolver (vertically stretched grids) on slabs and pencils against the oracle's
    FourierTridiagonalPoissonSolver (test/test_distributed_poisson_solvers.jl:150-170 runs (4,1,1), (1,4,1), (2,2,1) x stretched z)"""
    res = cpu_result(R, case)
    assert res["ranks"] == R and res["worst"] <= 1e-13, res


@pytest.mark.parametrize("R,case", PENCIL_POISSON_CASES, ids=[f"{r}-px{c['px']}-{c['topo']}" for r, c in PENCIL_POISSON_CASES])
def test_pencil_poisson_solver_matches_single_domain_solve(R, case):
    """Partition(px, R / px): (2,2), (2,1), (3,2) and (2,3) process grids (test/test_distributed_poisson_solvers.jl:70-89,128-148)"""
    res = cpu_result(R, case)
    assert res["ranks"] == R and res["worst"] <= 1e-13, res


@pytest.mark.parametrize("R,case", REFERENCE_POISSON_MATRIX, ids=[f"{c['topo']}-px{c['px']}-{'x'.join(map(str, c['N']))}" for r, c in REFERENCE_POISSON_MATRIX])
def test_distributed_poisson_solver_on_the_reference_test_matrix(R, case):
    """sizes, process grids and topologies of the reference's 3-D distributed solver tests (test/test_distributed_poisson_solvers.jl:128-148)"""
    res = cpu_result(R, case)
    assert res["ranks"] == R and res["worst"] <= 1e-13, res


@pytest.mark.parametrize("R,case", HALO_FILL_CASES, ids=[f"H{c['halo_fill']}-px{c['px']}" for r, c in HALO_FILL_CASES])
def test_halo_communication_fills_halos_with_the_neighbours_rank(R, case):
    """every field = local rank; the halos (corners included) must then hold the neighbouring ranks (test_distributed_models.jl:334-407)"""
    res = cpu_result(R, case)
    assert res["ranks"] == R and res["worst"] == 0, res


def test_distributed_rejects_unsupported_configurations():
    import oceananigans_b200 as ob
    with pytest.raises(NotImplementedError):
        ob.Partition(2, 2, 2)              # z is never partitioned
    with pytest.raises(ValueError):
        ob.Distributed(ob.B200(0), partition=ob.Partition(2, 2), rank=0, nranks=6)
    # a Flat dimension cannot be partitioned: an (x, y) model on slabs in y is refused by the library, loudly
    import __graft_entry__ as ge
    from oceananigans_b200 import _lib
    ge.build()
    lib = _lib.Library(ge.HOSTSIM)
    arch = ob.Distributed(ob.B200(0), partition=ob.Partition(1, 2), rank=0, nranks=2, exchange=lambda msgs: 0)
    grid = ob.RectilinearGrid(arch, np.float64, size=(8, 8), extent=(1, 1), topology=(ob.Periodic, ob.Periodic, ob.Flat))
    with pytest.raises(_lib.OceananigansB200Error, match="Flat"):
        ob.NonhydrostaticModel(grid=grid, advection=ob.Centered(), library=lib)
    # the distributed solver's divisibility constraints (distributed_fft_based_poisson_solver.jl:211-229)
    grid = ob.RectilinearGrid(arch, np.float64, size=(8, 8, 5), extent=(1, 1, 1), topology=(ob.Periodic, ob.Periodic, ob.Bounded))
    with pytest.raises(_lib.OceananigansB200Error, match="divisible"):
        ob.NonhydrostaticModel(grid=grid, advection=ob.Centered(), library=lib)
    with pytest.raises(ValueError):
        ob.Distributed(ob.B200(0), partition=ob.Partition(1, 3), rank=0, nranks=2)


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


NCCL_CASES = [
    (2, dict(N=(64, 48, 32), topo="PPP", scheme="weno", steps=2)),
    (2, dict(N=(40, 24, 16), topo="PPP", scheme="centered", f=1e-2, steps=3)),
    (2, dict(N=(48, 32, 16), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, steps=2)),
    # slabs tall enough (64 rows) for the interior / boundary-strip split: the end-of-stage exchange of stages 1 and 2 overlaps the next
    # stage's interior tile rows (interleave_communication_and_computation.jl:29-67)
    (2, dict(N=(40, 128, 16), topo="PPP", scheme="weno", steps=2)),
    (2, dict(N=(40, 96, 16), topo="PPB", scheme="centered", f=1e-2, bcs=True, steps=2)),
    # R >= 3: the two neighbours are different peers (R = 2 is the degenerate case), every rank sends R - 1 transposed-FFT chunks
    (4, dict(N=(64, 48, 32), topo="PPP", scheme="weno", steps=2)),
    (4, dict(N=(48, 32, 16), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, steps=2)),
    (4, dict(N=(40, 24, 16), topo="PPP", scheme="centered", f=1e-2, steps=3, ts="QuasiAdamsBashforth2")),
    (8, dict(N=(64, 48, 32), topo="PPP", scheme="weno", steps=2)),
    (8, dict(N=(48, 32, 16), topo="PPB", scheme="weno", closure="lilly", f=("beta", 0.3, 2.0), bcs=True, steps=2)),
    (8, dict(N=(40, 24, 16), topo="PPP", scheme="weno", steps=2, f32=True)),
]

# Added after the round's last multi-GPU run: verified on the host simulation (gloo and thread ranks) but NOT yet on GPUs — the single 2-GPU
# call made for them ran into the end of the GPU budget without a result (profiles/README.md).  They run when OC_NCCL_NEW_CASES=1
# (scripts/gpu.sh disttests sets it), so that an unvalidated multi-GPU case cannot stop a `pytest -m gpu -x` run of the validated suite.
NCCL_NEW_CASES = [
    # Bounded x and y on slabs (walls on the outer ranks; DCTs in x on the slab and in y in the transposed layout)
    (2, dict(N=(48, 64, 16), topo="PBB", scheme="weno", closure="amd", f=1e-2, bcs="walls", steps=2)),
    (2, dict(N=(40, 24, 16), topo="BPB", scheme="weno", bcs="walls", steps=2)),
    (4, dict(N=(48, 64, 16), topo="BBB", scheme="weno", closure="lilly", f=("beta", 0.3, 2.0), bcs="walls", steps=2)),
    (4, dict(N=(40, 32, 16), topo="PBP", scheme="centered", bcs="walls", steps=2, ts="QuasiAdamsBashforth2")),
    (8, dict(N=(48, 64, 16), topo="BBB", scheme="weno", bcs="walls", steps=2)),
    (2, dict(N=(36, 40, 20), topo="BBB", poisson=True)),
    (4, dict(N=(36, 40, 20), topo="PBB", poisson=True)),
    (8, dict(N=(30, 48, 24), topo="BBB", poisson=True)),
    # pencils
    (4, dict(N=(64, 48, 32), topo="PPP", scheme="weno", steps=2, px=2)),
    (4, dict(N=(48, 64, 16), topo="BBB", scheme="weno", closure="amd", f=1e-2, bcs="walls", steps=2, px=2)),
    (2, dict(N=(48, 32, 16), topo="PPB", scheme="weno", bcs=True, steps=2, px=2)),
    (8, dict(N=(64, 48, 32), topo="PPB", scheme="weno", closure="lilly", f=("beta", 0.3, 2.0), bcs=True, steps=2, px=2)),
    (8, dict(N=(64, 48, 32), topo="PBB", scheme="centered", bcs="walls", steps=2, px=4, ts="QuasiAdamsBashforth2")),
    (4, dict(N=(36, 40, 20), topo="BBB", poisson=True, px=2)),
    (8, dict(N=(32, 48, 24), topo="PBP", poisson=True, px=4)),
    # vertically stretched grids
    (2, dict(N=(48, 32, 16), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, stretch="smooth", steps=2)),
    (4, dict(N=(48, 32, 16), topo="PPB", scheme="weno", closure="lilly", bcs=True, stretch="smooth", steps=2, px=2)),
    (4, dict(N=(36, 40, 20), topo="BBB", poisson=True, stretch="facr")),
]

def _nccl_case(R, case):
    if _gpu_count() < R:
        pytest.skip(f"needs {R} GPUs")
    res = run_ranks(R, dict(case), timeout=300, backend="nccl")      # one launch per case: one model and one communicator per process
    assert res["ranks"] == R and res["worst"] <= (1e-4 if case.get("f32") else 1e-11), res


@pytest.mark.gpu
@pytest.mark.parametrize("R,case", NCCL_CASES)
def test_nccl_slab_decomposition_matches_oracle(R, case):
    """The CUDA library on R GPUs of one box (NCCL halo exchange + transposed distributed FFT) against the oracle."""
    _nccl_case(R, case)


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("OC_NCCL_NEW_CASES") != "1", reason="not yet validated on GPUs: set OC_NCCL_NEW_CASES=1 (scripts/gpu.sh disttests)")
@pytest.mark.parametrize("R,case", NCCL_NEW_CASES)
def test_nccl_bounded_and_pencil_decompositions_match_oracle(R, case):
    """Distributed Bounded x / y and pencil partitions on R GPUs against the oracle."""
    _nccl_case(R, case)
