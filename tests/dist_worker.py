"""One rank of the CPU multi-process test of the slab-decomposed path (launched by tests/test_distributed.py).

Each rank builds the host simulation of the kernels with Distributed(partition=Partition(1, R)) and a gloo transport
(torch.distributed isend/irecv on the library's buffers), steps the model, and compares its slab with the single-domain oracle.
Usage: python tests/dist_worker.py <case-json>   with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gloo_exchange(msgs):
    """oc_exchange_fn: (send_peer, recv_peer, tag, send_ptr, send_bytes, recv_ptr, recv_bytes) per message."""
    reqs, keep = [], []
    for sp, rp, tag, sptr, sb, rptr, rb in msgs:
        if rb == 0:                 # a wall side of a Bounded y: that half of the message is empty
            continue
        r = torch.frombuffer((C.c_char * rb).from_address(rptr), dtype=torch.uint8)
        keep.append(r)
        reqs.append(dist.irecv(r, src=rp, tag=tag))
    for sp, rp, tag, sptr, sb, rptr, rb in msgs:
        if sb == 0:
            continue
        s = torch.frombuffer((C.c_char * sb).from_address(sptr), dtype=torch.uint8)
        keep.append(s)
        reqs.append(dist.isend(s, dst=sp, tag=tag))
    for q in reqs:
        q.wait()
    return 0


def halo_fill(kw, rank, arch, lib, FT):
    """test/test_distributed_models.jl:334-407: every field filled with the local rank; after fill_halo_regions! the halos of the
    partitioned directions hold the neighbour's rank (the process grid wraps: triply periodic), the others the local rank, and — on a
    (2, 2) process grid — the corners hold the diagonal neighbour's.  Returns the number of mismatching halo cells."""
    import oceananigans_b200 as ob
    H = int(kw["halo_fill"])
    part = arch.partition
    grid = ob.RectilinearGrid(arch, FT, size=tuple(kw["N"]), extent=(1, 2, 3), halo=(H, H, H), topology=(ob.Periodic,) * 3)
    model = ob.NonhydrostaticModel(grid=grid, advection=ob.Centered() if H < 3 else ob.WENO(), tracers=("c",), closure=None, library=lib)
    rank_of = lambda i, j: (i % part.x) * part.y + (j % part.y)            # index2rank, distributed_architectures.jl:354
    i, j = arch.rx, arch.ry
    bad = 0
    for f in list(model.fields.values()) + [model.pressures.pNHS]:
        f.set(float(rank))
        ob.fill_halo_regions_(f)
        p = f.parent()
        inner = (slice(H, -H),) * 3
        want = np.full(p.shape, float(rank))
        lo, hi = slice(0, H), slice(-H, None)
        mid = slice(H, -H)
        for sx, di in ((lo, -1), (mid, 0), (hi, +1)):
            for sy, dj in ((lo, -1), (mid, 0), (hi, +1)):
                want[sx, sy, :] = rank_of(i + di, j + dj)
        assert np.all(p[inner] == rank)
        bad += int(np.count_nonzero(p != want))
    return float(bad)


def run_rank(kw, rank, R, arch, lib, nccl=False):
    """Build this rank's model, run the case, return the worst relative error of its slab against the single-domain oracle."""
    import oceananigans_b200 as ob
    import parity_harness as ph
    kw = dict(kw)
    steps = kw.pop("steps", 2)
    FT = np.float32 if kw.pop("f32", False) else np.float64
    N, topo = tuple(kw["N"]), kw["topo"]
    scheme = kw.get("scheme", "weno")
    if isinstance(scheme, list):                 # JSON: a tuple of per-direction schemes arrives as a list
        scheme = tuple(scheme)
    f = kw.get("f")
    if isinstance(f, list):                      # ("beta", f₀, β): JSON turns the tuple into a list
        f = tuple(f)
    kw["f"] = f
    case = dict(N=N, topo=topo, scheme=scheme, FT=FT, f=kw.get("f"), closure=kw.get("closure", "scalar"), bcs=kw.get("bcs", False),
                ts=kw.get("ts", "RungeKutta3"), buoy=kw.get("buoy", "seawater"), tilt=tuple(kw["tilt"]) if kw.get("tilt") else None,
                stretch=kw.get("stretch"), halo=tuple(kw["halo"]) if kw.get("halo") else None,
                extent=tuple(kw["extent"]) if kw.get("extent") else ph.EXTENT)
    if kw.get("halo_fill"):
        return halo_fill(kw, rank, arch, lib, FT)
    model = ph.build_product(library=lib, arch=arch, **case)
    om = ph.build_oracle(**case)
    ic = ph.initial_conditions(om, tracer_noise=kw.get("tracer_noise", 0.01))
    part = arch.partition
    nxl, nyl = N[0] // part.x, N[1] // part.y
    cols = slice(arch.rx * nxl, (arch.rx + 1) * nxl)
    rows = slice(arch.ry * nyl, (arch.ry + 1) * nyl)
    if kw.get("poisson"):
        # the distributed solver on its own (test/test_distributed_poisson_solvers.jl:70-89,128-148): a seeded zero-mean right-hand side,
        # every rank solving for its rows, against the single-domain solve
        rng = np.random.default_rng(99)
        worst = 0.0
        for _ in range(2):
            rhs = rng.standard_normal(N).astype(FT)
            rhs -= rhs.mean()
            if kw.get("stretch"):          # solve!(ϕ, FourierTridiagonalPoissonSolver, rhs): the source term is Δzᶜ · rhs (:239-246)
                dzc = om.grid.dz_at("c", np.arange(1, N[2] + 1))
                rhs -= (rhs * dzc).sum() / (dzc.sum() * N[0] * N[1])          # compatible source: Σ Δz·b = 0
                want = om.solve_poisson_tridiagonal(rhs * dzc)
            else:
                want = om.solve_poisson(rhs)
            got = ob.solve_poisson(model, rhs[cols, rows, :])
            worst = max(worst, float(np.abs(got - want[cols, rows, :]).max() / np.abs(want).max()))
        return worst

    def sl_of(n):
        # this rank's share of field n; the last rank of a Bounded partitioned dimension also owns the wall face of the velocity normal to
        # it (LeftConnected: N_l + 1 faces, grid_utils.jl:43)
        cs, rs = cols, rows
        if n == "u" and topo[0] == "B" and arch.rx == part.x - 1:
            cs = slice(cols.start, cols.stop + 1)
        if n == "v" and topo[1] == "B" and arch.ry == part.y - 1:
            rs = slice(rows.start, rows.stop + 1)
        return cs, rs, slice(None)

    ob.set_(model, **{n: a[sl_of(n)] for n, a in ic.items()})
    om.set(**ic)
    dmin = [float(om.grid.D[d]) for d in range(3) if not om.grid.flat(d) and om.grid.D[d] is not None]
    if om.grid.stretched:
        dmin.append(float(np.min(om.grid.dz_at("c", np.arange(1, om.grid.Nz + 1)))))
    dt = 0.1 * min(dmin)
    worst = 0.0
    for s in range(steps + 1):
        if s:
            ob.time_step_(model, dt)
            om.time_step(dt)
        for n in om.fields:
            sl = sl_of(n)
            worst = max(worst, ph.rel_linf(model.fields[n].interior(), om.fields[n].interior[sl]) * (np.abs(om.fields[n].interior[sl]).max() / np.abs(om.fields[n].interior).max()))
        if scheme != "none":      # without advection the pressure correction is round-off sized: its RELATIVE error is meaningless
            worst = max(worst, float(np.abs(model.pressures.pNHS.interior() - om.pNHS.interior[cols, rows, :]).max() / np.abs(om.pNHS.interior).max()))
    if max(float(np.abs(om.fields[n].interior).max()) for n in ("u", "v", "w")) > 1e3:
        return float("nan")       # the harness's time step is unstable for this configuration (the fuzzer skips it)
    return worst


def main():
    cases = json.loads(sys.argv[1])              # one case, or a list of cases run one after the other on the same process group
    many = isinstance(cases, list)
    if not many:
        cases = [cases]
    nccl = os.environ.get("OC_DIST_BACKEND", "gloo") == "nccl"      # -m gpu variant: the CUDA library, NCCL over NVLink
    if nccl:
        torch.cuda.set_device(int(os.environ["RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["RANK"])))
    else:
        dist.init_process_group("gloo")
    rank, R = dist.get_rank(), dist.get_world_size()
    import oceananigans_b200 as ob
    from oceananigans_b200 import _lib
    import __graft_entry__ as ge
    lib = None if nccl else _lib.Library(ge.HOSTSIM)
    out = []
    for kw in cases:
        px = int(kw.get("px", 1))                   # Partition(px, R / px): ranks along x
        arch = ob.Distributed(ob.B200(rank if nccl else 0), partition=ob.Partition(px, R // px), rank=rank, nranks=R,
                              exchange=None if nccl else gloo_exchange)
        worst = run_rank(kw, rank, R, arch, lib, nccl)
        t = torch.tensor([worst], dtype=torch.float64, device="cuda" if nccl else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append({"worst": float(t.item()), "ranks": R, "steps": kw.get("steps", 2)})
    if rank == 0:
        print(json.dumps(out if many else out[0]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
