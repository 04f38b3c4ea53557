"""Randomised differential test of the pressure solvers alone (CPU only: the host simulation against the oracle's solvers): random
sizes (odd, prime, tile-crossing), every mix of Periodic / Bounded dimensions, regular and vertically stretched grids, Float64 / Float32,
on one domain, slabs in y, slabs in x and pencils up to 3 x 2 (thread ranks with NCCL-like message matching,
tests/test_distributed_threads.py).   python scripts/fuzz_poisson.py [cases] [seed]"""
import sys, os, threading, warnings
warnings.simplefilter("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
os.environ["OC_HOSTSIM_THREADS"] = "1"
import numpy as np
import __graft_entry__ as ge, oceananigans_b200 as ob, parity_harness as ph, dist_worker
from oceananigans_b200 import _lib
from test_distributed_threads import Mailbox
lib = _lib.Library(ge.HOSTSIM)
CASES = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
bad = 0; n = 0
for it in range(CASES):
    topo = "".join(rng.choice(["P", "B"]) for _ in range(3))
    part = [(1, 1), (1, 2), (2, 1), (2, 2), (1, 3), (3, 1), (1, 4), (2, 3), (3, 2), (4, 1)][int(rng.integers(0, 10))]
    loc = [int(rng.integers(3, 14)) for _ in range(3)]
    N = [loc[0] * part[0], loc[1] * part[1], loc[2]]
    N[2] = ((N[2] + part[1] - 1) // part[1]) * part[1]
    m = part[0] * part[1]
    N[1] = ((N[1] + m - 1) // m) * m
    stretch = str(rng.choice(["smooth", "facr"])) if topo[2] == "B" and rng.random() < 0.4 else None
    f32 = rng.random() < 0.15
    case = dict(N=tuple(N), topo=topo, poisson=True, px=part[0], f32=bool(f32))
    if stretch: case["stretch"] = stretch
    R = m
    box = Mailbox(); out = [None] * R; errs = []
    def body(rank):
        try:
            arch = ob.Distributed(ob.B200(0), partition=ob.Partition(*part), rank=rank, nranks=R, exchange=box.exchange_for(rank)) if R > 1 else None
            if R == 1:
                FT = np.float32 if f32 else np.float64
                mm, om = ph.build_pair(library=lib, N=tuple(N), topo=topo, FT=FT, stretch=stretch)
                r2 = np.random.default_rng(99); w = 0.0
                rhs = r2.standard_normal(tuple(N)).astype(FT); rhs -= rhs.mean()
                if stretch:
                    dzc = om.grid.dz_at("c", np.arange(1, N[2] + 1)); rhs -= (rhs * dzc).sum() / (dzc.sum() * N[0] * N[1]); want = om.solve_poisson_tridiagonal(rhs * dzc)
                else:
                    want = om.solve_poisson(rhs)
                got = ob.solve_poisson(mm, rhs)
                out[rank] = float(np.abs(got - want).max() / np.abs(want).max())
            else:
                out[rank] = dist_worker.run_rank(case, rank, R, arch, lib)
        except BaseException as e:
            errs.append(repr(e)[:300])
    ts = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(R)]
    [t.start() for t in ts]; [t.join(timeout=300) for t in ts]
    if errs:
        print(it, case, "ERR", errs[0]); bad += 1; continue
    w = max(out); tol = 2e-4 if f32 else 1e-12
    n += 1
    if not w <= tol:
        print(it, case, "FAIL", w); bad += 1
print(n, "solves compared,", bad, "bad")
sys.exit(1 if bad else 0)
