"""The peer-memory path of the distributed pressure solve on CPU.

On a GPU box with up to four ranks the transposes of the distributed FFT are ONE kernel each over CUDA-IPC peer memory (TransposePutKernel,
csrc/oc_dist.h) instead of pack -> all-to-all -> unpack; the gloo multi-process tests (tests/test_distributed.py) cannot reach that kernel
because separate processes share no address space.  Here the ranks are THREADS of one process running the host simulation: a peer's buffer
is simply its pointer (HostTransport with OC_HOSTSIM_THREADS=1), the collectives (pointer exchange, barriers, halo messages) go through an
in-process mailbox, and every slab is compared with the single-domain oracle exactly like the multi-process tests do."""
import ctypes as C
import os
import queue
import sys
import threading

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Mailbox:
    """One FIFO per (source, destination): like ncclSend / ncclRecv, messages between a pair of ranks carry NO tag and are matched
    purely in posting order, and the two sides must agree on the size — so an exchange whose correctness leaned on the tags of the
    host-callback interface (which gloo honours) fails here, as it would on NCCL."""

    def __init__(self):
        self.lock = threading.Lock()
        self.queues = {}

    def of(self, key):
        with self.lock:
            return self.queues.setdefault(key, queue.Queue())

    def exchange_for(self, rank):
        def exchange(msgs):
            for sp, rp, tag, sptr, sb, rptr, rb in msgs:
                if sb:
                    self.of((rank, sp)).put(C.string_at(sptr, sb))
            for sp, rp, tag, sptr, sb, rptr, rb in msgs:
                if rb:
                    data = self.of((rp, rank)).get(timeout=300)
                    if len(data) != rb:
                        raise RuntimeError(f"rank {rank}: a {rb}-byte receive from rank {rp} met a {len(data)}-byte send (tag {tag})")
                    C.memmove(rptr, data, rb)
            return 0
        return exchange


def run_threads(R, case, p2p=True):
    import __graft_entry__ as ge
    import oceananigans_b200 as ob
    from oceananigans_b200 import _lib
    import dist_worker
    ge.build()
    os.environ["OC_HOSTSIM_THREADS"] = "1"
    os.environ["OC_DIST_P2P"] = "1" if p2p else "0"
    try:
        lib = _lib.Library(ge.HOSTSIM)
        box = Mailbox()
        out, errs = [None] * R, []

        def body(rank):
            try:
                px = int(case.get("px", 1))
                arch = ob.Distributed(ob.B200(0), partition=ob.Partition(px, R // px), rank=rank, nranks=R, exchange=box.exchange_for(rank))
                out[rank] = dist_worker.run_rank(case, rank, R, arch, lib)
            except BaseException as e:       # noqa: BLE001 — reported by the main thread
                errs.append((rank, repr(e)))

        threads = [threading.Thread(target=body, args=(r,)) for r in range(R)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=900)
        assert not errs, errs
        assert all(o is not None for o in out), "a rank did not finish"
        return max(out)
    finally:
        os.environ.pop("OC_HOSTSIM_THREADS", None)
        os.environ.pop("OC_DIST_P2P", None)


@pytest.mark.parametrize("R,case", [
    (2, dict(N=(16, 12, 8), topo="PPP", scheme="weno", steps=2)),
    (4, dict(N=(12, 16, 8), topo="PPB", scheme="weno", closure="amd", f=1e-2, bcs=True, steps=1)),
    # Bounded y: the Makhoul permutation of the y line happens inside the transposed put
    (2, dict(N=(16, 12, 8), topo="PBB", scheme="weno", bcs="walls", steps=2)),
    (3, dict(N=(12, 18, 9), topo="BBB", scheme="centered", bcs="walls", steps=1)),
    (4, dict(N=(10, 16, 12), topo="BBB", poisson=True)),
    (4, dict(N=(10, 16, 12), topo="PBP", poisson=True)),
    (3, dict(N=(9, 15, 6), topo="BPB", poisson=True)),
    # pencils with thread ranks (send / receive transposes: the peer-memory kernel is the slab path's)
    (4, dict(N=(16, 12, 8), topo="PBB", scheme="weno", bcs="walls", steps=1, px=2)),
    # eight ranks, as in the driver's scaling run (C5 is 8 slabs): slabs through peer memory, 2 x 4 pencils, a stretched grid
    (8, dict(N=(16, 32, 8), topo="PPP", scheme="weno", steps=1)),
    (8, dict(N=(16, 32, 8), topo="PPB", scheme="weno", steps=1, px=2)),
    (8, dict(N=(16, 32, 16), topo="PPB", scheme="weno", closure="lilly", bcs=True, stretch="smooth", steps=1)),
    # a (y, z) model on slabs: the first stage of the peer-memory solve transforms z alone
    (2, dict(N=(1, 12, 8), topo="FPB", scheme="weno", buoy="tracer", f=0.2, steps=2)),
])
def test_peer_memory_transposes_match_single_domain_oracle(R, case):
    worst = run_threads(R, case)
    assert worst <= 1e-11, worst


def test_thread_ranks_with_all_to_all_agree_with_peer_memory(capfd, monkeypatch):
    """the same case through the send / receive transposes (what more than four ranks use) and through peer memory: both against the
    oracle, and the library says which path it took"""
    monkeypatch.setenv("OC_VERBOSE", "1")
    case = dict(N=(12, 16, 8), topo="PBB", scheme="weno", steps=1)
    assert run_threads(4, case, p2p=False) <= 1e-11
    err = capfd.readouterr().err
    assert err.count("transposes over send/receive all-to-all") == 4 and "peer memory" not in err
    assert run_threads(4, case, p2p=True) <= 1e-11
    err = capfd.readouterr().err
    assert err.count("transposes over peer memory") == 4 and "all-to-all" not in err


def test_tall_slabs_split_their_tendency_launches(monkeypatch):
    """interleave_communication_and_computation.jl:29-67: after a deferred end-of-stage exchange the next stage launches every field's
    interior tile rows first and the two boundary strips afterwards — more launches, the same numbers (parity: tests/test_distributed.py)"""
    import __graft_entry__ as ge
    import oceananigans_b200 as ob
    from oceananigans_b200 import _lib
    import parity_harness as ph
    ge.build()
    monkeypatch.setenv("OC_HOSTSIM_THREADS", "1")
    lib = _lib.Library(ge.HOSTSIM)

    def launches(split):
        monkeypatch.setenv("OC_XCHG_OVERLAP", "1" if split else "0")
        box, R, out = Mailbox(), 2, [None, None]

        def body(rank):
            arch = ob.Distributed(ob.B200(0), partition=ob.Partition(1, R), rank=rank, nranks=R, exchange=box.exchange_for(rank))
            m = ph.build_product(library=lib, arch=arch, N=(16, 64, 8), topo="PBB", scheme="weno", bcs="walls")
            n0 = m.launch_count()
            ob.time_step_(m, 1e-3)
            out[rank] = (m.launch_count() - n0, m.fields["T"].interior().copy())
        ts = [threading.Thread(target=body, args=(r,)) for r in range(R)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=600)
        return out

    a, b = launches(True), launches(False)
    assert a[0][0] > b[0][0] and a[1][0] > b[1][0], (a[0][0], b[0][0])
    for r in range(2):
        assert (a[r][1] == b[r][1]).all()           # bit for bit: the same kernels on the same cells, in two or three launches


@pytest.mark.parametrize("R,px,kw", [
    (2, 1, dict(N=(12, 12, 8), topo="PPB", scheme="weno", closure="amd", bcs=True, f=1e-2)),
    (4, 2, dict(N=(12, 12, 8), topo="BBB", scheme="weno", closure="lilly", bcs="walls")),
    (3, 1, dict(N=(12, 12, 9), topo="PBB", scheme="centered", bcs="array", stretch="smooth")),
], ids=["slabs", "pencils", "stretched slabs with array BCs"])
def test_staged_entry_points_match_fused_step_on_distributed_models(R, px, kw, monkeypatch):
    """time_step! assembled from the staged C entry points (runge_kutta_3.jl:93-170) — every one of them collective on a distributed model —
    against the fused oc_time_step_rk3: two models per rank (each with its own mailbox), two steps, every field"""
    import numpy as np
    import __graft_entry__ as ge
    import oceananigans_b200 as ob
    from oceananigans_b200 import _lib
    import parity_harness as ph
    ge.build()
    monkeypatch.setenv("OC_HOSTSIM_THREADS", "1")
    lib = _lib.Library(ge.HOSTSIM)
    boxes = (Mailbox(), Mailbox())
    out, errs = [None] * R, []
    om = ph.build_oracle(**kw)
    ic = ph.initial_conditions(om)
    N, topo, py = kw["N"], kw["topo"], R // px
    nxl, nyl = N[0] // px, N[1] // py

    def body(rank):
        try:
            models = []
            for box in boxes:
                arch = ob.Distributed(ob.B200(0), partition=ob.Partition(px, py), rank=rank, nranks=R, exchange=box.exchange_for(rank))
                m = ph.build_product(library=lib, arch=arch, **kw)
                cs = lambda n: slice(arch.rx * nxl, (arch.rx + 1) * nxl + (1 if n == "u" and topo[0] == "B" and arch.rx == px - 1 else 0))
                rs = lambda n: slice(arch.ry * nyl, (arch.ry + 1) * nyl + (1 if n == "v" and topo[1] == "B" and arch.ry == py - 1 else 0))
                ob.set_(m, **{n: a[cs(n), rs(n), :] for n, a in ic.items()})
                models.append(m)
            m1, m2 = models
            dt, g, z = 0.02, [8 / 15, 5 / 12, 3 / 4], [0.0, -17 / 60, -5 / 12]
            worst = 0.0
            for _ in range(2):
                ob.time_step_(m1, dt)
                ob.update_state_(m2, True)
                for st in (1, 2, 3):
                    ob.compute_flux_bc_tendencies_(m2)
                    ob.rk3_substep_(m2, dt, st)
                    sdt = dt * (g[st - 1] + z[st - 1])
                    ob.compute_pressure_correction_(m2, sdt)
                    ob.make_pressure_correction_(m2, sdt)
                    if st < 3:
                        ob.cache_previous_tendencies_(m2)
                    ob.update_state_(m2, True)
                for n in m1.fields:
                    worst = max(worst, ph.rel_linf(m1.fields[n].interior(), m2.fields[n].interior()))
            out[rank] = worst
        except BaseException as e:       # noqa: BLE001
            errs.append((rank, repr(e)))

    ts = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(R)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not errs, errs
    assert all(o is not None and np.isfinite(o) for o in out) and max(out) <= 1e-13, out
