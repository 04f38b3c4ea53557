"""CPU test of bench.py's contract: the reference arm prints exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                        "--warmup", "0", "--cpu-size", "16"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key


def test_workload_byte_model_matches_baseline_table():
    """BASELINE.md §2 / SURVEY §8d: reals per cell per step 87 (C2), 119 (C3), 152 (C4), 66 (C1)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.reals_per_cell_step(bench.WORKLOADS["c2"]) == 87
    assert bench.reals_per_cell_step(bench.WORKLOADS["c3"]) == 119
    assert bench.reals_per_cell_step(bench.WORKLOADS["c4"]) == 152
    assert bench.reals_per_cell_step(bench.WORKLOADS["c1"]) == 66


def test_parity_check_of_the_bench_line_on_the_host_simulation():
    """bench.py's `parity_check` (the correctness figure next to every throughput number) run on CPU: the host simulation of the kernel
    sources stands in for the CUDA library, one rank."""
    sys.path.insert(0, ROOT)
    import bench
    from test_host_api import _hostsim
    for name in ("c3", "c2"):
        r = bench.parity_check(bench.WORKLOADS[name], 0, 0, 1, steps=1, library=_hostsim())
        assert r["ok"] and r["worst"] <= 1e-11 and r["grid"] == [64, 16, 64], r
    assert bench.parity_check(bench.WORKLOADS["c4"], 0, 0, 1) is None        # outside the C twin's scope: no figure rather than a wrong one
