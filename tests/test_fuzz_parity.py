"""A short, seeded run of the configuration fuzzer (scripts/fuzz_parity.py): random combinations of topology, sizes, scheme, closure,
buoyancy, Coriolis form, boundary conditions, time stepper, float type, stretched z, tilted gravity and domain decomposition — the host
simulation of the kernels against the oracle.  Longer runs (hundreds of cases, other seeds) are a command away; what they found is pinned
as explicit cases in tests/parity_harness.py and tests/test_distributed.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("flags", [["--single-only", "--seed", "5", "--cases", "16"], ["--dist-only", "--seed", "6", "--cases", "10"],
                                   ["--staged", "--seed", "7", "--cases", "10"], ["--checkpoint", "--seed", "8", "--cases", "8"]],
                         ids=["single-domain", "thread-rank decompositions", "staged entry points vs fused step", "checkpoint pickup bit for bit"])
def test_seeded_fuzz_run_agrees_with_the_oracle(flags):
    import __graft_entry__ as ge
    ge.build()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_parity.py")] + flags, capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " 0 failures" in r.stdout and "compared" in r.stdout, tail
