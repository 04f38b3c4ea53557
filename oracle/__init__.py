"""CPU oracle: a NumPy/SciPy restatement of the Oceananigans NonhydrostaticModel time step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``oldoceananigans.jl_b200`` / ``oceananigans_b200``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / CPU baseline.

Parity status: the Julia reference cannot be executed in the build container (no
``julia``), and its regression golden files are downloaded at test time
(``test/data_dependencies.jl:17-38``), so they are not available.  The oracle is
pinned against the reference's *known-answer tests* instead (SURVEY.md §8c items
1-7; see ``tests/test_oracle_*.py``).  End-to-end "reference CPU after N steps"
fixtures do not exist: **parity against the Julia binary itself is unpinned**.

Every function cites the reference file:line it restates (paths relative to
``/root/reference``).
"""
from .grid import Grid, Field, PERIODIC, BOUNDED, FLAT  # noqa: F401
from .model import OracleModel  # noqa: F401
