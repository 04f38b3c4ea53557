import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import parity_harness as ph
scheme = sys.argv[1] if len(sys.argv) > 1 else "weno"
buoy = sys.argv[2] if len(sys.argv) > 2 else "seawater"
out, m, om = ph.run_case(steps=(1,), N=(16, 12, 8), topo="PPP", scheme=scheme, FT=np.float32, buoy=buoy, closure="none")
print(scheme, buoy, {s: max(v.values()) for s, v in out.items()})
