mkdir -p gpurun_out
for ty in 8 16; do
OC_MARCH_TY=$ty timeout 600 python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('TY=$ty', 'ms/step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
    else: print(l.rstrip()[:300])
"
done
OC_MARCH_TY=16 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
OC_MARCH_TY=16 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "C3 or C2 or full_size" 2>&1 | tail -3
