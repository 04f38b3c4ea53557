mkdir -p gpurun_out
for ov in 0 1; do
if [ $ov = 0 ]; then export OC_NO_OVERLAP=1; else unset OC_NO_OVERLAP; fi
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('1gpu overlap=$ov', 'ms/step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('2gpu overlap=$ov', 'ms/step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
done
