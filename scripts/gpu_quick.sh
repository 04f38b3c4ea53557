mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log
for w in ${WL:-c4 c4s}; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${w}_q.json 2> gpurun_out/bench_${w}_q.err; echo "$w rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_${w}_q.json"))
print("$w", d["ms_per_step"], {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"])
P
done
