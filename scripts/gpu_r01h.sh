mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_h.log
for w in c3 c4 c4s c2 c3f32; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${w}_h.json 2> gpurun_out/bench_${w}_h.err; echo "$w rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_${w}_h.json"))
print("$w", d["ms_per_step"], {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"])
P
done
