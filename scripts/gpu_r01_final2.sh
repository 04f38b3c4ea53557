mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
for w in c3u5 c2c4 c3 c4 c4s; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e > gpurun_out/r01g_bench_${w}.json 2> gpurun_out/bench_${w}.err; echo "$w rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/r01g_bench_${w}.json"))
print("$w", d["ms_per_step"], {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"], d.get("cpu_baseline",{}).get("value"))
P
done
