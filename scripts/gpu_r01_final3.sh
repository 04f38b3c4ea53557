mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r01h_bench_c3_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"; cat gpurun_out/r01h_bench_c3_default.json; tail -3 gpurun_out/bench_default.err
for w in c3f32 c3u5; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01h_bench_${w}.json 2> gpurun_out/bench_${w}.err; echo "$w rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/r01h_bench_${w}.json"))
print("$w", d["ms_per_step"], {k:round(v,2) for k,v in d["kernel_ms_per_step"].items() if v}, d["clocks"])
P
done
