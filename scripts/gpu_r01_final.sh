mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01f_bench_reference_arm.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r01f_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r01f_bench_c3_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"; cat gpurun_out/r01f_bench_c3_default.json; tail -3 gpurun_out/bench_default.err
for w in c4 c4s c2 c3f32 c1; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e > gpurun_out/r01f_bench_${w}.json 2> gpurun_out/bench_${w}.err; echo "$w rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/r01f_bench_${w}.json"))
print("$w", d["ms_per_step"], {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["clocks"])
P
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_c3.log 2>&1; echo "list c3 rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches_c4s.csv python bench.py --workload c4s --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_c4s.log 2>&1; echo "list c4s rc=$?"
