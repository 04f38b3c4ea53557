mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "matches_oracle or golden" > gpurun_out/pytest_cpt.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_cpt.log
for w in c3 c2 c4 c3f32; do
for cpt in 2 1; do
OC_MARCH_CPT=$cpt timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${w}_cpt$cpt.json 2> gpurun_out/bench_${w}_cpt$cpt.err; echo "$w cpt=$cpt rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_${w}_cpt$cpt.json"))
print("$w cpt=$cpt", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items() if v}, d["clocks"].get("reasons"))
P
done
done
