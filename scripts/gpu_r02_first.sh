# First GPU call of round 2: evidence for what round 1 could only verify on the host-simulation build.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_r02_first.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_widening_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r02a_pytest_widening.log; cat gpurun_out/r02a_pytest_widening.log
# AMD buoyancy modification (AmdKernel<FT, STR, CB = true>): the C4 physics with Cb = 1 next to plain C4
timeout 120 python bench.py --workload c4cb --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_c4cb.json 2> gpurun_out/r02a_bench_c4cb.err; echo "c4cb rc=$?"
# the C4 physics with (SmagorinskyLilly, ScalarDiffusivity): the division-free SmagorinskyKernel has no timing yet
timeout 120 python bench.py --workload c4l --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_c4l.json 2> gpurun_out/r02a_bench_c4l.err; echo "c4l rc=$?"
# launch list of one c4l step (per-kernel share) and one full capture of the SmagorinskyKernel
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_c4l.csv \
    python bench.py --workload c4l --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02a_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:Smagorinsky -c 1 -o gpurun_out/r02a_prof_smag \
    python bench.py --workload c4l --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02a_ncu_smag.log 2>&1; echo "ncu smag rc=$?"
