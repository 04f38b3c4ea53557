mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "stretched or golden or C4" > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_g.log
timeout 300 python bench.py --workload c4s --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4s_g.json 2> gpurun_out/bench_c4s_g.err; echo "c4s rc=$?"; cat gpurun_out/bench_c4s_g.json; tail -3 gpurun_out/bench_c4s_g.err
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4_g.json 2> gpurun_out/bench_c4_g.err; echo "c4 rc=$?"; cat gpurun_out/bench_c4_g.json; tail -3 gpurun_out/bench_c4_g.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"march_entry|AmdKernel" -s 12 -c 6 -o gpurun_out/prof_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c4.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_c4.log; ls -la gpurun_out
