mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c4s --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_c4s.json 2> gpurun_out/bench_c4s.err; echo "c4s rc=$?"; cat gpurun_out/bench_c4s.json; tail -3 gpurun_out/bench_c4s.err
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4s.csv python bench.py --workload c4s --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c4s.log 2>&1; echo "ncu rc=$?"
