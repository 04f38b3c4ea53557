mkdir -p gpurun_out
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 600 python bench.py --workload c3f32 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3f32.json 2> gpurun_out/bench_c3f32.err; echo "c3f32 rc=$?"; cat gpurun_out/bench_c3f32.json; tail -3 gpurun_out/bench_c3f32.err
