mkdir -p gpurun_out
set -x
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
