mkdir -p gpurun_out
for wl in c4 c1; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; cat gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
