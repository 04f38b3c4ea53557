mkdir -p gpurun_out
set -x
timeout 600 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:march_entry -s 15 -c 5 -o gpurun_out/prof_march_c3 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/ncu.log; ls -la gpurun_out
