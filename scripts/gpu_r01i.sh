mkdir -p gpurun_out
set -x
timeout 600 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_c3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_entry -s 15 -c 5 -o gpurun_out/prof_c3_i python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c3_i.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_c3_i.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_i.csv python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c3_list.log 2>&1; echo "list rc=$?"
