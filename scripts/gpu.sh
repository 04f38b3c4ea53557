#!/bin/bash
# One parametrised runner for every GPU-side job of this repo (run under gpurun; outputs land in gpurun_out/).
#   gpurun --timeout 900 -- 'bash scripts/gpu.sh <tag> <task> [<task> ...]'
# tasks:
#   tests                 python -m pytest tests -m gpu
#   smoke                 __graft_entry__.smoke()
#   fuzz:<seed>:<cases>   scripts/fuzz_parity.py --cuda (random configurations, CUDA library vs oracle)   -> <tag>_fuzz_cuda.log
#   bench:<wl>[:steps]    python bench.py --workload <wl> (no e2e / cpu baseline)  -> <tag>_bench_<wl>.json
#   default               python bench.py exactly as the driver runs it            -> <tag>_bench_default.json
#   reference             python bench.py --impl reference                         -> <tag>_bench_reference.json
#   list:<wl>             ncu launch list (gpu__time_duration) of one step         -> <tag>_launches_<wl>.csv
#   ncu:<wl>:<regex>:<skip>:<count>   ncu --set full --import-source on            -> <tag>_ncu_<wl>_<n>.ncu-rep
#   disttests             NCCL parity tests on the box's GPUs (gpurun --gpus N)
#   scale:<N>[:steps]     torchrun bench.py --gpus N                                -> <tag>_bench_n<N>.json
#   memcheck | racecheck  compute-sanitizer on smoke()                             -> <tag>_<tool>.log
#   py:<script>           python <script> (free-form experiments under scripts/)
set -u
mkdir -p gpurun_out
TAG=$1; shift
n=0
for task in "$@"; do
  IFS=: read -r kind a b c d <<< "$task"
  echo "=== $task"
  case $kind in
    fuzz)            # the configuration fuzzer on the CUDA library: fuzz:<seed>:<cases>
      timeout 1500 python scripts/fuzz_parity.py --cuda --seed ${a:-7} --cases ${b:-60} > gpurun_out/${TAG}_fuzz_cuda.log 2>&1; echo "fuzz rc=$?"; tail -6 gpurun_out/${TAG}_fuzz_cuda.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${TAG}_smoke.log ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest_gpu.log ;;
    bench)
      timeout 600 python bench.py --workload $a --steps ${b:-5} --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-other-workloads > gpurun_out/${TAG}_bench_$a.json 2> gpurun_out/${TAG}_bench_$a.err
      echo "bench $a rc=$?"; cat gpurun_out/${TAG}_bench_$a.json; tail -3 gpurun_out/${TAG}_bench_$a.err ;;
    default)
      timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "default rc=$?"; cat gpurun_out/${TAG}_bench_default.json; tail -3 gpurun_out/${TAG}_bench_default.err ;;
    reference)
      timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"; cat gpurun_out/${TAG}_bench_reference.json ;;
    list)
      timeout 300 python bench.py --workload $a --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity-check --no-other-workloads > gpurun_out/${TAG}_plain_$a.log 2>&1 &&
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_$a.csv \
          python bench.py --workload $a --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity-check --no-other-workloads > gpurun_out/${TAG}_ncu_list_$a.log 2>&1; echo "list $a rc=$?" ;;
    ncu)
      n=$((n+1))
      timeout 300 python bench.py --workload $a --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity-check --no-other-workloads > gpurun_out/${TAG}_plain_$a.log 2>&1 &&
      timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$b" -s ${c:-0} -c ${d:-3} -f \
          -o gpurun_out/${TAG}_ncu_${a}_$n python bench.py --workload $a --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity-check --no-other-workloads > gpurun_out/${TAG}_ncu_${a}_$n.log 2>&1
      echo "ncu $a /$b/ rc=$?"; tail -3 gpurun_out/${TAG}_ncu_${a}_$n.log
      # the .ncu-rep with imported source is ~10 MB per launch and gpurun_out/ is capped at 64 MiB: export the pages here, drop the report
      rep=gpurun_out/${TAG}_ncu_${a}_$n.ncu-rep
      if [ -f $rep ]; then
        ncu -i $rep --page raw --csv > gpurun_out/${TAG}_ncu_${a}_${n}_raw.csv 2>/dev/null
        for ((i=0; i<${d:-3}; i++)); do
          ncu -i $rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 2>/dev/null | gzip -9 > gpurun_out/${TAG}_ncu_${a}_${n}_source_$i.csv.gz
        done
        [ "${KEEP_REP:-0}" = 1 ] || rm -f $rep
      fi ;;
    disttests)       # NCCL parity (tests/test_distributed.py -m gpu): needs gpurun --gpus N
      OC_NCCL_NEW_CASES=1 timeout 1500 python -m pytest tests/test_distributed.py -m gpu -q -rs > gpurun_out/${TAG}_pytest_dist.log 2>&1; echo "disttests rc=$?"; tail -8 gpurun_out/${TAG}_pytest_dist.log ;;
    scale)           # bench.py under torchrun on $a GPUs, as the driver launches it
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $a --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $a --steps ${b:-5} --warmup 3 \
          > gpurun_out/${TAG}_bench_n$a.json 2> gpurun_out/${TAG}_bench_n$a.err; echo "scale $a rc=$?"; cat gpurun_out/${TAG}_bench_n$a.json; tail -3 gpurun_out/${TAG}_bench_n$a.err ;;
    memcheck|racecheck)
      timeout 1200 compute-sanitizer --tool $kind --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_$kind.log 2>&1; echo "$kind rc=$?"; tail -6 gpurun_out/${TAG}_$kind.log ;;
    py)
      timeout 1500 python $a > gpurun_out/${TAG}_$(basename $a .py).log 2>&1; echo "py $a rc=$?"; tail -40 gpurun_out/${TAG}_$(basename $a .py).log ;;
    *) echo "unknown task $task" ;;
  esac
done
ls -la gpurun_out | tail -30
