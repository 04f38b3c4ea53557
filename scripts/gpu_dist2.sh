mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; echo "n2 rc=$?"; cat gpurun_out/bench_c3_n2.json; tail -5 gpurun_out/bench_c3_n2.err
