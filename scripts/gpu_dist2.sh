mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; echo "n2 rc=$?"; cat gpurun_out/bench_c3_n2.json; tail -5 gpurun_out/bench_c3_n2.err
