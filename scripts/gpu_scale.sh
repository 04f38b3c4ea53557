mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "n$N rc=$?"; cat gpurun_out/bench_c3_n$N.json; tail -5 gpurun_out/bench_c3_n$N.err
