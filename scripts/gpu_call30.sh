#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c30_bench_n8.json 2> gpurun_out/c30_bench_n8.err; echo "bench n8 exit $?"; tail -3 gpurun_out/c30_bench_n8.err; cat gpurun_out/c30_bench_n8.json
