#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c29_bench_n2.json 2> gpurun_out/c29_bench_n2.err; echo "bench n2 exit $?"; tail -3 gpurun_out/c29_bench_n2.err; cat gpurun_out/c29_bench_n2.json
