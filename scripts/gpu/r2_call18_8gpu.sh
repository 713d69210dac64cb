#!/bin/bash
# round 2, call 18 (8 GPUs): where the end-to-end step spends its extra time at N = 8 (per-rank trace)
mkdir -p gpurun_out
O=gpurun_out/r2c18
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NELE_BENCH_TRACE=1 NELE_TRACE=0 timeout 400 $TR --nproc-per-node 8 --master-port 29918 bench.py --gpus 8 --steps 6 --warmup 3 --no-general-case > ${O}_bench_n8.json 2> ${O}_bench_n8.err; echo "exit $?"
grep "bench trace" ${O}_bench_n8.err
nproc; cat /proc/cpuinfo | grep "model name" | head -1; taskset -p $$ 
