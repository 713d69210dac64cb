#!/bin/bash
# BASELINE configs[3] (latency of one GAN sampling round) and configs[4] sample on one GPU, final code
mkdir -p gpurun_out
timeout 60 python bench.py --config ganround --steps 5 --warmup 2 > gpurun_out/late_ganround_n1.json 2> gpurun_out/late_ganround_n1.err; echo "ganround exit $?"; cut -c1-500 gpurun_out/late_ganround_n1.json
