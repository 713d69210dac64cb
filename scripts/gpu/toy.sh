#!/bin/bash
# BASELINE configs[1]: the toy corpus on one B200 against the CPU oracle
mkdir -p gpurun_out
timeout 600 python bench.py --config toy --steps 10 --warmup 3 > gpurun_out/toy_n1.json 2> gpurun_out/toy_n1.err; echo "exit $?"; tail -2 gpurun_out/toy_n1.err; cat gpurun_out/toy_n1.json
