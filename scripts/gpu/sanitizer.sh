#!/bin/bash
# racecheck / memcheck of the final KLT kernels and the resynthesis kernel on tiny batches
mkdir -p gpurun_out
O=gpurun_out/san
timeout 900 compute-sanitizer --tool racecheck python scripts/kernel_times.py 2 47999 siib > ${O}_racecheck.txt 2>&1; tail -3 ${O}_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python scripts/kernel_times.py 6 47999 > ${O}_memcheck.txt 2>&1; tail -2 ${O}_memcheck.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_api.py -q -k "inloop or dataset" > ${O}_memcheck_inloop.txt 2>&1; tail -4 ${O}_memcheck_inloop.txt
