#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python scripts/dbg_metrics.py 1 2>&1 | grep -A6 "^batch"
timeout 1200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
tail -3 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print(d['kernels_ms_per_step']); print(d['roofline'])"
