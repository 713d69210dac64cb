#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/dbg_metrics.py > gpurun_out/dbg_metrics.log 2>&1
echo "exit $?" >> gpurun_out/dbg_metrics.log
grep -v "^  n10\|kept equal\|tob rel\|x10 err\|logspec err\|Sxx rel\|Fa oracle" gpurun_out/dbg_metrics.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
