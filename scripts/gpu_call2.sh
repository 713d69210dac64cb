#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/dbg_metrics.py > gpurun_out/dbg_metrics.log 2>&1
echo "exit $?" >> gpurun_out/dbg_metrics.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/dbg_metrics.log
