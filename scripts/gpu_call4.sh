#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
# launch list of one whole step (second call), both workloads
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 20 --csv --log-file gpurun_out/launches_1024x48000.csv python scripts/prof_batch.py 1024 48000 > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 20 --csv --log-file gpurun_out/launches_592x52345.csv python scripts/prof_batch.py 592 52345 > gpurun_out/ncu_l2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -s 20 -c 20 -o gpurun_out/prof_all_592x52345 python scripts/prof_batch.py 592 52345 > gpurun_out/ncu_f1.log 2>&1
ls -la gpurun_out
