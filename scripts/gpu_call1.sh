#!/bin/bash
# first GPU call of the session: HASPI parity tests, timing, launch list, ncu full on the ear kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/quick_time.py 4096 3 haspi > gpurun_out/time_haspi_4096.log 2>&1
NELE_HASPI_F64=1 timeout 600 python scripts/quick_time.py 1024 3 haspi > gpurun_out/time_haspi_f64_1024.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_haspi.csv python scripts/quick_time.py 1024 3 haspi > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'haspi_(ear|modcorr|control|prep|cep)' -c 5 -o gpurun_out/prof_haspi python scripts/quick_time.py 1024 3 haspi > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/time_haspi_4096.log
