#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -s 21 -c 21 -o /tmp/prof_r01b python scripts/prof_batch.py 1024 48000 > gpurun_out/ncu_f2.log 2>&1
ncu -i /tmp/prof_r01b.ncu-rep --page raw --csv > gpurun_out/raw_r01b_1024x48000.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 21 --csv --log-file gpurun_out/launches_r01b_1024x48000.csv python scripts/prof_batch.py 1024 48000 > gpurun_out/ncu_l3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'haspi_ear|siib_cov' -s 2 -c 2 --csv --page raw --log-file gpurun_out/traffic_4096x48000.csv python scripts/prof_batch.py 4096 48000 > gpurun_out/ncu_t.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ls -la gpurun_out | tail -8
