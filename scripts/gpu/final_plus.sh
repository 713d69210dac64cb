#!/bin/bash
# final check (what the driver runs) + the per-kernel times and the ncu launch list that profiles/ quotes
bash scripts/gpu/final_check.sh
timeout 120 python scripts/kernel_times.py 4096 48000 > gpurun_out/final_times_4096x48000.txt 2>&1; head -3 gpurun_out/final_times_4096x48000.txt
timeout 120 python scripts/kernel_times.py 1024 47999 siib > gpurun_out/final_times_siib_1024x47999.txt 2>&1; head -2 gpurun_out/final_times_siib_1024x47999.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_launch.log 2>&1; wc -l gpurun_out/final_launches_bench.csv
