#!/bin/bash
mkdir -p gpurun_out
K=${1:-haspi_prep}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o /tmp/prof_k python scripts/kernel_times.py 1024 48000 ${2:-haspi} > gpurun_out/c25_ncu.log 2>&1
ncu -i /tmp/prof_k.ncu-rep --page details > gpurun_out/c25_${K}_details.txt 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page source --csv > gpurun_out/c25_${K}_source.csv 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page raw --csv > gpurun_out/c25_${K}_raw.csv 2>/dev/null
grep -E "Duration|Issue Slots Busy|Executed Ipc Active|Achieved Occupancy|Registers Per|No Eligible|FP64 is|Warp Cycles Per Issued" gpurun_out/c25_${K}_details.txt | head
wc -l gpurun_out/c25_${K}_source.csv
