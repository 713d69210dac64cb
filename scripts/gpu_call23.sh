#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:estoi_resample58 -s 1 -c 1 -o /tmp/prof_rs python scripts/kernel_times.py 1024 48000 estoi > gpurun_out/c23_ncu.log 2>&1
ncu -i /tmp/prof_rs.ncu-rep --page raw --csv > gpurun_out/c23_rs_raw.csv 2>/dev/null
ncu -i /tmp/prof_rs.ncu-rep --page details > gpurun_out/c23_rs_details.txt 2>/dev/null
grep -E "Duration|Issue Slots Busy|Executed Ipc|Achieved Occupancy|Theoretical Occupancy|Registers|Stall|Bank|FP64|L1/TEX Hit|Warp Cycles Per Issued|No Eligible" gpurun_out/c23_rs_details.txt | head -40
