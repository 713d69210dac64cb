#!/bin/bash
# usage: scripts/gpu/ncu_kernel.sh <tag> <kernel regex> <n> <L> <metrics> [skip]
# One `ncu --set full` capture of one kernel of a kernel_times.py run; leaves details / raw csv / source csv and
# a one-line summary under gpurun_out/<tag>_*.
tag=$1; K=$2; n=$3; L=$4; met=$5; skip=${6:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $skip -c 1 -o /tmp/prof_$tag python scripts/kernel_times.py $n $L $met > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_raw.csv "$K n=$n L=$L" > gpurun_out/${tag}_summary.txt 2>&1
cat gpurun_out/${tag}_summary.txt
grep -E "Duration|Issue Slots Busy|Executed Ipc Active|Achieved Occupancy|Registers Per|No Eligible|Warp Cycles Per Issued|L2 Hit|DRAM Throughput|L2 Cache Throughput|Mem Busy|Max Bandwidth" gpurun_out/${tag}_details.txt | head -20
