#!/bin/bash
# round 2, call 8: backtf5 with padded row-part stride; ncu --set full of every kernel of a general-case step (592 pairs)
mkdir -p gpurun_out
O=gpurun_out/r2c08
timeout 300 python -m pytest tests/test_gpu_estoi_siib.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -13 ${O}_times_siib.txt
timeout 1200 ncu --set full --clock-control none -c 40 -o /tmp/prof_step python scripts/kernel_times.py 592 47999 > ${O}_ncu.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > ${O}_step_raw.csv 2>/dev/null
python scripts/ncu_summary.py ${O}_step_raw.csv "general-case step, 592 x 47999, first iteration (cold)" > ${O}_step_summary.txt; cat ${O}_step_summary.txt
