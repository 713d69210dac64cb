#!/bin/bash
# state check after container re-creation: gpu tests, bench N=1, launch list, one full ncu step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/c11_pytest.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/c11_bench_n1.json 2> gpurun_out/c11_bench_n1.err; echo "bench exit $?"
tail -3 gpurun_out/c11_bench_n1.err; cat gpurun_out/c11_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c11_launches_1024.csv python scripts/prof_batch.py 1024 48000 > gpurun_out/c11_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none -s 25 -c 30 -o /tmp/prof_step python scripts/prof_batch.py 1024 48000 > gpurun_out/c11_ncu.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/c11_step_raw.csv 2>gpurun_out/c11_step_raw.err
python scripts/ncu_summary.py gpurun_out/c11_step_raw.csv "ncu --set full, one step, 1024x48000" > gpurun_out/c11_ncu_summary.txt; cat gpurun_out/c11_ncu_summary.txt
