#!/bin/bash
# round-1 evidence run, state s6 (chol lower triangle, expand coalesced mirror, gram triangle, mask batching): tests, bench,
# launch list of the bench command, one full ncu step at bench size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c28_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/c28_pytest.log
timeout 1200 python bench.py > gpurun_out/c28_bench_n1.json 2> gpurun_out/c28_bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/c28_bench_n1.err
cat gpurun_out/c28_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c28_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c28_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none -s 34 -c 34 -o /tmp/prof_step python scripts/prof_batch.py 4096 48000 2 > gpurun_out/c28_ncu.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/c28_step_raw.csv 2>gpurun_out/c28_step_raw.err
python scripts/ncu_summary.py gpurun_out/c28_step_raw.csv "ncu --set full --clock-control none, one step, 4096 x 48000 (bench workload)" > gpurun_out/c28_ncu_summary.txt; cat gpurun_out/c28_ncu_summary.txt
python scripts/make_traffic.py gpurun_out/c28_step_raw.csv 4096 3.0 gpurun_out/c28_traffic.json | tail -3
timeout 600 python scripts/kernel_times.py 1024 52345 siib > gpurun_out/c28_fullrank_times.txt 2>&1; head -14 gpurun_out/c28_fullrank_times.txt
