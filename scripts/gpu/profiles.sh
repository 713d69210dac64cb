#!/bin/bash
# the evidence set for profiles/ -- bench line (driver settings), reference arm, launch list, ncu --set full of a
# whole step at bench size and of a general-case step, DRAM traffic per kernel
mkdir -p gpurun_out
O=gpurun_out/prof
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench exit $?"; tail -2 ${O}_bench_n1.err; cut -c1-600 ${O}_bench_n1.json
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > ${O}_bench_reference_arm.json 2> ${O}_bench_ref.err; echo "ref exit $?"; cut -c1-700 ${O}_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_launch.log 2>&1; wc -l ${O}_launches_bench.csv
timeout 1200 ncu --set full --clock-control none -c 80 -o /tmp/prof_bench python scripts/prof_batch.py 4096 48000 2 > ${O}_ncu_bench.log 2>&1
ncu -i /tmp/prof_bench.ncu-rep --page raw --csv > ${O}_step_bench_raw.csv 2>/dev/null
python scripts/ncu_summary.py ${O}_step_bench_raw.csv "one step, 4096 x 48000 (bench workload), two calls (first cold)" > ${O}_ncu_full_4096x48000.txt; cat ${O}_ncu_full_4096x48000.txt
python scripts/make_traffic.py ${O}_step_bench_raw.csv 4096 3.0 ${O}_traffic.json > /dev/null
timeout 1200 ncu --set full --clock-control none -c 80 -o /tmp/prof_gen python scripts/prof_batch.py 4096 47999 2 > ${O}_ncu_gen.log 2>&1
ncu -i /tmp/prof_gen.ncu-rep --page raw --csv > ${O}_step_general_raw.csv 2>/dev/null
python scripts/ncu_summary.py ${O}_step_general_raw.csv "one step, 4096 x 47999 (general case), two calls (first cold)" > ${O}_ncu_full_4096x47999.txt; cat ${O}_ncu_full_4096x47999.txt
python scripts/make_traffic.py ${O}_step_general_raw.csv 4096 2.9999375 ${O}_traffic_general.json > /dev/null
