#!/bin/bash
# round-1 evidence run (state s5: feature front-end, ESTOI resampler fast path, prep/smalltri passes): tests, both bench arms,
# launch list of the bench command, one full ncu step at bench size, feature front-end probe + ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c27_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/c27_pytest.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c27_bench_ref.json 2> gpurun_out/c27_bench_ref.err; echo "ref exit $?"
timeout 1200 python bench.py > gpurun_out/c27_bench_n1.json 2> gpurun_out/c27_bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/c27_bench_n1.err
cat gpurun_out/c27_bench_ref.json gpurun_out/c27_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c27_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c27_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none -s 34 -c 34 -o /tmp/prof_step python scripts/prof_batch.py 4096 48000 2 > gpurun_out/c27_ncu.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/c27_step_raw.csv 2>gpurun_out/c27_step_raw.err
python scripts/ncu_summary.py gpurun_out/c27_step_raw.csv "ncu --set full --clock-control none, one step, 4096 x 48000 (bench workload)" > gpurun_out/c27_ncu_summary.txt; cat gpurun_out/c27_ncu_summary.txt
python scripts/make_traffic.py gpurun_out/c27_step_raw.csv 4096 3.0 gpurun_out/c27_traffic.json | tail -5
timeout 300 python scripts/feature_times.py 4096 48000 8 > gpurun_out/c27_feature_times.json 2> gpurun_out/c27_feature_times.err; cat gpurun_out/c27_feature_times.json
timeout 600 ncu --set full --clock-control none -k regex:feat_ -s 3 -c 3 -o /tmp/prof_feat python scripts/feature_times.py 4096 48000 0 > gpurun_out/c27_feat_ncu.log 2>&1
ncu -i /tmp/prof_feat.ncu-rep --page raw --csv > gpurun_out/c27_feat_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c27_feat_raw.csv "ncu --set full --clock-control none, nele_features, 4096 x 48000" > gpurun_out/c27_feat_summary.txt; cat gpurun_out/c27_feat_summary.txt
