#!/bin/bash
# feature front-end: parity tests, timing probe, ncu of both kernels; then the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py -x -q > gpurun_out/c20_feat_pytest.log 2>&1; echo "feat pytest exit $?"; tail -25 gpurun_out/c20_feat_pytest.log
timeout 300 python scripts/feature_times.py 4096 48000 8 > gpurun_out/c20_feature_times.json 2> gpurun_out/c20_feature_times.err; echo "times exit $?"; cat gpurun_out/c20_feature_times.json; tail -3 gpurun_out/c20_feature_times.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:feat_ -s 3 -c 3 -o /tmp/prof_feat python scripts/feature_times.py 1024 48000 0 > gpurun_out/c20_ncu.log 2>&1
ncu -i /tmp/prof_feat.ncu-rep --page raw --csv > gpurun_out/c20_feat_raw.csv 2>gpurun_out/c20_feat_raw.err
python scripts/ncu_summary.py gpurun_out/c20_feat_raw.csv "ncu --set full --clock-control none, nele_features, 1024 x 48000" > gpurun_out/c20_feat_summary.txt; cat gpurun_out/c20_feat_summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/c20_pytest.log
