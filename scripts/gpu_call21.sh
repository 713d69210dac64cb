#!/bin/bash
# feature front-end parity + the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py -q > gpurun_out/c21_feat_pytest.log 2>&1; echo "feat pytest exit $?"; tail -25 gpurun_out/c21_feat_pytest.log | cut -c1-600
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_features.py -k nothing > gpurun_out/c21_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/c21_pytest.log
