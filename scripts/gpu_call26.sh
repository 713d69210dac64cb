#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_estoi_siib.py tests/test_gpu_api.py tests/test_gpu_scale.py -x -q > gpurun_out/c26_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/c26_pytest.log
timeout 600 python scripts/kernel_times.py 4096 48000 > gpurun_out/c26_times.txt 2>&1; head -34 gpurun_out/c26_times.txt
