#!/bin/bash
# SIIB suite + per-kernel times of the general case and of the headline workload after a change to a SIIB kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_estoi_siib.py -q -x > gpurun_out/siibchk_pytest.log 2>&1; echo "pytest exit $? $(tail -1 gpurun_out/siibchk_pytest.log)"
timeout 200 python scripts/kernel_times.py 1024 47999 siib 2>&1 | tee gpurun_out/siibchk_times.txt | head -16
timeout 200 python scripts/kernel_times.py 4096 48000 siib 2>&1 | tee gpurun_out/siibchk_times_headline.txt | head -22
