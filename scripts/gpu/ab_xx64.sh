#!/bin/bash
# A/B of NELE_COV_XX64=1 (FP64 xx lag products for non-periodic pairs): SIIB suite, times, scores of 32 pairs against the default
mkdir -p gpurun_out
NELE_COV_XX64=1 timeout 200 python -m pytest tests/test_gpu_estoi_siib.py -q -x > gpurun_out/abxx_pytest.log 2>&1; echo "pytest exit $? $(tail -1 gpurun_out/abxx_pytest.log)"
timeout 200 python scripts/ab_switches.py 256 47999 default cov_xx64 > gpurun_out/abxx_switches.txt 2>&1; tail -12 gpurun_out/abxx_switches.txt
