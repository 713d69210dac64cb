#!/bin/bash
# projection route for periodic pairs: SIIB tests, then A/B of the bench step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_estoi_siib.py tests/test_gpu_api.py -x -q > gpurun_out/c18_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/c18_pytest.log
timeout 300 python scripts/kernel_times.py 4096 48000 2>&1 | tee gpurun_out/c18_kt_proj.log | head -40
NELE_SIIB_QUADFORM=1 timeout 300 python scripts/kernel_times.py 4096 48000 2>&1 | tee gpurun_out/c18_kt_quad.log | head -12
