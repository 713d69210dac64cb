#!/bin/bash
# round 2, call 12: backtf6 (warp-independent) vs backtf5 / backtf4
mkdir -p gpurun_out
O=gpurun_out/r2c12
timeout 300 python -m pytest tests/test_gpu_estoi_siib.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -8 ${O}_times_siib.txt
NELE_BACKTF5=1 timeout 300 python scripts/kernel_times.py 1024 47999 siib 2>&1 | head -4
bash scripts/gpu/ncu_kernel.sh r2c12_backtf6 backtf6 592 47999 siib 1
