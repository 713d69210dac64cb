#!/bin/bash
# round 2, call 3: quadform GEMM kernel, 4-blocked back-transformation, spec staging; A/B against the old kernels; ncu of tridiag32
mkdir -p gpurun_out
O=gpurun_out/r2c03
timeout 300 python -m pytest tests/test_gpu_estoi_siib.py tests/test_gpu_api.py tests/test_gpu_scale.py -x -q > ${O}_pytest_siib.log 2>&1; echo "pytest siib exit $?"; tail -15 ${O}_pytest_siib.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_new.txt 2>&1; head -14 ${O}_times_new.txt
NELE_SIIB_QUAD_OLD=1 NELE_BACKTF_OLD=1 timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_old.txt 2>&1; head -8 ${O}_times_old.txt
timeout 300 python scripts/kernel_times.py 4096 48000 > ${O}_times_bench.txt 2>&1; head -30 ${O}_times_bench.txt
bash scripts/gpu/ncu_kernel.sh r2c03_tridiag32 tridiag32 592 47999 siib 1
bash scripts/gpu/ncu_kernel.sh r2c03_quadform quadform 592 47999 siib 1
bash scripts/gpu/ncu_kernel.sh r2c03_backtf4 backtf4 592 47999 siib 1
