#!/bin/bash
# round 2, call 13: tridiag32 with cp.async tile pipeline and warp-level correction dots
mkdir -p gpurun_out
O=gpurun_out/r2c13
timeout 300 python -m pytest tests/test_gpu_estoi_siib.py tests/test_gpu_scale.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -6 ${O}_times_siib.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/kernel_times.py 8 47999 siib > ${O}_memcheck.txt 2>&1; tail -2 ${O}_memcheck.txt
bash scripts/gpu/ncu_kernel.sh r2c13_tridiag32 tridiag32 592 47999 siib 1
