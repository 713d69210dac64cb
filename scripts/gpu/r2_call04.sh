#!/bin/bash
# round 2, call 4: tridiag32 fast paths, backtf4 with cp.async staging, quadform with f32x2
mkdir -p gpurun_out
O=gpurun_out/r2c04
timeout 300 python -m pytest tests/test_gpu_estoi_siib.py tests/test_gpu_api.py tests/test_gpu_scale.py -x -q > ${O}_pytest_siib.log 2>&1; echo "pytest siib exit $?"; tail -15 ${O}_pytest_siib.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_new.txt 2>&1; head -14 ${O}_times_new.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/kernel_times.py 8 47999 siib > ${O}_memcheck.txt 2>&1; tail -3 ${O}_memcheck.txt
timeout 300 compute-sanitizer --tool racecheck python scripts/kernel_times.py 4 47999 siib > ${O}_racecheck.txt 2>&1; tail -3 ${O}_racecheck.txt
bash scripts/gpu/ncu_kernel.sh r2c04_tridiag32 tridiag32 592 47999 siib 1
bash scripts/gpu/ncu_kernel.sh r2c04_backtf4 backtf4 592 47999 siib 1
bash scripts/gpu/ncu_kernel.sh r2c04_quadform quadform 592 47999 siib 1
