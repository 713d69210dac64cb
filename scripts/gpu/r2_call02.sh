#!/bin/bash
# round 2, call 2: first run of the FP32 lower-triangle tridiagonalisation (siib_klt.cu): SIIB tests, A/B vs FP64, times; mma.sync tf32 rate
mkdir -p gpurun_out
O=gpurun_out/r2c02
./scripts/micro/mma_tf32 > ${O}_mma_tf32.txt 2>&1; cat ${O}_mma_tf32.txt
timeout 300 python -m pytest tests/test_gpu_estoi_siib.py tests/test_gpu_api.py -x -q > ${O}_pytest_siib.log 2>&1; echo "pytest siib exit $?"; tail -15 ${O}_pytest_siib.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_f32.txt 2>&1; head -12 ${O}_times_f32.txt
NELE_TRIDIAG_F64=1 timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_f64.txt 2>&1; head -6 ${O}_times_f64.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/kernel_times.py 8 47999 siib > ${O}_memcheck.txt 2>&1; tail -5 ${O}_memcheck.txt
