#!/bin/bash
# round 2, call 5: tridiag32 (384 threads, constant tile lists, no edge paths), nele_resyn + ragged in-loop round, full suite
mkdir -p gpurun_out
O=gpurun_out/r2c05
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -13 ${O}_times_siib.txt
timeout 300 python scripts/kernel_times.py 4096 47999 > ${O}_times_general.txt 2>&1; head -24 ${O}_times_general.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/kernel_times.py 8 47999 siib > ${O}_memcheck.txt 2>&1; tail -3 ${O}_memcheck.txt
bash scripts/gpu/ncu_kernel.sh r2c05_tridiag32 tridiag32 592 47999 siib 1
