#!/bin/bash
# round 2, call 6: tridiag32 with unmasked diagonal tiles; new parity tests (toy corpus, HL != 0, seed distribution, ragged in-loop round)
mkdir -p gpurun_out
O=gpurun_out/r2c06
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -25 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -13 ${O}_times_siib.txt
free -g | head -2; nproc
