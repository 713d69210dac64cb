#!/bin/bash
# round 2, call 22: occupancy caps on the warp-per-pair HASPI kernels (wave quantisation at 4096 pairs)
mkdir -p gpurun_out
O=gpurun_out/r2c22
timeout 300 python -m pytest tests/test_gpu_haspi.py -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 4096 48000 haspi > ${O}_times_haspi.txt 2>&1; head -8 ${O}_times_haspi.txt
timeout 300 python scripts/kernel_times.py 3000 61234 haspi 2>&1 | head -5
