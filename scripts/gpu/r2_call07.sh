#!/bin/bash
# round 2, call 7: backtf5 (register tile), 46 bisection steps; full suite; first bench.py line with general_case
mkdir -p gpurun_out
O=gpurun_out/r2c07
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -25 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -13 ${O}_times_siib.txt
NELE_BACKTF4=1 timeout 300 python scripts/kernel_times.py 1024 47999 siib 2>&1 | head -4
bash scripts/gpu/ncu_kernel.sh r2c07_backtf5 backtf5 592 47999 siib 1
timeout 900 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"; tail -3 ${O}_bench.err; cat ${O}_bench.json
