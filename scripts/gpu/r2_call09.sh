#!/bin/bash
# round 2, call 9: backtf5 f32x2, trieig with fewer instructions; bench configs ganround / sweep at N = 1 (validation of bench.py)
mkdir -p gpurun_out
O=gpurun_out/r2c09
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -13 ${O}_times_siib.txt
timeout 600 python bench.py --config ganround --steps 5 --warmup 2 > ${O}_ganround_n1.json 2> ${O}_ganround_n1.err; echo "ganround exit $?"; tail -3 ${O}_ganround_n1.err; cat ${O}_ganround_n1.json
timeout 900 python bench.py --config sweep --sweep-pairs 8192 --steps 1 --warmup 1 > ${O}_sweep8k_n1.json 2> ${O}_sweep8k_n1.err; echo "sweep exit $?"; tail -3 ${O}_sweep8k_n1.err; cat ${O}_sweep8k_n1.json
