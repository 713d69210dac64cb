#!/bin/bash
# round 2, call 11: full suite (PCM-16 path, device data set, f32 middle-ear output), bench line with PCM-16 e2e
mkdir -p gpurun_out
O=gpurun_out/r2c11
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -12 ${O}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"; tail -3 ${O}_bench.err; cat ${O}_bench.json | cut -c1-2600
timeout 300 python scripts/kernel_times.py 4096 48000 haspi > ${O}_times_haspi.txt 2>&1; head -9 ${O}_times_haspi.txt
