#!/bin/bash
# round 2, call 20: full suite with concurrent metric pipelines on small chunks (new default); GAN round latency at N = 1
mkdir -p gpurun_out
O=gpurun_out/r2c20
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 ${O}_pytest.log
timeout 300 compute-sanitizer --tool racecheck python scripts/quick_time.py 3 2.1 haspi,siib,estoi > ${O}_racecheck.txt 2>&1; tail -2 ${O}_racecheck.txt
timeout 600 python bench.py --config ganround --steps 5 --warmup 2 > ${O}_ganround_n1.json 2> ${O}_ganround_n1.err; echo "ganround exit $?"; cut -c1-200 ${O}_ganround_n1.json; python -c "
import json; d=json.load(open('${O}_ganround_n1.json')); print(d['latency_ms'], d['kernel_ms_per_round_max_rank'])"
python -c "import __graft_entry__ as g; g.smoke()"
