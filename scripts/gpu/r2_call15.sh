#!/bin/bash
# round 2, call 15: full suite on the final code (joblib-pool test, trieig reciprocals, PCM-16 expansion on the compute stream), bench line
mkdir -p gpurun_out
O=gpurun_out/r2c15
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 ${O}_pytest.log
timeout 300 python scripts/kernel_times.py 1024 47999 siib > ${O}_times_siib.txt 2>&1; head -8 ${O}_times_siib.txt
timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"; tail -3 ${O}_bench.err; python - <<'PY'
import json
b = json.load(open("gpurun_out/r2c15_bench.json"))
print("bench: value %.0f ms %.2f e2e %.0f (%.2f ms; f32 host %.0f) general %.0f (%.1f ms) e2e %.0f cpu %.1f nulldrop %d" % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["e2e"]["ms_per_step"], b["e2e"]["float32_host"]["value"], b["general_case"]["value"], b["general_case"]["ms_per_step"], b["general_case"]["e2e"]["value"], b["cpu_baseline"]["value"], b["pairs_siib_nullspace_dropped"]))
PY
python -c "import __graft_entry__ as g; g.smoke()"
