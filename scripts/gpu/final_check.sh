#!/bin/bash
# what the driver runs at round end -- GPU suite, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 ${O}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "bench exit $?"; python - <<'PY'
import json
b = json.loads([l for l in open("gpurun_out/final_bench_default.json").read().splitlines() if l.startswith("{")][-1])
print("value %.0f (%.2f ms) e2e %.0f general %.0f / %.0f launches %d clocks %s cpu %.1f" % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["general_case"]["value"], b["general_case"]["e2e"]["value"], b["gpu_launches"], b["clocks"], b["cpu_baseline"]["value"]))
PY
