#!/bin/bash
# the default bench line with the per-step wall-time trace of the e2e loop on stderr
mkdir -p gpurun_out
NELE_BENCH_TRACE=1 NELE_TRACE=1 timeout 600 python bench.py --no-general-case --steps 6 --warmup 3 > gpurun_out/trace_bench.json 2> gpurun_out/trace_bench.err; echo "bench exit $?"
grep -v "^\[nele\]" gpurun_out/trace_bench.err | tail -5
grep "^\[nele\]" gpurun_out/trace_bench.err | tail -40
python - <<'PY'
import json
b = json.loads([l for l in open("gpurun_out/trace_bench.json").read().splitlines() if l.startswith("{")][-1])
print("value %.0f e2e %.0f (%.1f ms) f32 %.0f (%.1f ms) cpu %.1f" % (b["value"], b["e2e"]["value"], b["e2e"]["ms_per_step"], b["e2e"]["float32_host"]["value"], b["e2e"]["float32_host"]["ms_per_step"], b["cpu_baseline"]["value"]))
PY
