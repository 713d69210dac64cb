#!/bin/bash
# round 2, call 16 (8 GPUs): the driver's weak-scaling line at N = 8 and N = 1 on the same box with e2e from int16 host buffers
mkdir -p gpurun_out
O=gpurun_out/r2c16
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29908 bench.py --gpus 8 --steps 10 --warmup 3 > ${O}_bench_n8.json 2> ${O}_bench_n8.err; echo "bench N=8 exit $?"; tail -2 ${O}_bench_n8.err | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench N=1 exit $?"
python - <<'PY'
import json
for n in (8, 1):
    b = json.loads([l for l in open("gpurun_out/r2c16_bench_n%d.json" % n).read().splitlines() if l.startswith("{")][-1])
    print("N=%d: value %.0f (%.2f ms) e2e %.0f (%.2f ms; f32 host %.0f, %.2f ms) general %.0f e2e %.0f" % (n, b["value"], b["ms_per_step"], b["e2e"]["value"], b["e2e"]["ms_per_step"], b["e2e"]["float32_host"]["value"], b["e2e"]["float32_host"]["ms_per_step"], b["general_case"]["value"], b["general_case"]["e2e"]["value"]))
PY
