#!/bin/bash
# round 2, call 21 (8 GPUs): GAN round latency at N = 8 and 4 with the concurrent small-chunk default
mkdir -p gpurun_out
O=gpurun_out/r2c21
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29800+N)) bench.py --gpus $N --config ganround --steps 5 --warmup 2 > ${O}_ganround_n$N.json 2> ${O}_ganround_n$N.err; echo "ganround N=$N exit $?"
  python -c "
import json; d=json.loads([l for l in open('${O}_ganround_n$N.json').read().splitlines() if l.startswith('{')][-1]); print($N, d['latency_ms'], d['kernel_ms_per_round_max_rank']); json.dump(d, open('${O}_ganround_n$N.json','w'))"
done
