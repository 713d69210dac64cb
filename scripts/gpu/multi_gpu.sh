#!/bin/bash
# BASELINE configs[3] (GAN round latency) and configs[4] (3-10 s sweep, strong scaling) at N = 1, 2, 4, 8;
# the driver's weak-scaling line at N = 8; score_sharded on real engines (pytest -m gpu includes the 2-rank test)
mkdir -p gpurun_out
O=gpurun_out/mgpu
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > ${O}_topo.txt 2>&1
for N in 8 4 2; do
  timeout 900 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --config sweep --steps 1 > ${O}_sweep_n$N.json 2> ${O}_sweep_n$N.err; echo "sweep N=$N exit $?"; tail -2 ${O}_sweep_n$N.err | cut -c1-300; cat ${O}_sweep_n$N.json | cut -c1-900
done
timeout 900 python bench.py --config sweep --steps 1 > ${O}_sweep_n1.json 2> ${O}_sweep_n1.err; echo "sweep N=1 exit $?"; tail -2 ${O}_sweep_n1.err | cut -c1-300; cat ${O}_sweep_n1.json | cut -c1-900
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N --config ganround --steps 5 --warmup 2 > ${O}_ganround_n$N.json 2> ${O}_ganround_n$N.err; echo "ganround N=$N exit $?"; tail -2 ${O}_ganround_n$N.err | cut -c1-300; cat ${O}_ganround_n$N.json | cut -c1-700
done
timeout 900 $TR --nproc-per-node 8 --master-port 29808 bench.py --gpus 8 --steps 5 --warmup 3 > ${O}_bench_n8.json 2> ${O}_bench_n8.err; echo "bench N=8 exit $?"; tail -2 ${O}_bench_n8.err | cut -c1-300; cat ${O}_bench_n8.json | cut -c1-1500
timeout 600 python -m pytest tests/test_gpu_scale.py -q > ${O}_pytest_scale.log 2>&1; tail -3 ${O}_pytest_scale.log
