#!/bin/bash
# ncu --set full of the full-rank Jacobi (cluster kernel), the rank-96 Jacobi, the FP64/FP32 lag-product kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'siib_jacobi2' -c 2 -o /tmp/prof_jc python scripts/prof_batch.py 296 47999 1 > gpurun_out/c12_ncu_jc.log 2>&1
ncu -i /tmp/prof_jc.ncu-rep --page raw --csv > gpurun_out/c12_jc_raw.csv 2>/dev/null
ncu -i /tmp/prof_jc.ncu-rep --page details > gpurun_out/c12_jc_details.txt 2>&1
ncu -i /tmp/prof_jc.ncu-rep --page source --csv --print-source sass > gpurun_out/c12_jc_sass.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'siib_jacobi2|siib_cov|siib_chol|siib_quad|siib_expand' -c 7 -o /tmp/prof_s96 python scripts/prof_batch.py 592 48000 1 > gpurun_out/c12_ncu_s96.log 2>&1
ncu -i /tmp/prof_s96.ncu-rep --page raw --csv > gpurun_out/c12_s96_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c12_s96_raw.csv "592x48000 siib kernels" ; python scripts/ncu_summary.py gpurun_out/c12_jc_raw.csv "296x47999 jacobi"
ls -la gpurun_out | tail -8
