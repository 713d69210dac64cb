#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'siib_jacobi2' -s 2 -c 1 -o /tmp/prof_jac python scripts/prof_batch.py 592 48000 > gpurun_out/ncu_j.log 2>&1
ncu -i /tmp/prof_jac.ncu-rep --page source --csv --print-source sass > gpurun_out/jac_source_sass.csv 2>gpurun_out/jac_source.err
ncu -i /tmp/prof_jac.ncu-rep --page details > gpurun_out/jac_details.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'haspi_ear' -s 1 -c 1 -o /tmp/prof_ear python scripts/prof_batch.py 592 48000 > gpurun_out/ncu_e.log 2>&1
ncu -i /tmp/prof_ear.ncu-rep --page source --csv --print-source sass > gpurun_out/ear_source_sass.csv 2>gpurun_out/ear_source.err
ncu -i /tmp/prof_ear.ncu-rep --page details > gpurun_out/ear_details.txt 2>&1
ls -la gpurun_out
