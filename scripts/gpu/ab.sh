#!/bin/bash
# A/B references of DESIGN.md section 9: times / score deviations of every switch, and the SIIB + HASPI suites under each
mkdir -p gpurun_out
O=gpurun_out/ab
timeout 900 python scripts/ab_switches.py 1024 47999 > ${O}_switches_1024x47999.txt 2>&1; cat ${O}_switches_1024x47999.txt
for sw in NELE_TRIDIAG_F64=1 NELE_BACKTF_OLD=1 NELE_BACKTF4=1 NELE_BACKTF5=1 NELE_COV_XX64=1 NELE_F32X2=0 NELE_RESAMPLE_F32=0 NELE_CONCURRENT=0 NELE_CONCURRENT=1; do
  env $sw timeout 300 python -m pytest tests/test_gpu_estoi_siib.py tests/test_gpu_haspi.py -q > ${O}_pytest_${sw}.log 2>&1; echo "$sw: exit $? $(tail -1 ${O}_pytest_${sw}.log)"
done
