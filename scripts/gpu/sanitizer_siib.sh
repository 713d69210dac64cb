#!/bin/bash
# memcheck / racecheck of the SIIB kernels only (general and periodic case), bounded: for late changes to siib*.cu
mkdir -p gpurun_out
O=gpurun_out/san2
timeout 70 compute-sanitizer --tool memcheck python scripts/kernel_times.py 3 47999 siib > ${O}_memcheck_general.txt 2>&1; echo "memcheck general exit $?"; grep -E "ERROR SUMMARY|L=47999" ${O}_memcheck_general.txt | tail -2
timeout 50 compute-sanitizer --tool memcheck python scripts/kernel_times.py 3 48000 siib > ${O}_memcheck_periodic.txt 2>&1; echo "memcheck periodic exit $?"; grep -E "ERROR SUMMARY|L=48000" ${O}_memcheck_periodic.txt | tail -2
timeout 110 compute-sanitizer --tool racecheck python scripts/kernel_times.py 2 47999 siib > ${O}_racecheck_general.txt 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|L=47999" ${O}_racecheck_general.txt | tail -2
