#!/bin/bash
# round 2, call 1: baseline GPU suite; the three never-run switches (suite + A/B times); general-case baseline;
# attempt to obtain the un-vendored packages (network probe)
mkdir -p gpurun_out
O=gpurun_out/r2c01
timeout 300 python -m pytest tests -m gpu -x -q > ${O}_pytest_default.log 2>&1; echo "pytest default exit $?"; tail -2 ${O}_pytest_default.log
for sw in NELE_F32X2 NELE_RESAMPLE_F32 NELE_TRIDIAG_MV32; do
  env $sw=1 timeout 300 python -m pytest tests -m gpu -q > ${O}_pytest_${sw}.log 2>&1; echo "pytest $sw exit $?"; tail -3 ${O}_pytest_${sw}.log
done
timeout 600 python scripts/ab_switches.py 4096 48000 > ${O}_ab_4096x48000.txt 2>&1; cat ${O}_ab_4096x48000.txt
timeout 600 python scripts/ab_switches.py 1024 47999 > ${O}_ab_1024x47999.txt 2>&1; cat ${O}_ab_1024x47999.txt
timeout 300 python scripts/kernel_times.py 4096 47999 > ${O}_times_4096x47999.txt 2>&1; cat ${O}_times_4096x47999.txt
(timeout 60 python -m pip download --no-deps -d /tmp/whl pystoi pysiib resampy==0.2.2 2>&1; echo "pip exit $?"; timeout 10 python -c "import pystoi" 2>&1; timeout 10 python -c "import pysiib" 2>&1; timeout 10 python -c "import resampy" 2>&1; timeout 10 python -c "import librosa" 2>&1; ls /opt/wheelhouse 2>&1 | grep -i -E "stoi|siib|resampy|librosa|soxr|soundfile" ; echo "wheelhouse grep exit $?") > ${O}_pip_thirdparty.log 2>&1; tail -12 ${O}_pip_thirdparty.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
