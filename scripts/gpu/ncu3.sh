#!/bin/bash
# source-level ncu captures of the mid-sized general-case SIIB kernels
bash scripts/gpu/ncu_kernel.sh n3_vad siib_vad_kernel 296 47999 siib
bash scripts/gpu/ncu_kernel.sh n3_cov "siib_cov_kernel" 296 47999 siib
