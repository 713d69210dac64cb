#!/bin/bash
# source-level ncu captures of the mid-sized general-case SIIB kernels
bash scripts/gpu/ncu_kernel.sh n3_spec siib_spec_kernel 296 47999 siib
bash scripts/gpu/ncu_kernel.sh n3_trieig siib_trieig 296 47999 siib
bash scripts/gpu/ncu_kernel.sh n3_cov32 siib_cov32_kernel 296 47999 siib
