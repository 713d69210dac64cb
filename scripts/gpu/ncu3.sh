#!/bin/bash
# source-level ncu captures of the mid-sized general-case SIIB kernels
bash scripts/gpu/ncu_kernel.sh n4_spec siib_spec_kernel 296 47999 siib
