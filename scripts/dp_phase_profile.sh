#!/bin/bash
# Phase breakdown of the DP row loop from in-kernel clocks (lane 0 of every warp of CTA 0), printed per alignment.
# Builds an instrumented copy of the library (-DSVS_DP_PROFILE) next to the product one; run on the GPU box:
#   bash scripts/dp_phase_profile.sh build      (here, no GPU needed)
#   NWIN=148 CFGS=384,8,2 bash scripts/dp_phase_profile.sh run > gpurun_out/dpprof.log
set -e
cd "$(dirname "$0")/.."
C=svscope_b200/csrc; O=$C/_build_prof; mkdir -p $O
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -DSVS_DP_PROFILE"
if [ "$1" = build ]; then
  for f in api poa_kernels poa_window msa_features em myers misscore; do nvcc $FL -c $C/$f.cu -o $O/$f.o & done; wait
  nvcc -shared -o svscope_b200/_C/libsvscope_b200_prof.so $O/*.o -lcudart -lpthread
  echo built svscope_b200/_C/libsvscope_b200_prof.so
else
  SVS_LIB=$PWD/svscope_b200/_C/libsvscope_b200_prof.so python scripts/poa_probe.py
fi
