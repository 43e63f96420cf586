#!/bin/bash
# Build a variant of the library whose photometry kernel prints per-phase cycle counts (RB_PHOT_PROFILE), for use as
#   REDBACK_B200_LIB=$PWD/redback_b200/_lib/exp/prof.so python tools/bench_paths.py --only magnitudes
set -e
cd "$(dirname "$0")/.."
mkdir -p redback_b200/_lib/exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -shared \
  -DRB_PHOT_PROFILE -I include -I redback_b200/csrc -o redback_b200/_lib/exp/prof.so redback_b200/csrc/capi.cu
