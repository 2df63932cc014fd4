#!/bin/sh
# builds tools/microbench/overlap against the library objects
set -e
cd "$(dirname "$0")/../.."
C=hybrid-monte-carlo-for-d-wave-sc_b200/csrc
make -C $C -j8 >/dev/null
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude -I$C \
  -o tools/microbench/overlap tools/microbench/overlap.cu $C/api.o $C/assemble.o $C/hetrd.o $C/stedc.o $C/backtransform.o $C/gemm_dmma.o $C/force.o
