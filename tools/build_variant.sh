#!/bin/bash
# developer tool: build variants/libemc_<name>.so from the working tree with extra nvcc flags (kernel A/B runs, tools/ab.sh)
# usage: tools/build_variant.sh <name> [-DEMC_SLIM -DFOO=1 ...]
set -e
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared "$@" \
  -o variants/libemc_$name.so erpl_monte_carlo_sim_b200/csrc/emc_engine.cu
echo built variants/libemc_$name.so
