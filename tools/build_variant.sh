#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]   -> variants/lib_<name>.so (+ ptxas log)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xptxas -v \
  -Xcompiler -fPIC -shared -ccbin /usr/bin/g++ "$@" -o variants/lib_${name}.so \
  tfhe-research_b200/csrc/tfhe_b200.cu tfhe-research_b200/csrc/host_api.cpp tfhe-research_b200/csrc/tfhe_mgpu.cpp -ldl 2> variants/lib_${name}.ptxas.log
grep -A2 "pbs_kernel" variants/lib_${name}.ptxas.log | grep -E "registers|spill" | paste - - | sed 's/ptxas info *: //g' | cut -c1-200
