#!/bin/bash
# build a kernel variant for profiles/ab.py: profiles/build_variant.sh NAME [-DFLAG ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -o build/variants/$name.so p-pad_b200/csrc/ppad_kernels.cu
