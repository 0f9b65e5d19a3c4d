#!/bin/bash
# build a kernel variant of the whole library for profiles/ab.py: profiles/build_variant.sh NAME [-DFLAG ...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/variants/$name
mkdir -p $out
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
pids=""
for tu in ppad_api ppad_step ppad_agent ppad_policy; do
  nvcc $FLAGS "$@" -c -o $out/$tu.o p-pad_b200/csrc/$tu.cu & pids="$pids $!"
done
for g in 0 1 2 3 4 5 6 7; do
  nvcc $FLAGS "$@" -DPPAD_GEOM=$g -c -o $out/inst_g$g.o p-pad_b200/csrc/ppad_step_inst.cu & pids="$pids $!"
done
for p in $pids; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/$name.so $out/*.o
echo build/variants/$name.so
