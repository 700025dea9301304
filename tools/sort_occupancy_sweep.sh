#!/bin/bash
# Radix-sort scatter kernel: keys per tile (rounds of 32 per warp) and resident CTAs per SM vs kd-tree / evidence time.
set -e
cd "$(dirname "$0")/.."
CS=mcmc_ocaml_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --extended-lambda -Xcompiler -fPIC"
OBJS=$(ls $CS/build/*.o | grep -v "/kdtree.o" | grep -v "/evidence.o")
for v in "8 4" "4 4" "4 6" "6 5" "12 3"; do
  set -- $v
  nvcc $FLAGS -DMG_RS_ROUNDS=$1 -DMG_RS_MINBLOCKS=$2 -Xptxas -v -c $CS/kdtree.cu -o gpurun_out/kdx.o 2>&1 | grep -A2 "rs_scatter_kernel" | grep -E "Used" | tr '\n' ' '
  nvcc $FLAGS -DMG_RS_ROUNDS=$1 -DMG_RS_MINBLOCKS=$2 -c $CS/evidence.cu -o gpurun_out/evx.o 2>/dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_out/libkdx.so gpurun_out/kdx.o gpurun_out/evx.o $OBJS -ldl
  MCMC_GPU_LIB=$PWD/gpurun_out/libkdx.so python tools/bench_evidence.py --reps 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rounds $1 minblocks $2', 'lebesgue_s', round(d['lebesgue_s'],4), 'direct_s', round(d['direct_s'],4), 'tree64_s', round(d['tree64_s'],4), 'tree_full_s', round(d['tree_full_s'],4))"
done
rm -f gpurun_out/kdx.o gpurun_out/evx.o gpurun_out/libkdx.so
