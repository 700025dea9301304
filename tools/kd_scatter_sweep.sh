#!/bin/bash
# Resident CTAs per SM of the kd-tree's scatter kernel (register cap through __launch_bounds__): rebuild
# kdtree_build2.cu with MG_V2_SCATTER_MINBLOCKS and time config 3.  Run under gpurun after `make`; needs nvcc on the box.
cd "$(dirname "$0")/.."
CS=mcmc_ocaml_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --extended-lambda -Xcompiler -fPIC"
OBJS=$(ls $CS/build/*.o | grep -v kdtree_build2.o)
for mb in 2 3 4 5; do
  nvcc $FLAGS -DMG_V2_SCATTER_MINBLOCKS=$mb -c $CS/kdtree_build2.cu -o gpurun_out/kb2.o 2>/dev/null || { echo "minblocks $mb: does not compile"; continue; }
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_out/libkb2.so gpurun_out/kb2.o $OBJS -ldl
  MCMC_GPU_LIB=$PWD/gpurun_out/libkb2.so python tools/bench_evidence.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('minblocks $mb', 'lebesgue', round(1e3*d['lebesgue_s'],2), 'tree64', round(1e3*d['tree64_s'],2), 'full', round(1e3*d['tree_full_s'],2), 'Z', d['lebesgue_Z'])"
done
rm -f gpurun_out/kb2.o gpurun_out/libkb2.so
