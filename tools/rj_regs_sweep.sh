#!/bin/bash
# RJ ensemble kernel: register cap vs throughput (latency-bound tree descents want more resident warps).
set -e
cd "$(dirname "$0")/.."
CS=mcmc_ocaml_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --extended-lambda -Xcompiler -fPIC"
OBJS=$(ls $CS/build/*.o | grep -v "/rj.o")
for r in 255 128 96 80 64; do
  nvcc $FLAGS "-DMG_RJ_MAXNREG(D)=$r" -Xptxas -v -c $CS/rj.cu -o gpurun_out/rjx.o 2>&1 | grep -A2 "rj_ensemble_kernelILi4E" | grep -E "spill|Used" | tr '\n' ' '
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_out/librjx.so gpurun_out/rjx.o $OBJS -ldl
  MCMC_GPU_LIB=$PWD/gpurun_out/librjx.so python tools/bench_rjmcmc.py 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maxnreg $r', 'rj_s', round(d['rj_s'],4), 'steps/s %.4g' % d['chain_steps_per_s'], d['counts'])"
done
rm -f gpurun_out/rjx.o gpurun_out/librjx.so
