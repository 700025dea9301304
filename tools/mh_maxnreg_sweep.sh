#!/bin/bash
# Register cap of the balanced D = 10 sampler (three one-warp CTAs per scheduler leave room for 170 registers).
set -e
cd "$(dirname "$0")/.."
CS=mcmc_ocaml_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --extended-lambda -Xcompiler -fPIC"
OBJS=$(ls $CS/build/*.o | grep -v mcmc_static_10.o)
for r in 128 136 144 152 160 168; do
  nvcc $FLAGS -DMG_SD=10 "-DMG_MHB_MAXNREG(D)=$r" -c $CS/mcmc_static.cu -o gpurun_out/abl.o 2>/dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_out/libabl.so gpurun_out/abl.o $OBJS -ldl
  for rep in 1 2; do
  MCMC_GPU_LIB=$PWD/gpurun_out/libabl.so python bench.py --steps 8 --warmup 3 --no-cpu --no-evidence --no-rjmcmc 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maxnreg $r', 'ms', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'clk', d['clocks']['sm_mhz'])"
  done
done
rm -f gpurun_out/abl.o gpurun_out/libabl.so
