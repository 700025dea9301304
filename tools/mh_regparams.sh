#!/bin/bash
# Parameters of the D = 10 step kept in registers instead of re-read from shared memory every step (proposal pairs,
# mu), at 128 and 168 registers: rebuild the static instantiation with the switch and time the bench pass.
# Run under gpurun after `make -C mcmc_ocaml_b200/csrc`; needs nvcc on the box.
set -e
cd "$(dirname "$0")/.."
CS=mcmc_ocaml_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --extended-lambda -Xcompiler -fPIC"
OBJS=$(ls $CS/build/*.o | grep -v mcmc_static_10.o)
for v in BASE "-DMG_EXP_REG_PROP" "-DMG_EXP_REG_PROP -DMG_EXP_REG_MU" "-DMG_EXP_REG_PROP -DMG_MH_MAXNREG(D)=168" "-DMG_EXP_REG_PROP -DMG_EXP_REG_MU -DMG_MH_MAXNREG(D)=168" "-DMG_MH_MAXNREG(D)=168"; do
  def=$v; [ "$v" = BASE ] && def=""
  nvcc $FLAGS -DMG_SD=10 $def -c $CS/mcmc_static.cu -o gpurun_out/abl.o 2>/dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_out/libabl.so gpurun_out/abl.o $OBJS -ldl
  MCMC_GPU_LIB=$PWD/gpurun_out/libabl.so python bench.py --steps 5 --warmup 3 --no-cpu --no-evidence --no-rjmcmc 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'accept', d['accept_rate_per_rank'], 'clk', d['clocks']['sm_mhz'])"
done
rm -f gpurun_out/abl.o gpurun_out/libabl.so
