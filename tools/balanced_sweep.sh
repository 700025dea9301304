#!/bin/bash
run() { env "$@" MCMC_GPU_DEBUG=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-evidence 2> gpurun_out/err.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2))"; grep "mean cycles" gpurun_out/err.log | tail -1; }
run MCMC_GPU_MH_SEG=128
run MCMC_GPU_MH_SEG=64
run MCMC_GPU_MH_SEG=256
run MCMC_GPU_MH_SEG=128 MCMC_GPU_MH_GRID=1480
run MCMC_GPU_MH_SEG=128 MCMC_GPU_MH_GRID=1850
