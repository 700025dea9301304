set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 600 python -m pytest tests/test_mcmc_gpu.py tests/test_misc_gpu.py -x -q > gpurun_out/r2_run29_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run29_tests.log
timeout 600 python tools/bench_cfg1.py > gpurun_out/r2_run29_cfg1.json 2> gpurun_out/r2_run29_cfg1.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run29_bench.json 2> gpurun_out/r2_run29_bench.err
echo finished
