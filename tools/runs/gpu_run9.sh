set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 900 python -m pytest tests/test_mcmc_gpu.py tests/test_misc_gpu.py -x -q > gpurun_out/r2_run9_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run9_tests.log
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -k config2 > gpurun_out/r2_run9_tests_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run9_tests_full.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run9_bench.json 2> gpurun_out/r2_run9_bench.err
MCMC_GPU_MH_TMA=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run9_bench_notma.json 2>> gpurun_out/r2_run9_bench.err
timeout 300 python tools/dmma_decision.py gpurun_out/r02_dmma_decision.json > gpurun_out/r2_run9_dmma.log 2>&1
timeout 300 python tools/stress_parity.py --seconds 60 > gpurun_out/r2_run9_stress.log 2>&1
echo finished
