set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 900 python -m pytest tests/test_misc_gpu.py tests/test_mcmc_gpu.py -x -q > gpurun_out/r2_run35_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run35_tests.log
echo finished
