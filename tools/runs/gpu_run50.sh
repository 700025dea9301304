set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_run50_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run50_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run50_tests.log
echo finished
