set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run43_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run43_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_run43_smoke.log 2>&1
( time timeout 900 python bench.py > gpurun_out/r2_run43_bench.json 2> gpurun_out/r2_run43_bench.err ) 2> gpurun_out/r2_run43_bench.time
echo finished
