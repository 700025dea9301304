set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run10_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run10_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run10_bench.json 2> gpurun_out/r2_run10_bench.err
timeout 300 python tools/dmma_decision.py gpurun_out/r02_dmma_decision.json > gpurun_out/r2_run10_dmma.log 2>&1
timeout 300 python tools/stress_parity.py --seconds 60 --seed 5 > gpurun_out/r2_run10_stress.log 2>&1
echo finished
