set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 600 python -m pytest tests/test_mcmc_gpu.py -x -q > gpurun_out/r2_run31_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run31_tests.log
for sg in 1 2 4 8; do
MCMC_GPU_D2H_SEGMENTS=$sg timeout 300 python bench.py --steps 5 --warmup 3 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run31_bench_seg$sg.json 2> gpurun_out/r2_run31_bench.err
done
echo finished
