set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 300 python -m pytest tests/test_ellipse_gpu.py -x -q > gpurun_out/r2_run40_ellipse_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run40_ellipse_tests.log
timeout 300 python tools/bench_ellipse.py > gpurun_out/r2_run40_ellipse.json 2> gpurun_out/r2_run40_ellipse.err
timeout 300 python tools/bench_ellipse.py --n 1000000 --d 16 --cpu-n 30000 > gpurun_out/r2_run40_ellipse_d16.json 2>> gpurun_out/r2_run40_ellipse.err
echo finished
