set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run34_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run34_tests.log
( time timeout 900 python bench.py > gpurun_out/r2_run34_bench.json 2> gpurun_out/r2_run34_bench.err ) 2> gpurun_out/r2_run34_bench.time
timeout 300 python tools/bench_cfg1.py > gpurun_out/r2_run34_cfg1.json 2> gpurun_out/r2_run34_cfg1.err
timeout 300 python tools/bench_ellipse.py > gpurun_out/r2_run34_ellipse.json 2> gpurun_out/r2_run34_ellipse.err
echo finished
