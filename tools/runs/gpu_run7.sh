set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run7_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run7_tests.log
timeout 300 python tools/stress_tree.py --seconds 45 --seed 4 > gpurun_out/r2_run7_stress.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run7_stress.log
timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run7_cfg3.json 2> gpurun_out/r2_run7_cfg3.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_run7_bench.json 2> gpurun_out/r2_run7_bench.err
echo finished
