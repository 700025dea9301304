set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 300 python tools/stress_tree.py --seconds 60 --seed 45 > gpurun_out/r2_run45_stress.log 2>&1
timeout 600 python -m pytest tests/test_kdtree_gpu.py tests/test_evidence_gpu.py -x -q > gpurun_out/r2_run45_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run45_tests.log
timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run45_cfg3.json 2> gpurun_out/r2_run45_cfg3.err
timeout 300 python tools/bench_evidence.py --reps 3 --dups 0.3 > gpurun_out/r2_run45_cfg3_dups.json 2>> gpurun_out/r2_run45_cfg3.err
timeout 120 python tools/bench_truncated.py 10000000 4 2 > gpurun_out/r2_run45_d4.log 2>&1
echo finished
