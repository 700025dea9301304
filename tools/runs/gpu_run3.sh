set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 600 python -m pytest tests/test_kdtree_gpu.py tests/test_evidence_gpu.py -x -q > gpurun_out/r2_run3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run3_tests.log
timeout 300 python tools/stress_tree.py --seconds 100 > gpurun_out/r2_run3_stress.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run3_stress.log
timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run3_cfg3.json 2> gpurun_out/r2_run3_cfg3.err
MCMC_GPU_KD_BUILD=1 timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run3_cfg3_v1.json 2>> gpurun_out/r2_run3_cfg3.err
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -k "config3" > gpurun_out/r2_run3_tests_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run3_tests_full.log
echo finished
