set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 600 python -m pytest tests/test_kdtree_gpu.py tests/test_evidence_gpu.py tests/test_misc_gpu.py -x -q > gpurun_out/r2_run12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run12_tests.log
timeout 300 python tools/stress_tree.py --seconds 40 --seed 6 > gpurun_out/r2_run12_stress.log 2>&1
timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run12_cfg3.json 2> gpurun_out/r2_run12_cfg3.err
MCMC_GPU_KD_SMEM_KB=110 timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run12_cfg3_smem110.json 2>> gpurun_out/r2_run12_cfg3.err
MCMC_GPU_KD_SMEM_KB=110 timeout 300 python tools/stress_tree.py --seconds 30 --seed 7 > gpurun_out/r2_run12_stress110.log 2>&1
echo finished
