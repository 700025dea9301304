set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 600 python -m pytest tests/test_kdtree_gpu.py tests/test_evidence_gpu.py -x -q > gpurun_out/r2_run5_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run5_tests.log
timeout 300 python tools/stress_tree.py --seconds 60 --seed 2 > gpurun_out/r2_run5_stress.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run5_stress.log
timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run5_cfg3.json 2> gpurun_out/r2_run5_cfg3.err
python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run5_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_lebesgue_v2b_launches.csv python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run5_ncu.log 2>&1
echo finished
