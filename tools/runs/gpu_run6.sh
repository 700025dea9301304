set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 600 python -m pytest tests/test_kdtree_gpu.py tests/test_evidence_gpu.py -x -q > gpurun_out/r2_run6_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run6_tests.log
timeout 300 python tools/stress_tree.py --seconds 45 --seed 3 > gpurun_out/r2_run6_stress.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run6_stress.log
timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run6_cfg3.json 2> gpurun_out/r2_run6_cfg3.err
for d in 2 4 8 32 64; do timeout 120 python tools/profile_tree.py 10000000 $d >> gpurun_out/r2_run6_dims.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_tree_full_v2c_launches.csv python tools/profile_tree.py 10000000 20 > gpurun_out/r2_run6_ncu.log 2>&1
echo finished
