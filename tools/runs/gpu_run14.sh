set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run14_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run14_tests.log
timeout 300 python tools/stress_tree.py --seconds 60 --seed 14 > gpurun_out/r2_run14_stress.log 2>&1
timeout 300 python tools/bench_evidence.py --reps 3 > gpurun_out/r2_run14_cfg3.json 2> gpurun_out/r2_run14_cfg3.err
timeout 300 python tools/bench_evidence.py --reps 3 --dups 0.3 > gpurun_out/r2_run14_cfg3_dups.json 2>> gpurun_out/r2_run14_cfg3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_run14_lebesgue_launches.csv python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run14_ncu_l.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_run14_tree_launches.csv python tools/profile_tree.py 10000000 20 > gpurun_out/r2_run14_ncu_t.log 2>&1
echo finished
