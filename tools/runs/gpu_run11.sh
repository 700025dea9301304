set -x
cd $GRAFT_REPO_ROOT
export MCMC_GPU_SKIP_FULL_PARITY=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run11_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run11_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_run11_bench.json 2> gpurun_out/r2_run11_bench.err
timeout 300 python tools/dmma_decision.py gpurun_out/r02_dmma_decision.json > gpurun_out/r2_run11_dmma.log 2>&1
timeout 600 python tools/cfg4_batches.py --seeds 20 --out gpurun_out/r02_cfg4_batches_20seeds.json > gpurun_out/r2_run11_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_balanced -s 1 -c 1 -f -o gpurun_out/r02_mh_tma python bench.py --steps 1 --warmup 1 --no-evidence --no-rjmcmc --no-cpu > gpurun_out/r2_run11_ncu_mh.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v2_scatter -s 20 -c 1 -f -o gpurun_out/r02_v2_scatter python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run11_ncu_sc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v2_bottom -s 1 -c 1 -f -o gpurun_out/r02_v2_bottom python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run11_ncu_bt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v2_hist -s 60 -c 1 -f -o gpurun_out/r02_v2_hist python tools/profile_lebesgue.py 10000000 20 > gpurun_out/r2_run11_ncu_hi.log 2>&1
echo finished
