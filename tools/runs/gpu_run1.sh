set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_run1_env.log; nproc >> gpurun_out/r2_run1_env.log; free -g >> gpurun_out/r2_run1_env.log
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_run1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_run1_tests.log
timeout 300 python tools/cfg4_batches.py --seeds 6 --out gpurun_out/r02_cfg4_batches.json > gpurun_out/r2_run1_cfg4.log 2>&1
timeout 200 python tools/bench_rjmcmc.py --da 8 --db 16 --chains 262144 > gpurun_out/r2_run1_cfg5_8_16.json 2> gpurun_out/r2_run1_cfg5.err
timeout 200 python tools/bench_rjmcmc.py --da 8 --db 16 --chains 262144 --nstop 64 > gpurun_out/r2_run1_cfg5_8_16_nstop64.json 2>> gpurun_out/r2_run1_cfg5.err
timeout 200 python tools/bench_rjmcmc.py --da 2 --db 4 > gpurun_out/r2_run1_cfg5_2_4.json 2>> gpurun_out/r2_run1_cfg5.err
timeout 900 python tools/parity_full_size.py --n 10000000 --d 20 --dups 0.0 --out gpurun_out/r02_parity_cfg3_1e7.json > gpurun_out/r2_run1_parity.log 2>&1
echo finished
