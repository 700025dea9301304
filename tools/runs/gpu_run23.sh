set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/bench_dist_build.py --check --reps 6 > gpurun_out/r2_run23_dist8.json 2> gpurun_out/r2_run23_dist8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/bench_dist_build.py --check --reps 4 --min-split 64 > gpurun_out/r2_run23_dist8_ms64.json 2> gpurun_out/r2_run23_dist8_ms64.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 tools/multi_gpu_check.py --samples 10000000 --dim 20 --out gpurun_out/r02_multi_gpu_check_8_run23.json > gpurun_out/r2_run23_mg8.log 2>&1
echo finished
