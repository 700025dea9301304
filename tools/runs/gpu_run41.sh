set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 tools/multi_gpu_check.py --samples 2000000 --dim 8 --out gpurun_out/r02_multi_gpu_check_2_run41.json > gpurun_out/r2_run41_mg2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/bench_dist_build.py --check --reps 4 > gpurun_out/r2_run41_dist2.json 2> gpurun_out/r2_run41_dist2.err
echo finished
